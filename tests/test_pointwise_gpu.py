"""GPU parity of the memory-bound kernels (through the C ABI) against torch ops on identical inputs."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import bf16_round, from_internal, rel_l2, rel_to_max, strict_fp32, to_internal

pytestmark = pytest.mark.gpu


def _ops():
    from unet_bssfp_b200 import ops
    return ops


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, device="cuda", generator=g) * scale


@pytest.mark.parametrize("ca,cb,shape", [(24, 0, (2, 4, 6, 10)), (24, 6, (1, 8, 8, 8)), (6, 6, (2, 2, 4, 4)), (6, 0, (1, 3, 5, 7))])
def test_pack_unpack_roundtrip(ca, cb, shape):
    ops = _ops()
    n, d, h, w = shape
    a = _rand((n, ca, d, h, w), 0)
    b = _rand((n, cb, d, h, w), 1) if cb else None
    pk = ops.pack_ncdhw(a, b)
    assert pk.shape == (n, d, h, w, 32) and pk.dtype == torch.bfloat16
    ref = torch.cat([a, b], 1) if cb else a
    assert torch.equal(from_internal(pk, ca + cb), bf16_round(ref))          # bit-exact bf16 rounding
    assert pk[..., ca + cb:].abs().max().item() == 0.0
    assert torch.equal(ops.unpack_ncdhw(pk, ca + cb), bf16_round(ref))
    if cb:
        assert torch.equal(ops.unpack_ncdhw(pk, cb, ca), bf16_round(b))


@pytest.mark.parametrize("ca,cb,shape", [(24, 6, (2, 4, 6, 10)), (6, 6, (1, 8, 8, 8)), (24, 0, (1, 2, 4, 18))])
def test_pack_space_to_depth(ca, cb, shape):
    """s2d layout [n][(pd,ph,pw)][d/2][h/2][w/2][cp] == plain pack regrouped by parity (bit-exact)."""
    ops = _ops()
    n, d, h, w = shape
    a = _rand((n, ca, d, h, w), 0)
    b = _rand((n, cb, d, h, w), 1) if cb else None
    plain = ops.pack_ncdhw(a, b)                                      # (n,d,h,w,32)
    s2d = ops.pack_ncdhw(a, b, s2d=True)
    assert s2d.shape == (n, 8, d // 2, h // 2, w // 2, 32)
    want = plain.view(n, d // 2, 2, h // 2, 2, w // 2, 2, 32).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(s2d.shape)
    assert torch.equal(s2d, want)


@pytest.mark.parametrize("c,shape", [(32, (2, 4, 6, 10)), (64, (1, 8, 8, 8)), (256, (2, 2, 4, 2))])
def test_to_space_to_depth_copy(c, shape):
    """ub_to_s2d == the plain tensor regrouped by parity (bit-exact)."""
    ops = _ops()
    n, d, h, w = shape
    a = to_internal(bf16_round(_rand((n, c, d, h, w), 4)))
    got = ops.to_s2d(a)
    want = a.view(n, d // 2, 2, h // 2, 2, w // 2, 2, c).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(got.shape)
    assert got.shape == (n, 8, d // 2, h // 2, w // 2, c) and torch.equal(got, want)


@pytest.mark.parametrize("mode", ["instance", "batch_train", "batch_eval"])
@pytest.mark.parametrize("c,shape,pool", [(32, (2, 4, 16, 8), True), (24, (3, 4, 8, 8), False), (64, (1, 8, 8, 16), True)])
def test_norm_act_forward_backward(mode, c, shape, pool):
    """conv stats -> finalize -> norm+LeakyReLU(+pool) and the whole backward, vs torch autograd."""
    strict_fp32()
    ops = _ops()
    from unet_bssfp_b200 import _lib
    n, d, h, w = shape
    cp = (c + 31) // 32 * 32
    slope = 0.1
    y = bf16_round(_rand((n, c, d, h, w), 3) * 1.7 + 0.3)
    gamma = _rand((c,), 4) * 0.5 + 1.0
    beta = _rand((c,), 5) * 0.2
    rm = _rand((c,), 6) * 0.1
    rv = _rand((c,), 7).abs() + 0.5
    # statistics partials exactly as the conv epilogue would emit them: use a 1x1x1 identity conv
    spec = ops.ConvSpec(_lib.UB_CONV_K1, c, c)
    eye = torch.eye(c, device="cuda").view(c, c, 1, 1, 1)
    yi, stats = ops.conv_fwd(spec, to_internal(y), None, ops.pack_conv_weights(spec, eye, 0), None, want_stats=True)
    assert yi.dtype == torch.float16 and (from_internal(yi, c) - y).abs().max().item() < 1e-6   # y is stored as fp16
    imode = {"instance": _lib.UB_NORM_INSTANCE, "batch_train": _lib.UB_NORM_BATCH_TRAIN, "batch_eval": _lib.UB_NORM_BATCH_EVAL}[mode]
    rm2, rv2 = rm.clone(), rv.clone()
    scale, shift, mean, rstd = ops.norm_finalize(stats, n, d * h * w, cp, c, gamma, beta, 1e-5, imode, 0.1, rm2, rv2)
    a, pooled = ops.norm_act_fwd(yi, scale, shift, slope, pool=pool)

    yr = y.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm3, rv3 = rm.clone(), rv.clone()
    if mode == "instance":
        z = F.instance_norm(yr, weight=gr, bias=br, eps=1e-5)
    else:
        z = F.batch_norm(yr, rm3, rv3, gr, br, training=(mode == "batch_train"), momentum=0.1, eps=1e-5)
    ar = F.leaky_relu(z, slope)
    assert rel_to_max(from_internal(a, c), ar) < 6e-3
    if mode == "batch_train":
        assert rel_to_max(rm2, rm3) < 1e-5 and rel_to_max(rv2, rv3) < 1e-5
    out_r = F.max_pool3d(ar, 2) if pool else ar
    if pool:
        # pooled = max over the stored (bf16) activations
        assert torch.equal(from_internal(pooled, c), F.max_pool3d(from_internal(a, c), 2))
    # backward with a smooth upstream gradient
    dO = bf16_round(_rand(out_r.shape, 8))
    out_r.backward(dO)
    dOi = to_internal(dO)
    dA = ops.maxpool_bwd(a, dOi) if pool else dOi
    dy, dgamma, dbeta, dbias = ops.norm_act_bwd(dA, a, yi, imode, mean, rstd, scale, slope, 0.0, 0, c)
    torch.cuda.synchronize()
    # sign flips of LeakyReLU at bf16-rounded zeros are possible but rare at these sizes
    assert rel_l2(from_internal(dy, c), yr.grad) < 2e-2
    assert rel_l2(dgamma, gr.grad) < 2e-2
    assert rel_l2(dbeta, br.grad) < 2e-2
    if mode != "batch_eval":
        assert dbias.abs().max().item() == 0.0
    # the product path never reads `a`: LeakyReLU branch recomputed from sign(y * scale + shift)
    dy2, dgamma2, dbeta2, _ = ops.norm_act_bwd(dA, None, yi, imode, mean, rstd, scale, slope, 0.0, 0, c, shift=shift)
    assert torch.equal(dy2, dy) and torch.equal(dgamma2, dgamma) and torch.equal(dbeta2, dbeta)


def test_maxpool_bwd_ties_and_accumulate():
    ops = _ops()
    n, c, d, h, w = 1, 32, 4, 4, 4
    a = torch.zeros((n, c, d, h, w), device="cuda")            # all ties: first voxel of each window wins
    a[:, :, 1::2, 1::2, 1::2] = 0.0
    ai = to_internal(a)
    dP = bf16_round(_rand((n, c, d // 2, h // 2, w // 2), 2))
    dA = ops.maxpool_bwd(ai, to_internal(dP))
    ar = a.clone().requires_grad_(True)
    F.max_pool3d(ar, 2).backward(dP)
    assert torch.equal(from_internal(dA, c), ar.grad)
    base = bf16_round(_rand((n, c, d, h, w), 3))
    acc = ops.maxpool_bwd(ai, to_internal(dP), to_internal(base))
    assert rel_to_max(from_internal(acc, c), base + ar.grad) < 5e-3


@pytest.mark.parametrize("mode,drop_p", [("instance", 0.05), ("instance", 0.0), ("batch_train", 0.0)])
@pytest.mark.parametrize("c,shape,accumulate", [(32, (2, 8, 16, 8), True), (64, (2, 4, 8, 8), True), (128, (1, 4, 4, 8), False),
                                                (24, (1, 36, 40, 24), True)])
def test_maxpool_bwd_fused_with_norm_backward_reduction(mode, drop_p, c, shape, accumulate):
    """ub_maxpool_bwd_fused: the gradient it writes is bit-identical to ub_maxpool_bwd on the materialised
    activations, and ub_norm_act_bwd fed with its partial records agrees with the separate reduction pass
    (different summation order only)."""
    strict_fp32()
    ops = _ops()
    from unet_bssfp_b200 import _lib
    n, d, h, w = shape
    cp = (c + 31) // 32 * 32
    g = torch.Generator(device="cuda").manual_seed(11)
    y = to_internal(torch.randn((n, c, d, h, w), device="cuda", generator=g) * 1.3 + 0.2, dtype=torch.float16)
    gamma = torch.rand((c,), device="cuda", generator=g) + 0.5
    beta = torch.randn((c,), device="cuda", generator=g) * 0.2
    yf = from_internal(y, c)
    imode = {"instance": _lib.UB_NORM_INSTANCE, "batch_train": _lib.UB_NORM_BATCH_TRAIN}[mode]
    dims = (2, 3, 4) if mode == "instance" else (0, 2, 3, 4)
    mean_t = yf.mean(dims, keepdim=True).expand(n, c, 1, 1, 1).reshape(n, c)
    var_t = yf.var(dims, unbiased=False, keepdim=True).expand(n, c, 1, 1, 1).reshape(n, c)
    rstd_t = (var_t + 1e-5).rsqrt()
    pad = lambda t, v=0.0: torch.cat([t, torch.full((n, cp - c), v, device="cuda")], 1).contiguous()
    mean, rstd = pad(mean_t), pad(rstd_t)
    scale, shift = pad(gamma * rstd_t), pad(beta - mean_t * gamma * rstd_t)
    seed = 777
    a, _ = ops.norm_act_fwd(y, scale, shift, 0.1, drop_p, seed)
    dP = to_internal(bf16_round(torch.randn((n, c, d // 2, h // 2, w // 2), device="cuda", generator=g)))
    base = to_internal(bf16_round(torch.randn((n, c, d, h, w), device="cuda", generator=g))) if accumulate else None
    want = ops.maxpool_bwd(a, dP, base.clone() if accumulate else None)
    fuse = ops.NormBwdFusion(y, scale, shift, mean, rstd, 0.1, drop_p, seed)
    assert ops.maxpool_bwd_fuse_records(n, d, h, w, cp) > 0
    got, partial = ops.maxpool_bwd_fused(dP, base.clone() if accumulate else None, fuse)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    ref = ops.norm_act_bwd(want, None, y, imode, mean, rstd, scale, 0.1, drop_p, seed, c, shift=shift)
    fus = ops.norm_act_bwd(got, None, y, imode, mean, rstd, scale, 0.1, drop_p, seed, c, shift=shift, partial=partial)
    torch.cuda.synchronize()
    assert rel_l2(fus[1], ref[1]) < 1e-4 and rel_l2(fus[2], ref[2]) < 1e-4          # dgamma, dbeta
    assert rel_l2(fus[0].float(), ref[0].float()) < 1e-3
    assert rel_to_max(fus[0].float(), ref[0].float()) < 1e-2


@pytest.mark.parametrize("drop_p", [0.0, 0.05])
@pytest.mark.parametrize("c,shape", [(32, (2, 8, 16, 8)), (64, (1, 5, 7, 9)), (512, (2, 2, 2, 2)), (24, (1, 33, 16, 20))])
def test_lean_norm_backward_apply_equals_generic_kernel(drop_p, c, shape, monkeypatch):
    """The instruction-lean apply kernel (statistics present, sign from y) is bit-identical to the generic one."""
    ops = _ops()
    from unet_bssfp_b200 import _lib
    n, d, h, w = shape
    cp = (c + 31) // 32 * 32
    g = torch.Generator(device="cuda").manual_seed(3)
    y = to_internal(torch.randn((n, c, d, h, w), device="cuda", generator=g), dtype=torch.float16)
    dA = to_internal(bf16_round(torch.randn((n, c, d, h, w), device="cuda", generator=g)))
    mean, rstd = torch.randn((n, cp), device="cuda", generator=g) * 0.1, torch.rand((n, cp), device="cuda", generator=g) + 0.5
    scale, shift = torch.rand((n, cp), device="cuda", generator=g) + 0.5, torch.randn((n, cp), device="cuda", generator=g) * 0.1
    args = (dA, None, y, _lib.UB_NORM_INSTANCE, mean, rstd, scale, 0.1, drop_p, 99, c)
    lean = ops.norm_act_bwd(*args, shift=shift)
    monkeypatch.setenv("UB_NAB_GENERIC", "1")
    generic = ops.norm_act_bwd(*args, shift=shift)
    torch.cuda.synchronize()
    assert torch.equal(lean[0], generic[0])
    assert torch.equal(lean[1], generic[1]) and torch.equal(lean[2], generic[2])


def test_dropout_statistics_and_backward_mask():
    ops = _ops()
    from unet_bssfp_b200 import _lib
    n, c, d, h, w = 2, 32, 16, 16, 16
    y = to_internal(torch.ones((n, c, d, h, w), device="cuda"), dtype=torch.float16)
    p = 0.05
    a, _ = ops.norm_act_fwd(y, None, None, 0.1, drop_p=p, drop_seed=1234)
    af = a.float()
    keep = (af != 0)
    rate = 1.0 - keep.float().mean().item()
    assert abs(rate - p) < 2e-3                                      # keep-rate
    assert rel_to_max(af[keep], torch.full_like(af[keep], 1.0 / (1.0 - p))) < 4e-3   # inverted scaling
    a2, _ = ops.norm_act_fwd(y, None, None, 0.1, drop_p=p, drop_seed=1234)
    assert torch.equal(a, a2)                                        # counter-based: reproducible
    a3, _ = ops.norm_act_fwd(y, None, None, 0.1, drop_p=p, drop_seed=99)
    assert not torch.equal(a, a3)
    # backward regenerates the same mask
    dA = to_internal(torch.ones((n, c, d, h, w), device="cuda"))
    dy, *_ = ops.norm_act_bwd(dA, a, None, _lib.UB_NORM_NONE, None, None, None, 0.1, p, 1234, c)
    assert torch.equal(dy.float() != 0, keep)


def test_colsum():
    ops = _ops()
    x = bf16_round(_rand((3, 24, 4, 8, 8), 1))
    got = ops.colsum(to_internal(x), 24)
    assert rel_to_max(got, x.sum(dim=(0, 2, 3, 4))) < 1e-5


def test_l1_and_bce_losses():
    from unet_bssfp_b200 import L1Loss, BCEWithLogitsLoss
    a = _rand((2, 6, 16, 16, 16), 0).requires_grad_(True)
    b = _rand((2, 6, 16, 16, 16), 1)
    l = L1Loss()(a, b)
    (l * 50.0).backward()
    a2 = a.detach().clone().requires_grad_(True)
    l2 = F.l1_loss(a2, b)
    (l2 * 50.0).backward()
    assert abs(l.item() - l2.item()) < 1e-6 * abs(l2.item()) + 1e-7
    assert torch.allclose(a.grad, a2.grad, rtol=1e-6, atol=0)
    for tv in (0.0, 1.0):
        z = _rand((8, 1, 4, 4, 4), 2, 3.0).requires_grad_(True)
        t = torch.full_like(z, tv)
        lb = BCEWithLogitsLoss()(z, t)
        (lb * 0.5).backward()
        z2 = z.detach().clone().requires_grad_(True)
        lb2 = F.binary_cross_entropy_with_logits(z2, t)
        (lb2 * 0.5).backward()
        assert abs(lb.item() - lb2.item()) < 1e-5 * abs(lb2.item())
        assert torch.allclose(z.grad, z2.grad, rtol=1e-4, atol=1e-8)


@pytest.mark.parametrize("co,ci,shape", [(6, 32, (2, 4, 6, 10)), (1, 32, (1, 3, 5, 7)), (8, 24, (1, 4, 4, 8))])
def test_output_head_conv1x1_fused(co, ci, shape):
    """ub_conv1x1_to_ncdhw / ub_conv1x1_from_ncdhw_bwd vs F.conv3d(k=1) + autograd on the bf16-rounded input
    (fp32 weights, fp32 math): forward exact to fp32 rounding, du to bf16 rounding, dW / db to 1e-4."""
    strict_fp32()
    ops = _ops()
    n, d, h, w = shape
    u = bf16_round(_rand((n, ci, d, h, w), 0))
    wt = _rand((co, ci, 1, 1, 1), 1) * 0.3
    b = _rand((co,), 2)
    out = ops.conv1x1_to_ncdhw(to_internal(u), wt, b)
    ur, wr, br = u.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv3d(ur, wr, br)
    assert out.shape == ref.shape and rel_to_max(out, ref) < 1e-5
    g = _rand(ref.shape, 3)
    ref.backward(g)
    du, dw, db = ops.conv1x1_from_ncdhw_bwd(g, to_internal(u), wt)
    assert rel_to_max(from_internal(du, ci), ur.grad) < 6e-3
    if ci < 32:
        assert du[..., ci:].abs().max().item() == 0.0
    assert rel_to_max(dw, wr.grad) < 1e-4 and rel_to_max(db, br.grad) < 1e-4
    du2, dw2, db2 = ops.conv1x1_from_ncdhw_bwd(g, to_internal(u), wt, need_input=True, need_params=False)
    assert torch.equal(du2, du) and dw2 is None and db2 is None
