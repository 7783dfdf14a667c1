"""Deferred activations: norm + dropout + LeakyReLU applied on the CONSUMER's operand path (VERDICT r1 item 1;
reference semantics: monai Convolution -> ADN("NDA") instantiated at ref:src/model.py:22-28).

A conv -> norm block keeps only its raw fp16 output y and per-(n, c) constants; the consumers -- the marching
forward conv (shared-memory operand transform between TMA arrival and tcgen05.mma), the marching weight gradient,
the pooling forward / backward passes and the fused 1x1x1 output head -- evaluate the activations themselves with
ONE shared device function. So every consumer must give the SAME BITS as when it is fed the materialised tensor
``ub_norm_act_fwd`` writes: that is what these tests assert, op by op and for the whole generator; the materialised
path itself is checked against torch in test_pointwise_gpu.py / test_conv_gpu.py / test_model_gpu.py.
"""
import pytest
import torch

from tests.util import no_dropout, rel_l2, to_internal

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    import unet_bssfp_b200 as ub
    return ub.ops


def _producer(n, d, h, w, c=32, seed=0, drop_p=0.0, slope=0.1):
    """A producer block's saved state: y (fp16, c real channels of 32), scale / shift [n][32], and both forms of its
    activations (deferred, materialised)."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(seed)
    y = to_internal(torch.randn((n, c, d, h, w), device=DEV, generator=g) * 1.5 + 0.25, dtype=torch.float16)
    scale = torch.zeros((n, 32), device=DEV)
    shift = torch.zeros((n, 32), device=DEV)
    scale[:, :c] = torch.rand((n, c), device=DEV, generator=g) + 0.5
    shift[:, :c] = torch.randn((n, c), device=DEV, generator=g) * 0.3
    seed_d = 4711 + seed
    a, _ = ops.norm_act_fwd(y, scale, shift, slope, drop_p, seed_d)
    return ops.DeferredAct(y, scale, shift, slope, drop_p, seed_d), a


CONV_CASES = [
    # c0, c1, co, (n, d, h, w)
    (32, 0, 32, (2, 9, 32, 16)),      # one K chunk, single-CTA marching kernel, several d segments
    (32, 0, 32, (1, 6, 20, 12)),      # ragged h / w tiles: halo rows outside the volume must stay zero AFTER activation
    (24, 0, 32, (1, 5, 16, 8)),       # 24 real channels (the input head's output): padded channels stay inert
    (32, 64, 32, (2, 7, 32, 32)),     # skip concat [deferred | plain], CTA-pair kernel (two w tiles per cluster)
    (32, 64, 32, (1, 4, 16, 24)),     # skip concat, odd number of w tiles: single-CTA kernel with three chunks
    (32, 32, 32, (1, 5, 16, 16)),     # two chunks
]


@pytest.mark.parametrize("drop_p", [0.0, 0.05])
@pytest.mark.parametrize("c0,c1,co,shape", CONV_CASES)
def test_conv_forward_on_deferred_source_is_bit_identical(c0, c1, co, shape, drop_p):
    ops = _ops()
    n, d, h, w = shape
    spec = ops.ConvSpec(0, c0, co, c1)
    assert ops.deferred_src0_ok(spec, n, d, h, w)
    lazy, a = _producer(n, d, h, w, c=c0, seed=1, drop_p=drop_p)
    g = torch.Generator(device=DEV).manual_seed(2)
    s1 = to_internal(torch.randn((n, c1, d, h, w), device=DEV, generator=g)) if c1 else None
    wt = torch.randn((co, c0 + c1, 3, 3, 3), device=DEV, generator=g) / ((c0 + c1) * 27) ** 0.5
    b = torch.randn((co,), device=DEV, generator=g)
    wpk = ops.pack_conv_weights(spec, wt, 0)
    ref, ref_stats = ops.conv_fwd(spec, a, s1, wpk, b, want_stats=True)
    got, got_stats = ops.conv_fwd(spec, lazy, s1, wpk, b, want_stats=True)
    torch.cuda.synchronize()
    assert got.dtype == torch.float16
    assert torch.equal(got, ref)
    assert torch.equal(got_stats, ref_stats)
    # and without statistics (bf16 output)
    ref2, _ = ops.conv_fwd(spec, a, s1, wpk, b)
    got2, _ = ops.conv_fwd(spec, lazy, s1, wpk, b)
    assert got2.dtype == torch.bfloat16 and torch.equal(got2, ref2)


@pytest.mark.parametrize("drop_p", [0.0, 0.05])
@pytest.mark.parametrize("c0,c1,co,shape", CONV_CASES)
def test_conv_wgrad_on_deferred_source_is_bit_identical(c0, c1, co, shape, drop_p):
    ops = _ops()
    n, d, h, w = shape
    spec = ops.ConvSpec(0, c0, co, c1)
    lazy, a = _producer(n, d, h, w, c=c0, seed=3, drop_p=drop_p)
    g = torch.Generator(device=DEV).manual_seed(4)
    s1 = to_internal(torch.randn((n, c1, d, h, w), device=DEV, generator=g)) if c1 else None
    dy = to_internal(torch.randn((n, co, d, h, w), device=DEV, generator=g))
    shape_w = (co, c0 + c1, 3, 3, 3)
    ref = ops.conv_wgrad(spec, a, s1, dy, shape_w)
    got = ops.conv_wgrad(spec, lazy, s1, dy, shape_w)
    torch.cuda.synchronize()
    assert torch.equal(got, ref)


def test_deferred_source_is_refused_where_unsupported():
    ops = _ops()
    lazy, a = _producer(1, 4, 16, 8, seed=5)
    spec = ops.ConvSpec(0, 32, 64)                      # 64 output channels: generic kernel, no operand transform
    assert not ops.deferred_src0_ok(spec, 1, 4, 16, 8)
    wpk = ops.pack_conv_weights(spec, torch.randn((64, 32, 3, 3, 3), device=DEV), 0)
    with pytest.raises(RuntimeError, match="deferred"):
        ops.conv_fwd(spec, lazy, None, wpk, None)
    with pytest.raises(RuntimeError, match="deferred"):
        ops.conv_wgrad(spec, lazy, None, to_internal(torch.randn((1, 64, 4, 16, 8), device=DEV)), (64, 32, 3, 3, 3))


@pytest.mark.parametrize("drop_p", [0.0, 0.05])
@pytest.mark.parametrize("shape", [(2, 4, 16, 8), (1, 6, 10, 12)])
def test_pooling_passes_on_deferred_activations(shape, drop_p):
    ops = _ops()
    n, d, h, w = shape
    lazy, a = _producer(n, d, h, w, seed=6, drop_p=drop_p)
    a2, pooled = ops.norm_act_fwd(lazy.y, lazy.scale, lazy.shift, lazy.slope, drop_p, lazy.drop_seed, pool=True)
    none, pooled_only = ops.norm_act_fwd(lazy.y, lazy.scale, lazy.shift, lazy.slope, drop_p, lazy.drop_seed, pool=True,
                                         materialize=False)
    assert none is None and torch.equal(a2, a) and torch.equal(pooled_only, pooled)
    g = torch.Generator(device=DEV).manual_seed(7)
    dP = to_internal(torch.randn((n, 32, d // 2, h // 2, w // 2), device=DEV, generator=g))
    base = to_internal(torch.randn((n, 32, d, h, w), device=DEV, generator=g))
    assert torch.equal(ops.maxpool_bwd(lazy, dP), ops.maxpool_bwd(a, dP))
    assert torch.equal(ops.maxpool_bwd(lazy, dP, base.clone()), ops.maxpool_bwd(a, dP, base.clone()))


@pytest.mark.parametrize("drop_p", [0.0, 0.05])
def test_output_head_on_deferred_activations(drop_p):
    ops = _ops()
    n, d, h, w = 2, 4, 6, 10
    lazy, a = _producer(n, d, h, w, seed=8, drop_p=drop_p)
    g = torch.Generator(device=DEV).manual_seed(9)
    wt = torch.randn((6, 32, 1, 1, 1), device=DEV, generator=g) * 0.3
    b = torch.randn((6,), device=DEV, generator=g)
    assert torch.equal(ops.conv1x1_to_ncdhw(lazy, wt, b), ops.conv1x1_to_ncdhw(a, wt, b))
    go = torch.randn((n, 6, d, h, w), device=DEV, generator=g)
    du, dw, db = ops.conv1x1_from_ncdhw_bwd(go, lazy, wt)
    du2, dw2, db2 = ops.conv1x1_from_ncdhw_bwd(go, a, wt)
    assert torch.equal(du, du2) and torch.equal(dw, dw2) and torch.equal(db, db2)


@pytest.mark.parametrize("drop_p", [0.0, 0.05])
@pytest.mark.parametrize("co,shape", [(6, (2, 4, 6, 10)), (6, (1, 16, 16, 32)), (3, (1, 2, 4, 4))])
def test_output_head_mma_form_equals_cuda_core_form(co, shape, drop_p, monkeypatch):
    """The warp-MMA output head (fp32 weights as three bf16 terms) against the CUDA-core kernel it replaces."""
    ops = _ops()
    n, d, h, w = shape
    lazy, a = _producer(n, d, h, w, seed=12, drop_p=drop_p)
    g = torch.Generator(device=DEV).manual_seed(13)
    wt = torch.randn((co, 32, 1, 1, 1), device=DEV, generator=g) * 0.3
    b = torch.randn((co,), device=DEV, generator=g)
    got = ops.conv1x1_to_ncdhw(lazy, wt, b)
    monkeypatch.setenv("UB_HEAD_CUDA_CORES", "1")
    ref = ops.conv1x1_to_ncdhw(lazy, wt, b)
    torch.cuda.synchronize()
    assert ((got - ref).abs().max() / ref.abs().max()).item() < 2e-6


@pytest.mark.parametrize("mode", ["instance", "batch_train"])
@pytest.mark.parametrize("drop_p", [0.0, 0.05])
@pytest.mark.parametrize("c,shape", [(32, (2, 4, 8, 10)), (32, (1, 16, 16, 32)), (24, (2, 2, 4, 4))])
def test_head_backward_fused_with_norm_backward(c, shape, drop_p, mode):
    """ub_head_bwd_fused (two warp-MMA passes over dout and y, du never written) against the chain it replaces:
    ub_conv1x1_from_ncdhw_bwd -> ub_norm_act_bwd on the deferred block. Differences: dout and W enter the MMAs as bf16
    and du is not rounded to bf16 in between."""
    ops = _ops()
    from unet_bssfp_b200 import _lib
    n, d, h, w = shape
    g = torch.Generator(device=DEV).manual_seed(21)
    y = to_internal(torch.randn((n, c, d, h, w), device=DEV, generator=g) * 1.5 + 0.25, dtype=torch.float16)
    yf = y[..., :c].float()
    dims = (1, 2, 3) if mode == "instance" else (0, 1, 2, 3)
    mean_c = yf.mean(dims, keepdim=True).expand(n, 1, 1, 1, c).reshape(n, c)
    rstd_c = (yf.var(dims, unbiased=False, keepdim=True).expand(n, 1, 1, 1, c).reshape(n, c) + 1e-5).rsqrt()
    gamma = torch.rand((c,), device=DEV, generator=g) + 0.5
    beta = torch.randn((c,), device=DEV, generator=g) * 0.3
    pad = lambda t: torch.cat([t, torch.zeros((n, 32 - c), device=DEV)], 1).contiguous()
    mean, rstd = pad(mean_c), pad(rstd_c)
    scale, shift = pad(gamma * rstd_c), pad(beta - mean_c * gamma * rstd_c)
    seed, slope = 555, 0.1
    imode = {"instance": _lib.UB_NORM_INSTANCE, "batch_train": _lib.UB_NORM_BATCH_TRAIN}[mode]
    lazy = ops.DeferredAct(y, scale, shift, slope, drop_p, seed)
    wt = torch.randn((6, c, 1, 1, 1), device=DEV, generator=g) * 0.3
    go = torch.randn((n, 6, d, h, w), device=DEV, generator=g)
    du, dw, db = ops.conv1x1_from_ncdhw_bwd(go, lazy, wt)
    ref = ops.norm_act_bwd(du, None, y, imode, mean, rstd, scale, slope, drop_p, seed, c, shift=shift)
    fuse = ops.NormBwdFusion(y, scale, shift, mean, rstd, slope, drop_p, seed)
    assert ops.head_bwd_fused_ok(go, fuse)
    dy, dgamma, dbeta, dbias, dw2, db2 = ops.head_bwd_fused(go, fuse, wt, imode, c)
    torch.cuda.synchronize()
    assert rel_l2(dy.float(), ref[0].float()) < 8e-3
    if c < 32:
        assert dy[..., c:].float().abs().max().item() == 0.0
    assert rel_l2(dgamma, ref[1]) < 5e-3 and rel_l2(dbeta, ref[2]) < 5e-3
    assert dbias.abs().max().item() == 0.0
    assert rel_l2(dw2, dw) < 5e-3 and rel_l2(db2, db) < 1e-5
    # and against fp32 torch autograd through the same (materialised) activations' mask / sign decisions
    nograd = ops.head_bwd_fused(go, fuse, wt, imode, c, need_params=False, want_block_grads=False)
    assert torch.equal(nograd[0], dy) and nograd[4] is None and nograd[1] is None


@pytest.mark.parametrize("f16", [False, True])
@pytest.mark.parametrize("drop_p", [0.0, 0.05])
@pytest.mark.parametrize("c0,c1,co,shape", CONV_CASES[:4])
def test_conv_forward_on_deferred_source_vs_torch(c0, c1, co, shape, drop_p, f16):
    """Against torch, not only against the materialising kernel: conv3d(cat[lrelu(dropout(y * scale + shift)), s1]) in
    fp32 on the same y / weights (dropout mask taken from the materialised tensor: the counter hash is ours). The
    fp16-operand form (packed half arithmetic, fp16 x fp16 MMA for that chunk) must be at least as close as the bf16
    form."""
    import torch.nn.functional as F
    ops = _ops()
    n, d, h, w = shape
    spec = ops.ConvSpec(0, c0, co, c1)
    lazy, a = _producer(n, d, h, w, c=c0, seed=11, drop_p=drop_p)
    lazy.f16_operand = f16
    g = torch.Generator(device=DEV).manual_seed(12)
    s1 = to_internal(torch.randn((n, c1, d, h, w), device=DEV, generator=g)) if c1 else None
    wt = torch.randn((co, c0 + c1, 3, 3, 3), device=DEV, generator=g) / ((c0 + c1) * 27) ** 0.5
    b = torch.randn((co,), device=DEV, generator=g)
    wpk = ops.pack_conv_weights(spec, wt, ops.UB_PACK_F16_SRC0 if f16 else 0)
    got, _ = ops.conv_fwd(spec, lazy, s1, wpk, b, want_stats=True)
    # fp32 truth from y
    yf = lazy.y[..., :c0].float().permute(0, 4, 1, 2, 3)
    z = yf * lazy.scale[:, :c0, None, None, None] + lazy.shift[:, :c0, None, None, None]
    act = F.leaky_relu(z, 0.1) / (1.0 - drop_p)
    keep = (a[..., :c0].float().permute(0, 4, 1, 2, 3) != 0) | (act == 0)
    act = act * keep
    src = act if not c1 else torch.cat([act, s1[..., :c1].float().permute(0, 4, 1, 2, 3)], 1)
    ref = F.conv3d(src, wt, b, padding=1)
    err = ((got[..., :co].float().permute(0, 4, 1, 2, 3) - ref).abs().max() / ref.abs().max()).item()
    ref_b, _ = ops.conv_fwd(spec, a, s1, ops.pack_conv_weights(spec, wt, 0), b, want_stats=True)
    err_b = ((ref_b[..., :co].float().permute(0, 4, 1, 2, 3) - ref).abs().max() / ref.abs().max()).item()
    assert err < 6e-3, (err, err_b)
    if f16:
        assert err <= err_b + 5e-4, (err, err_b)


def test_fp16_operand_form_is_forward_only():
    ops = _ops()
    lazy, a = _producer(1, 4, 16, 8, seed=13)
    lazy.f16_operand = True
    spec = ops.ConvSpec(0, 32, 32)
    with pytest.raises(RuntimeError, match="bf16"):
        ops.conv_wgrad(spec, lazy, None, to_internal(torch.randn((1, 32, 4, 16, 8), device=DEV)), (32, 32, 3, 3, 3))
    with pytest.raises(RuntimeError, match="UB_PACK_F16_SRC0"):
        ops.pack_conv_weights(ops.ConvSpec(0, 64, 64), torch.randn((64, 64, 3, 3, 3), device=DEV), ops.UB_PACK_F16_SRC0)


@pytest.mark.parametrize("mod,shape,train", [("bssfp", (1, 32, 32, 32), False), ("t1w", (2, 32, 48, 32), True),
                                            ("bssfp", (1, 64, 64, 64), True)])
def test_generator_with_and_without_deferral(mod, shape, train, monkeypatch):
    """The whole generator. WITH a backward to come only the block in front of the fused output head stays deferred
    (a memory-bound consumer evaluating the canonical bf16 form): with the unfused head backward outputs, input
    gradient and every parameter gradient agree BIT FOR BIT with the materialising path (dropout active in train mode: the mask is a counter
    hash, identical in both runs for the same seed). WITHOUT a backward (no_grad) all five full-resolution 32-channel
    blocks stay deferred and their conv consumers take the fp16-operand form: the output agrees with the
    materialising path to well below the bf16 noise of the network."""
    import unet_bssfp_b200 as ub
    from unet_bssfp_b200 import modules
    n, d, h, w = shape
    cin = 24 if mod == "bssfp" else 6
    torch.manual_seed(0)
    g = ub.Generator(mod).to(DEV)
    g.train(train)
    torch.manual_seed(5)
    x = torch.rand(n, cin, d, h, w, device=DEV)
    dY = torch.randn(n, 6, d, h, w, device=DEV)

    def run(defer, head_fuse=False):
        monkeypatch.setattr(modules, "_DEFER", defer)
        monkeypatch.setattr(modules, "_DEFER_CONV", defer)      # the opt-in conv operand path as well
        monkeypatch.setattr(modules, "_HEAD_FUSE", head_fuse)   # the head backward fused with the block's norm backward
        g._net()._defer_plans.clear()
        plan_bwd = g._net().defer_plan(n, d, h, w, True, True)[:2]
        plan_inf = g._net().defer_plan(n, d, h, w, True, False)[:2]
        for p in g.parameters():
            p.grad = None
        xr = x.clone().requires_grad_(True)
        torch.manual_seed(77)                       # the dropout seeds come from the host RNG
        out = g(xr)
        out.backward(dY)
        with torch.no_grad():
            torch.manual_seed(77)
            out_inf = g(x)
        torch.cuda.synchronize()
        grads = {k: p.grad.clone() for k, p in g.named_parameters() if p.grad is not None}
        return plan_bwd, plan_inf, out.detach(), xr.grad, grads, out_inf

    pb1, pi1, out1, dx1, gr1, inf1 = run(True)
    pb0, pi0, out0, dx0, gr0, inf0 = run(False)
    convs = {"conv_0.conv_0", "conv_0.conv_1", "upcat_1.conv_0", "upcat_1.conv_1"}
    # (the fp16 operand copies in front of the marching convs are a separate switch, on in both runs)
    assert pb0 == (frozenset(), convs) and pi0 == (frozenset(), convs)
    assert pb1 == (frozenset({"upcat_1.conv_1"}), convs)
    assert pi1[0] == {"head", "conv_0.conv_0", "conv_0.conv_1", "upcat_1.conv_0", "upcat_1.conv_1"}
    assert pi1[1] == convs
    assert torch.equal(out1, out0)
    assert torch.equal(dx1, dx0)
    assert gr1.keys() == gr0.keys() and len(gr1) > 80
    for k in gr1:
        assert torch.equal(gr1[k], gr0[k]), k
    assert torch.equal(inf0, out0)                               # materialising path: grad mode does not matter
    assert rel_l2(inf1, inf0) < 1.5e-2, rel_l2(inf1, inf0)       # fp16-operand form vs bf16 form of five blocks: bf16 noise
    # The default backward of the deferred head (ub_head_bwd_fused: two warp-MMA passes, du never written, dout and W
    # as bf16 MMA operands) is the same mathematics with different roundings: same forward bits, gradients within the
    # bf16 noise of the chain it replaces.
    _, _, out2, dx2, gr2, _ = run(True, head_fuse=True)
    assert torch.equal(out2, out0)
    assert rel_l2(dx2, dx0) < 3e-2, rel_l2(dx2, dx0)
    worst = max((rel_l2(gr2[k], gr0[k]), k) for k in gr0 if gr0[k].norm() > 0)
    assert worst[0] < 6e-2, worst
    tot = torch.cat([(gr2[k] - gr0[k]).flatten() for k in gr0]).norm() / torch.cat([gr0[k].flatten() for k in gr0]).norm()
    assert tot.item() < 2e-2, tot.item()


@pytest.mark.parametrize("drop_p", [0.0, 0.05])
@pytest.mark.parametrize("c0,c1,co,shape", CONV_CASES[:4])
def test_conv_forward_on_materialised_fp16_copy(c0, c1, co, shape, drop_p):
    """The default accuracy path of the full-resolution layers: norm_act_fwd writes an fp16 copy of the activations
    beside the bf16 tensor (same fp32 values, two roundings), the marching forward conv multiplies it against
    fp16-packed weight columns. Checked against torch in fp32 on the same y / weights; it must be closer than the
    bf16 path, and the bf16 tensor written alongside must be the canonical one."""
    import torch.nn.functional as F
    ops = _ops()
    n, d, h, w = shape
    spec = ops.ConvSpec(0, c0, co, c1)
    lazy, a = _producer(n, d, h, w, c=c0, seed=21, drop_p=drop_p)
    a_bf, _, a16 = ops.norm_act_fwd(lazy.y, lazy.scale, lazy.shift, lazy.slope, drop_p, lazy.drop_seed, f16_copy=True)
    none, _, a16_only = ops.norm_act_fwd(lazy.y, lazy.scale, lazy.shift, lazy.slope, drop_p, lazy.drop_seed,
                                         f16_copy=True, materialize=False)
    assert torch.equal(a_bf, a) and none is None and torch.equal(a16_only, a16) and a16.dtype == torch.float16
    assert torch.equal((a16 == 0), (a == 0)) or drop_p == 0.0            # the same dropout mask
    assert (a16.float() - a.float()).abs().max().item() <= 2.0 ** -8 * a.float().abs().max().item()
    g = torch.Generator(device=DEV).manual_seed(22)
    s1 = to_internal(torch.randn((n, c1, d, h, w), device=DEV, generator=g)) if c1 else None
    wt = torch.randn((co, c0 + c1, 3, 3, 3), device=DEV, generator=g) / ((c0 + c1) * 27) ** 0.5
    b = torch.randn((co,), device=DEV, generator=g)
    got, st16 = ops.conv_fwd(spec, a16, s1, ops.pack_conv_weights(spec, wt, ops.UB_PACK_F16_SRC0), b, want_stats=True)
    ref_b, _ = ops.conv_fwd(spec, a, s1, ops.pack_conv_weights(spec, wt, 0), b, want_stats=True)
    yf = lazy.y[..., :c0].float().permute(0, 4, 1, 2, 3)
    z = yf * lazy.scale[:, :c0, None, None, None] + lazy.shift[:, :c0, None, None, None]
    act = F.leaky_relu(z, 0.1) / (1.0 - drop_p) * (a[..., :c0].float().permute(0, 4, 1, 2, 3) != 0)
    src = act if not c1 else torch.cat([act, s1[..., :c1].float().permute(0, 4, 1, 2, 3)], 1)
    ref = F.conv3d(src, wt, b, padding=1)
    nchw = lambda t: t[..., :co].float().permute(0, 4, 1, 2, 3)
    e16, eb = rel_l2(nchw(got), ref), rel_l2(nchw(ref_b), ref)
    assert e16 < 1.5e-3 and eb < 6e-3, (e16, eb)
    if c1 == 0:
        assert e16 < 0.5 * eb, (e16, eb)          # every operand in fp16: well below the bf16 path's error
    else:
        assert e16 < eb, (e16, eb)                # the bf16 source 1 still contributes its share
    # a generic-kernel conv refuses fp16 sources
    with pytest.raises(RuntimeError, match="not supported|bfloat16"):
        ops.conv_fwd(ops.ConvSpec(0, 32, 64), a16, None, ops.pack_conv_weights(ops.ConvSpec(0, 32, 64),
                     torch.randn((64, 32, 3, 3, 3), device=DEV), 0), None)


def test_inference_path_defers_too():
    """``forward_packed`` (sliding-window inference) takes the same deferred path and agrees with ``forward``."""
    import unet_bssfp_b200 as ub
    torch.manual_seed(0)
    g = ub.Generator("bssfp").to(DEV).eval()
    x = torch.rand(2, 24, 32, 32, 32, device=DEV)
    with torch.no_grad():
        ref = g(x)
        got = g.forward_packed(ub.ops.pack_ncdhw(x))
    assert torch.equal(got, ref)
