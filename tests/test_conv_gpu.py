"""GPU parity: the tcgen05 implicit-GEMM convolutions (through the C ABI) against torch.nn.functional
on identical bf16-rounded inputs and weights (fp32 math, TF32 off). Tolerance: 1e-2 relative to the
largest magnitude (north_star bf16 bar); expected ~4e-3 from the bf16 output rounding."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import bf16_round, from_internal, rel_to_max, rel_l2, strict_fp32, to_internal

pytestmark = pytest.mark.gpu
TOL = 1e-2


def _ops():
    from unet_bssfp_b200 import ops
    return ops


def _ref_conv(kind, x, w, b):
    if kind == 0:
        return F.conv3d(x, w, b, stride=1, padding=1)
    if kind == 1:
        return F.conv3d(x, w, b)
    if kind in (2, 4):
        return F.conv3d(x, w, b, stride=2, padding=1)
    return F.conv_transpose3d(x, w, b, stride=2)


def _weight_shape(kind, ci, co):
    k = {0: 3, 1: 1, 2: 4, 3: 2, 4: 4}[kind]
    return (ci, co, k, k, k) if kind == 3 else (co, ci, k, k, k)


CASES = [
    # kind, c0, c1, co, (n, d, h, w)
    (0, 32, 0, 32, (2, 8, 16, 8)),
    (0, 24, 0, 32, (1, 6, 20, 12)),      # channel padding + ragged tiles in d, h, w
    (0, 64, 0, 64, (1, 4, 16, 16)),
    (0, 32, 0, 32, (1, 40, 16, 8)),      # marching kernel: ring wrap-around, 5 d-segments with halos
    (0, 32, 0, 64, (1, 9, 24, 16)),      # dgrad on the marching kernel with K = 64
    (0, 64, 0, 32, (2, 11, 16, 24)),     # forward on the marching kernel with two K chunks of one source
    (0, 32, 64, 32, (1, 5, 16, 8)),      # skip concat [32 | 64] as two sources, split dgrad destinations
    (0, 32, 64, 32, (2, 7, 32, 32)),     # marching kernel on CTA pairs (cta_group::2): 4 w tiles, 3 K chunks, two sources
    (0, 64, 0, 32, (1, 12, 16, 16)),     # CTA pairs, two K chunks, two d segments
    (0, 128, 0, 256, (1, 4, 8, 8)),      # several N tiles, plane smaller than the 16x8 tile
    (0, 256, 256, 256, (1, 2, 4, 4)),
    (1, 24, 0, 24, (2, 4, 16, 8)),
    (1, 32, 0, 6, (1, 4, 16, 8)),
    (1, 512, 0, 1, (2, 2, 2, 2)),
    (2, 30, 0, 32, (1, 16, 32, 16)),
    (2, 32, 0, 64, (2, 8, 8, 8)),
    (2, 256, 0, 512, (2, 4, 4, 4)),
    (4, 30, 0, 32, (1, 16, 32, 16)),     # PatchGAN stem reading the space-to-depth pack
    (4, 12, 0, 32, (2, 6, 36, 20)),      # ragged tiles
    (4, 64, 0, 64, (1, 8, 8, 8)),        # two 32-channel chunks per parity group
    (4, 32, 0, 64, (2, 16, 16, 16)),     # PatchGAN d2 on a space-to-depth copy: one chunk, N = 2 x 64 folded
    (4, 64, 0, 128, (1, 8, 16, 8)),      # d3: two chunks per group, N = 2 x 128
    (4, 128, 0, 256, (2, 8, 8, 8)),      # d4: two N tiles, depth shifts not folded
    (4, 256, 0, 512, (2, 4, 4, 4)),      # d5: planes smaller than the tile, four N tiles
    (3, 64, 0, 64, (1, 4, 16, 8)),
    (3, 512, 0, 256, (2, 2, 4, 4)),
    (3, 128, 0, 64, (1, 3, 6, 5)),
]


def _src(kind, x):
    """NCDHW fp32 -> the source tensor layout the conv kind reads."""
    t = to_internal(x)
    if kind == 4:
        n, d, h, w, cp = t.shape
        t = t.view(n, d // 2, 2, h // 2, 2, w // 2, 2, cp).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(
            n, 8, d // 2, h // 2, w // 2, cp).contiguous()
    return t


def _make(kind, c0, c1, co, shape, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    n, d, h, w = shape
    ci = c0 + c1
    x = bf16_round(torch.randn((n, ci, d, h, w), device="cuda", generator=g))
    wt = bf16_round(torch.randn(_weight_shape(kind, ci, co), device="cuda", generator=g) / (ci ** 0.5))
    b = torch.randn((co,), device="cuda", generator=g)
    return x, wt, b


@pytest.mark.parametrize("kind,c0,c1,co,shape", CASES)
def test_conv_forward_and_stats(kind, c0, c1, co, shape):
    strict_fp32()
    ops = _ops()
    x, wt, b = _make(kind, c0, c1, co, shape)
    spec = ops.ConvSpec(kind, c0, co, c1)
    wpk = ops.pack_conv_weights(spec, wt, 0)
    s0 = _src(kind, x[:, :c0])
    s1 = to_internal(x[:, c0:]) if c1 else None
    want_stats = kind != 3
    y, stats = ops.conv_fwd(spec, s0, s1, wpk, b, want_stats=want_stats)
    torch.cuda.synchronize()
    ref = _ref_conv(kind, x, wt, b)
    got = from_internal(y, co)
    assert got.shape == ref.shape
    assert rel_to_max(got, ref) < TOL
    # padded output channels are exactly zero
    if spec.cop > co:
        assert y[..., co:].abs().max().item() == 0.0
    if want_stats:
        n = shape[0]
        tiles = stats.shape[0] // n
        s = stats.view(n, tiles, 2, spec.cop).sum(1)
        ref_s = got.sum(dim=(2, 3, 4))
        ref_q = (got * got).sum(dim=(2, 3, 4))
        assert rel_to_max(s[:, 0, :co], ref_s) < 1e-3
        assert rel_to_max(s[:, 1, :co], ref_q) < 1e-3


@pytest.mark.parametrize("kind,c0,c1,co,shape", CASES)
def test_conv_dgrad(kind, c0, c1, co, shape):
    strict_fp32()
    ops = _ops()
    x, wt, b = _make(kind, c0, c1, co, shape, seed=1)
    spec = ops.ConvSpec(kind, c0, co, c1)
    x.requires_grad_(True)
    ref_y = _ref_conv(kind, x, wt, None)
    g = torch.Generator(device="cuda").manual_seed(7)
    dy = bf16_round(torch.randn(ref_y.shape, device="cuda", generator=g))
    (ref_dx,) = torch.autograd.grad(ref_y, x, dy)
    wpk = ops.pack_conv_weights(spec, wt, 1)
    d0, d1 = ops.conv_dgrad(spec, to_internal(dy), wpk, shape[1:])
    torch.cuda.synchronize()
    got = from_internal(d0, c0)
    if c1:
        got = torch.cat([got, from_internal(d1, c1)], dim=1)
    assert rel_to_max(got, ref_dx) < TOL


@pytest.mark.parametrize("kind,c0,c1,co,shape", CASES)
def test_conv_wgrad(kind, c0, c1, co, shape):
    strict_fp32()
    ops = _ops()
    x, wt, b = _make(kind, c0, c1, co, shape, seed=2)
    spec = ops.ConvSpec(kind, c0, co, c1)
    wt.requires_grad_(True)
    ref_y = _ref_conv(kind, x, wt, None)
    g = torch.Generator(device="cuda").manual_seed(9)
    dy = bf16_round(torch.randn(ref_y.shape, device="cuda", generator=g))
    (ref_dw,) = torch.autograd.grad(ref_y, wt, dy)
    s0 = _src(kind, x[:, :c0])
    s1 = to_internal(x[:, c0:]) if c1 else None
    dw = ops.conv_wgrad(spec, s0, s1, to_internal(dy), tuple(wt.shape))
    torch.cuda.synchronize()
    assert dw.shape == ref_dw.shape
    assert rel_to_max(dw, ref_dw) < 2e-3      # fp32 accumulate of exact bf16 products
    assert rel_l2(dw, ref_dw) < 1e-3


@pytest.mark.parametrize("mode,drop_p", [("instance", 0.05), ("instance", 0.0), ("batch_train", 0.0)])
@pytest.mark.parametrize("co,shape", [(32, (2, 9, 24, 16)), (64, (1, 6, 16, 8))])
def test_dgrad_fused_norm_backward_reduction(mode, drop_p, co, shape):
    """ub_conv_dgrad_fused: the marching dgrad epilogue accumulates the producer block's norm-backward
    sums. The gradient it writes is bit-identical to the unfused call, and ub_norm_act_bwd fed with the
    partial records agrees with its own reduction pass (different summation order only)."""
    strict_fp32()
    ops = _ops()
    from unet_bssfp_b200 import _lib
    n, d, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(5)
    spec = ops.ConvSpec(0, 32, co)
    wt = bf16_round(torch.randn((co, 32, 3, 3, 3), device="cuda", generator=g) / 30.0)
    wpk = ops.pack_conv_weights(spec, wt, 1)
    dy = to_internal(bf16_round(torch.randn((n, co, d, h, w), device="cuda", generator=g)))
    # the producer block: raw output y (24 real channels: padded channels must stay inert), its statistics
    c = 24
    y = to_internal(bf16_round(torch.randn((n, c, d, h, w), device="cuda", generator=g) * 1.3 + 0.2))
    gamma = torch.rand((c,), device="cuda", generator=g) + 0.5
    beta = torch.randn((c,), device="cuda", generator=g) * 0.2
    kspec = ops.ConvSpec(_lib.UB_CONV_K1, c, c)
    eye = torch.eye(c, device="cuda").view(c, c, 1, 1, 1)
    yi, stats = ops.conv_fwd(kspec, y, None, ops.pack_conv_weights(kspec, eye, 0), None, want_stats=True)
    imode = {"instance": _lib.UB_NORM_INSTANCE, "batch_train": _lib.UB_NORM_BATCH_TRAIN}[mode]
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    scale, shift, mean, rstd = ops.norm_finalize(stats, n, d * h * w, 32, c, gamma, beta, 1e-5, imode, 0.1, rm, rv)
    seed = 4242
    fuse = ops.NormBwdFusion(yi, scale, shift, mean, rstd, 0.1, drop_p, seed)
    assert ops.dgrad_fuse_records(spec, n, d, h, w) > 0
    d0, _ = ops.conv_dgrad(spec, dy, wpk, (d, h, w))
    f0, _, partial = ops.conv_dgrad(spec, dy, wpk, (d, h, w), fuse=fuse)
    torch.cuda.synchronize()
    assert torch.equal(d0, f0)
    ref = ops.norm_act_bwd(d0, None, yi, imode, mean, rstd, scale, 0.1, drop_p, seed, c, shift=shift)
    got = ops.norm_act_bwd(d0, None, yi, imode, mean, rstd, scale, 0.1, drop_p, seed, c, shift=shift, partial=partial)
    torch.cuda.synchronize()
    assert rel_l2(got[1], ref[1]) < 1e-4 and rel_l2(got[2], ref[2]) < 1e-4          # dgamma, dbeta
    assert rel_to_max(got[0].float(), ref[0].float()) < 1e-2                          # dy (bf16 roundings of near-ties)
    assert rel_l2(got[0].float(), ref[0].float()) < 1e-3
    # a conv whose source has 64 channels has no fused path
    assert ops.dgrad_fuse_records(ops.ConvSpec(0, 64, 64), n, d, h, w) == 0


@pytest.mark.parametrize("kind,c0,c1,co,shape", [
    (0, 24, 0, 32, (1, 6, 20, 12)),      # marching kernel, ragged h / w tiles (lane-pair stores at the border)
    (0, 64, 0, 64, (1, 5, 18, 10)),      # generic folded kernel, ragged d / h / w
    (0, 32, 64, 32, (1, 5, 20, 12)),     # two sources; dgrad with split destinations
    (3, 64, 0, 64, (1, 3, 6, 5)),        # transposed conv: scatter stores with stride 2
    (2, 32, 0, 64, (1, 6, 10, 12)),      # stride-2 conv: parity-class scatter in dgrad
    (4, 30, 0, 32, (1, 6, 36, 20)),      # stem on the space-to-depth source
])
def test_no_out_of_bounds_writes(kind, c0, c1, co, shape):
    """Guard bands around every output of the C ABI calls stay untouched (compute-sanitizer is not
    available on the pool): outputs live in the middle of NaN-filled arenas."""
    import ctypes as C
    from unet_bssfp_b200 import _lib
    ops = _ops()
    lib = _lib.load()
    x, wt, b = _make(kind, c0, c1, co, shape, seed=3)
    spec = ops.ConvSpec(kind, c0, co, c1)
    n, d, h, w = shape
    od, oh, ow = spec.out_dims(d, h, w)
    desc = spec.desc(n, d, h, w)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    GUARD = 4096

    def arena(numel, dtype):
        buf = torch.full((numel + 2 * GUARD,), float("nan"), dtype=dtype, device="cuda")
        return buf, buf[GUARD:GUARD + numel]

    def intact(buf):
        return bool(torch.isnan(buf[:GUARD]).all() and torch.isnan(buf[-GUARD:]).all())

    s0 = _src(kind, x[:, :c0])
    s1 = to_internal(x[:, c0:]) if c1 else None
    # forward (+ statistics)
    wf = ops.pack_conv_weights(spec, wt, 0)
    want_stats = kind != 3
    ybuf, y = arena(n * od * oh * ow * spec.cop, torch.float16 if want_stats else torch.bfloat16)   # with statistics: fp16 y
    tiles = lib.ub_conv_num_tiles(C.byref(desc))
    sbuf, stats = arena(tiles * 2 * spec.cop, torch.float32)
    _lib.check(lib.ub_conv_fwd(C.byref(desc), P(s0), P(s1), P(wf), P(b), 0, 0.0, P(y), P(stats) if want_stats else None,
                               None, 0, st))
    torch.cuda.synchronize()
    assert intact(ybuf) and intact(sbuf)
    assert not torch.isnan(y.float()).any()
    if want_stats:
        assert not torch.isnan(stats).any()          # every tile wrote its record
    # dgrad
    wd = ops.pack_conv_weights(spec, wt, 1)
    dy = torch.randn((n, od, oh, ow, spec.cop), device="cuda").to(torch.bfloat16)
    d0buf, d0 = arena(n * d * h * w * spec.c0p, torch.bfloat16)
    d1buf, d1 = arena(n * d * h * w * max(spec.c1p, 1), torch.bfloat16)
    _lib.check(lib.ub_conv_dgrad(C.byref(desc), P(dy), P(wd), P(d0), P(d1) if c1 else None, st))
    torch.cuda.synchronize()
    assert intact(d0buf) and intact(d1buf) and not torch.isnan(d0.float()).any()
    if c1:
        assert not torch.isnan(d1.float()).any()
    # wgrad
    nbytes = lib.ub_conv_wgrad_workspace_bytes(C.byref(desc))
    wsbuf, ws = arena(nbytes // 4, torch.float32)
    dwbuf, dw = arena(wt.numel(), torch.float32)
    dw.zero_()
    _lib.check(lib.ub_conv_wgrad(C.byref(desc), P(s0), P(s1), P(dy), P(ws), P(dw), None, st))
    torch.cuda.synchronize()
    assert intact(wsbuf) and intact(dwbuf) and not torch.isnan(dw).any()


def test_multi_tensor_weight_pack_equals_single_packs():
    """ub_pack_conv_weights_multi (one launch for a network's weights, coalesced (row, column)-per-thread mapping)
    writes exactly what ub_pack_conv_weights writes, for every conv kind, both directions and the fp16-column form."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(3)
    items = []
    for kind, c0, c1, co in [(0, 24, 0, 32), (0, 32, 64, 32), (0, 64, 64, 64), (0, 128, 0, 256), (1, 24, 0, 24), (1, 512, 0, 1),
                             (2, 32, 0, 64), (3, 128, 0, 64), (4, 30, 0, 32), (0, 32, 0, 32)]:
        spec = ops.ConvSpec(kind, c0, co, c1)
        wt = torch.randn(_weight_shape(kind, c0 + c1, co), device="cuda", generator=g)
        for direction in (0, 1):
            items.append((spec, wt, direction))
    for spec, c in ((ops.ConvSpec(0, 32, 32, 64), 96), (ops.ConvSpec(0, 24, 32), 24)):
        items.append((spec, torch.randn((32, c, 3, 3, 3), device="cuda", generator=g), ops.UB_PACK_F16_SRC0))
    multi = ops.pack_conv_weights_multi(items)
    torch.cuda.synchronize()
    assert len(multi) == len(items) > 16
    for (spec, wt, direction), got in zip(items, multi):
        assert torch.equal(got.view(torch.int16), ops.pack_conv_weights(spec, wt, direction).view(torch.int16)), (spec, direction)
