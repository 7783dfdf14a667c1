"""Guard bands around every output of the NEW C ABI entry points (fp32 mode, AdamW, NIfTI order, BatchNorm running
update, statistics / norm-backward finalisation) on ragged shapes: outputs live in the middle of NaN-filled arenas
and the bands must stay untouched, every output element must be written (compute-sanitizer is closed on the pool)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 2048


def arena(numel, dtype=torch.float32):
    buf = torch.full((numel + 2 * GUARD,), float("nan"), dtype=dtype, device="cuda")
    return buf, buf[GUARD:GUARD + numel]


def intact(*bufs):
    return all(bool(torch.isnan(b[:GUARD]).all() and torch.isnan(b[-GUARD:]).all()) for b in bufs)


def written(*views):
    return all(not bool(torch.isnan(v.float()).any()) for v in views)


P = lambda t: None if t is None else C.c_void_p(t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("k,s,p,c0,c1,co,shape", [(3, 1, 1, 5, 3, 7, (2, 5, 7, 9)), (1, 1, 0, 6, 0, 24, (1, 3, 5, 7)),
                                                  (4, 2, 1, 3, 2, 9, (2, 6, 10, 14))])
def test_fp32_conv_entries(k, s, p, c0, c1, co, shape):
    from unet_bssfp_b200 import _lib
    lib = _lib.load()
    n, d, h, w = shape
    od, oh, ow = [(v + 2 * p - k) // s + 1 for v in (d, h, w)]
    desc = _lib.F32ConvDesc(n, c0, c1, co, d, h, w, k, s, p)
    torch.manual_seed(0)
    s0 = torch.randn(n, c0, d, h, w, device="cuda")
    s1 = torch.randn(n, c1, d, h, w, device="cuda") if c1 else None
    wt = torch.randn(co, c0 + c1, k, k, k, device="cuda")
    b = torch.randn(co, device="cuda")
    ybuf, y = arena(n * co * od * oh * ow)
    _lib.check(lib.ub_f32_conv_fwd(C.byref(desc), P(s0), P(s1), P(wt), P(b), P(y), _st()))
    dy = torch.randn(n, co, od, oh, ow, device="cuda")
    d0buf, d0 = arena(s0.numel())
    d1buf, d1 = arena(max(1, s1.numel() if c1 else 1))
    _lib.check(lib.ub_f32_conv_dgrad(C.byref(desc), P(dy), P(wt), P(d0), P(d1) if c1 else None, _st()))
    dwbuf, dw = arena(wt.numel())
    dbbuf, db = arena(co)
    _lib.check(lib.ub_f32_conv_wgrad(C.byref(desc), P(s0), P(s1), P(dy), P(dw), P(db), _st()))
    torch.cuda.synchronize()
    assert intact(ybuf, d0buf, d1buf, dwbuf, dbbuf)
    assert written(y, d0, dw, db) and (not c1 or written(d1))


def test_fp32_deconv_and_norm_entries():
    from unet_bssfp_b200 import _lib
    lib = _lib.load()
    n, ci, co, d, h, w = 2, 5, 3, 3, 5, 7
    torch.manual_seed(0)
    x = torch.randn(n, ci, d, h, w, device="cuda")
    wt = torch.randn(ci, co, 2, 2, 2, device="cuda")
    b = torch.randn(co, device="cuda")
    ybuf, y = arena(n * co * 8 * d * h * w)
    _lib.check(lib.ub_f32_deconv2_fwd(n, ci, co, d, h, w, P(x), P(wt), P(b), P(y), _st()))
    dy = torch.randn(n, co, 2 * d, 2 * h, 2 * w, device="cuda")
    dxbuf, dx = arena(x.numel())
    _lib.check(lib.ub_f32_deconv2_dgrad(n, ci, co, d, h, w, P(dy), P(wt), P(dx), _st()))
    dwbuf, dw = arena(wt.numel())
    dbbuf, db = arena(co)
    _lib.check(lib.ub_f32_deconv2_wgrad(n, ci, co, d, h, w, P(x), P(dy), P(dw), P(db), _st()))
    # norm statistics + activation + pool, forward and backward, on the odd-sized (6, 10, 14) output
    D, H, W = 2 * d, 2 * h, 2 * w
    vol = D * H * W
    gam, bet = torch.rand(co, device="cuda") + 0.5, torch.randn(co, device="cuda")
    outs = [arena(n * co) for _ in range(4)]
    for mode, rm, rv in ((0, None, None), (1, torch.zeros(co, device="cuda"), torch.ones(co, device="cuda"))):
        _lib.check(lib.ub_f32_norm_stats(P(y), n, co, vol, mode, P(gam), P(bet), 1e-5, 0.1, P(rm), P(rv),
                                         *[P(v) for _, v in outs], _st()))
    (sb, scale), (hb, shift), (mb, mean), (rb, rstd) = outs
    abuf, a = arena(y.numel())
    pbuf, pooled = arena(n * co * (D // 2) * (H // 2) * (W // 2))
    _lib.check(lib.ub_f32_norm_act_fwd(P(y), P(scale), P(shift), 0.1, 0.05, 77, n, co, D, H, W, P(a), P(pooled), _st()))
    dA = torch.randn(n, co, D, H, W, device="cuda")
    dP = torch.randn(n, co, D // 2, H // 2, W // 2, device="cuda")
    c12buf, c12 = arena(2 * n * co)
    gbuf, g = arena(y.numel())
    dgbuf, dg = arena(co)
    dbb, dbe = arena(co)
    _lib.check(lib.ub_f32_norm_act_bwd(P(dA), P(dP), P(a), P(y), 1, P(mean), P(rstd), P(scale), P(shift), 0.1, 0.05, 77,
                                       n, co, D, H, W, P(c12), P(g), P(dg), P(dbe), _st()))
    torch.cuda.synchronize()
    assert intact(ybuf, dxbuf, dwbuf, dbbuf, sb, hb, mb, rb, abuf, pbuf, c12buf, gbuf, dgbuf, dbb)
    assert written(y, dx, dw, db, scale, shift, mean, rstd, a, pooled, c12, g, dg, dbe)


def test_adamw_nifti_bn_update_entries():
    from unet_bssfp_b200 import _lib
    lib = _lib.load()
    sizes = [1, 3, 4097, 8191, 12289] + [50] * 60            # > 48 tensors: two launches; odd tails
    bufs, table = [], (_lib.AdamWTensor * len(sizes))()
    torch.manual_seed(0)
    for i, nel in enumerate(sizes):
        quad = [arena(nel) for _ in range(3)]
        for _, v in quad:
            v.copy_(torch.rand(nel, device="cuda"))
        g = torch.randn(nel, device="cuda")
        bufs.append((quad, g))
        table[i] = _lib.AdamWTensor(quad[0][1].data_ptr(), g.data_ptr(), quad[1][1].data_ptr(), quad[2][1].data_ptr(), nel)
    _lib.check(lib.ub_adamw_step(table, len(sizes), 1e-3, 0.9, 0.999, 1e-8, 1e-2, 3, 1.0, _st()))
    torch.cuda.synchronize()
    for quad, _ in bufs:
        assert intact(*[b for b, _ in quad]) and written(*[v for _, v in quad])
    # NIfTI order with sizes that are not multiples of the 32 x 32 tile
    vol = torch.rand(3, 33, 5, 70, device="cuda")
    obuf, out = arena(vol.numel())
    _lib.check(lib.ub_denorm_to_nifti(P(vol), 3, 33, 5, 70, 2.0, 0.5, P(out), _st()))
    # BatchNorm running update
    c = 37
    rmb, rm = arena(c)
    rvb, rv = arena(c)
    rm.zero_(); rv.fill_(1.0)
    _lib.check(lib.ub_bn_running_update(P(torch.randn(64, device="cuda")), P(torch.rand(64, device="cuda") + 0.5), c, 1000.0,
                                        1e-5, 0.1, P(rm), P(rv), _st()))
    torch.cuda.synchronize()
    assert intact(obuf, rmb, rvb) and written(out, rm, rv)


@pytest.mark.parametrize("n,tiles,cp,c,mode", [(3, 700, 32, 24, 1), (2, 1500, 64, 40, 1), (3, 37, 32, 24, 0), (1, 5, 96, 70, 1)])
def test_norm_finalize_entries(n, tiles, cp, c, mode):
    """Both BatchNorm statistics paths (single block, two-stage with scratch in the output slots) and InstanceNorm."""
    from unet_bssfp_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(0)
    part = torch.rand(n * tiles, 2, cp, device="cuda")
    gam, bet = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    outs = [arena(n * cp) for _ in range(4)]
    _lib.check(lib.ub_norm_finalize(P(part), tiles, n, cp, c, 4096.0, P(gam), P(bet), 1e-5, mode, 0.1, P(rm), P(rv),
                                    *[P(v) for _, v in outs], _st()))
    torch.cuda.synchronize()
    assert intact(*[b for b, _ in outs]) and written(*[v for _, v in outs])
    # values: mean = sum / count over the reduced axes
    s = part[:, 0, :c].double().view(n, tiles, c).sum(1)
    q = part[:, 1, :c].double().view(n, tiles, c).sum(1)
    cnt = 4096.0 * (n if mode == 1 else 1)
    mean = (s.sum(0, keepdim=True).expand(n, c) if mode == 1 else s) / cnt
    msq = (q.sum(0, keepdim=True).expand(n, c) if mode == 1 else q) / cnt
    var = (msq - mean * mean).clamp_min(0)
    got_mean = outs[2][1].view(n, cp)[:, :c].double()
    got_rstd = outs[3][1].view(n, cp)[:, :c].double()
    torch.testing.assert_close(got_mean, mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got_rstd, 1.0 / torch.sqrt(var + 1e-5), rtol=1e-4, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------
# round 2: the warp-MMA output head (forward, fused backward), the space-to-depth copy, the fused pooling backward,
# the instruction-lean norm-backward apply -- ragged / minimal shapes, every output inside a NaN arena
# ---------------------------------------------------------------------------------------------------------------
def _norm_block(n, d, h, w, c, seed=0):
    """y (fp16, 32 padded channels, c real) and its per-(n, c) constants."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    y = torch.zeros((n, d, h, w, 32), dtype=torch.float16, device="cuda")
    y[..., :c] = (torch.randn((n, d, h, w, c), device="cuda", generator=g) * 1.5 + 0.2).half()
    pad = lambda t: torch.cat([t, torch.zeros((n, 32 - c), device="cuda")], 1).contiguous()
    scale, shift = pad(torch.rand((n, c), device="cuda", generator=g) + 0.5), pad(torch.randn((n, c), device="cuda", generator=g) * 0.2)
    mean, rstd = pad(torch.randn((n, c), device="cuda", generator=g) * 0.1), pad(torch.rand((n, c), device="cuda", generator=g) + 0.5)
    return y, scale, shift, mean, rstd


@pytest.mark.parametrize("co,c,shape,drop_p", [(6, 32, (2, 2, 4, 6), 0.05), (3, 24, (1, 1, 4, 4), 0.0), (8, 32, (1, 3, 2, 8), 0.0)])
def test_output_head_warp_mma_entries(co, c, shape, drop_p):
    from unet_bssfp_b200 import _lib
    lib = _lib.load()
    n, d, h, w = shape
    vox = d * h * w
    assert vox % 16 == 0
    y, scale, shift, mean, rstd = _norm_block(n, d, h, w, c)
    torch.manual_seed(1)
    wt = torch.randn(co, c, device="cuda") * 0.3
    bias = torch.randn(co, device="cuda")
    act = _lib.DeferredAct(scale.data_ptr(), shift.data_ptr(), 0.1, drop_p, 77, 0)
    ws = torch.empty(lib.ub_conv1x1_workspace_bytes() // 4, device="cuda")
    obuf, out = arena(n * co * vox)
    _lib.check(lib.ub_conv1x1_to_ncdhw(P(y), 32, P(wt), c, P(bias), co, n, vox, P(ws), P(out), C.byref(act), _st()))
    dout = torch.randn(n, co, vox, device="cuda")
    fuse = _lib.NormBwdFuse(y.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), 0.1,
                            drop_p, 77, None)
    ws2 = torch.empty(lib.ub_head_bwd_fused_workspace_bytes(n) // 4, device="cuda")
    dybuf, dy = arena(n * vox * 32, torch.bfloat16)
    dgbuf, dg = arena(c); dbtbuf, dbt = arena(c); dbsbuf, dbs = arena(c)
    dwbuf, dw = arena(co * c); dbbuf, db = arena(co)
    _lib.check(lib.ub_head_bwd_fused(P(dout), co, P(wt), c, n, vox, C.byref(fuse), _lib.UB_NORM_INSTANCE, c, P(ws2), P(dy),
                                     P(dg), P(dbt), P(dbs), P(dw), P(db), _st()))
    torch.cuda.synchronize()
    assert intact(obuf, dybuf, dgbuf, dbtbuf, dbsbuf, dwbuf, dbbuf)
    assert written(out, dy, dg, dbt, dbs, dw, db)


@pytest.mark.parametrize("cp,shape", [(32, (2, 2, 4, 6)), (64, (1, 4, 2, 2)), (512, (1, 2, 2, 2))])
def test_space_to_depth_copy_and_fused_pool_backward_entries(cp, shape):
    from unet_bssfp_b200 import _lib
    lib = _lib.load()
    n, d, h, w = shape
    torch.manual_seed(2)
    src = torch.randn(n, d, h, w, cp, device="cuda").bfloat16()
    obuf, out = arena(src.numel(), torch.bfloat16)
    _lib.check(lib.ub_to_s2d(P(src), n, d, h, w, cp, P(out), _st()))
    # fused pooling backward: dA (accumulated in place) and the partial records
    g = torch.Generator(device="cuda").manual_seed(3)
    y = (torch.randn((n, d, h, w, cp), device="cuda", generator=g)).half()
    scale, shift = torch.rand((n, cp), device="cuda") + 0.5, torch.randn((n, cp), device="cuda") * 0.2
    mean, rstd = torch.randn((n, cp), device="cuda") * 0.1, torch.rand((n, cp), device="cuda") + 0.5
    records = lib.ub_maxpool_bwd_fuse_records(n, d, h, w, cp)
    assert records > 0
    pbuf, part = arena(records * 2 * cp)
    dabuf, dA = arena(y.numel(), torch.bfloat16)
    dP = torch.randn(n, d // 2, h // 2, w // 2, cp, device="cuda").bfloat16()
    fuse = _lib.NormBwdFuse(y.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), 0.1,
                            0.05, 5, part.data_ptr())
    _lib.check(lib.ub_maxpool_bwd_fused(P(dP), P(dA), 0, n, d, h, w, cp, C.byref(fuse), _st()))
    torch.cuda.synchronize()
    assert intact(obuf, pbuf, dabuf)
    assert written(out, part, dA)


@pytest.mark.parametrize("c,shape,drop_p", [(24, (1, 3, 5, 7), 0.05), (512, (2, 1, 1, 3), 0.0), (64, (1, 9, 17, 5), 0.05)])
def test_lean_norm_backward_apply_entry(c, shape, drop_p):
    from unet_bssfp_b200 import _lib
    lib = _lib.load()
    n, d, h, w = shape
    cp = (c + 31) // 32 * 32
    vox = d * h * w
    torch.manual_seed(4)
    y = torch.randn(n, d, h, w, cp, device="cuda").half()
    dA = torch.randn(n, d, h, w, cp, device="cuda").bfloat16()
    scale, shift = torch.rand((n, cp), device="cuda") + 0.5, torch.randn((n, cp), device="cuda") * 0.2
    mean, rstd = torch.randn((n, cp), device="cuda") * 0.1, torch.rand((n, cp), device="cuda") + 0.5
    ws = torch.empty(lib.ub_norm_act_bwd_workspace_bytes(n, cp) // 4, device="cuda")
    dybuf, dy = arena(y.numel(), torch.bfloat16)
    dgbuf, dg = arena(c); dbbuf, db = arena(c); dsbuf, ds = arena(c)
    _lib.check(lib.ub_norm_act_bwd(P(dA), None, P(y), _lib.UB_NORM_INSTANCE, P(mean), P(rstd), P(scale), P(shift), 0.1, drop_p,
                                   9, n, vox, cp, c, P(ws), P(dy), P(dg), P(db), P(ds), None, 0, _st()))
    torch.cuda.synchronize()
    assert intact(dybuf, dgbuf, dbbuf, dsbuf)
    assert written(dy, dg, db, ds)
