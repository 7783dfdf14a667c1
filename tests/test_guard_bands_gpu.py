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
