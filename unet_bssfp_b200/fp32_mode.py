"""fp32 mode of the drop-in modules (north_star: outputs and gradients within 1e-5 of the reference).

``module.precision = "fp32"`` (or ``unet_bssfp_b200.set_precision(module, "fp32")``, or ``UB_PRECISION=fp32``)
routes ``Generator`` / ``Discriminator`` / ``DownSampleConv`` / ``BasicUNet`` through the ``ub_f32_*`` entry
points of ``libubssfp.so``: the same operator graph on the CUDA cores (fp32 products, fp64 sums) with tensors in
the reference's NCDHW fp32 layout. It is the verification path -- bf16 cannot express 1e-5 -- and is meant for
small volumes; the throughput path is the bf16 tcgen05 one. There is still no CPU and no cuDNN fallback.

One ``torch.autograd.Function`` per network block: conv(cat[src0, src1]) -> norm -> dropout -> LeakyReLU
(-> MaxPool3d(2)) with outputs ``(a, pooled)``; its backward receives both gradients, so the skip / pool fan-out
of the U-Net is resolved inside the kernels and autograd never has to add tensors.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (F32ConvDesc, UB_CONV_K1, UB_CONV_K3S1P1, UB_CONV_K4S2P1, UB_CONV_K4S2P1_S2D, UB_DECONV_K2S2,
                   UB_NORM_BATCH_EVAL, UB_NORM_BATCH_TRAIN, UB_NORM_INSTANCE, UB_NORM_NONE)

_KSP = {UB_CONV_K3S1P1: (3, 1, 1), UB_CONV_K1: (1, 1, 0), UB_CONV_K4S2P1: (4, 2, 1), UB_CONV_K4S2P1_S2D: (4, 2, 1)}


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("unet_bssfp_b200 kernels run on CUDA tensors only (there is no CPU fallback)")
    return t.detach().contiguous().float()


class _Cfg:
    """Static + per-call description of one block (not a tensor: passed through autograd untouched)."""
    __slots__ = ("kind", "mode", "slope", "act", "drop_p", "seed", "pool", "eps", "momentum", "running_mean",
                 "running_var")


class _BlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg: _Cfg, src0, src1, weight, bias, gamma, beta):
        s0, s1 = _f32(src0), _f32(src1)
        w, b = _f32(weight), _f32(bias)
        lib = _lib.load()
        st = _stream()
        n, c0, d, h, wd = s0.shape
        c1 = 0 if s1 is None else s1.shape[1]
        if cfg.kind == UB_DECONV_K2S2:
            co = w.shape[1]
            y = torch.empty((n, co, 2 * d, 2 * h, 2 * wd), dtype=torch.float32, device=s0.device)
            _lib.check(lib.ub_f32_deconv2_fwd(n, c0, co, d, h, wd, _p(s0), _p(w), _p(b), _p(y), st), "ub_f32_deconv2_fwd")
            desc = None
        else:
            k, s, p = _KSP[cfg.kind]
            co = w.shape[0]
            desc = F32ConvDesc(n, c0, c1, co, d, h, wd, k, s, p)
            od, oh, ow = [(v + 2 * p - k) // s + 1 for v in (d, h, wd)]
            y = torch.empty((n, co, od, oh, ow), dtype=torch.float32, device=s0.device)
            _lib.check(lib.ub_f32_conv_fwd(C.byref(desc), _p(s0), _p(s1), _p(w), _p(b), _p(y), st), "ub_f32_conv_fwd")
        _, _, od, oh, ow = y.shape
        scale = shift = mean = rstd = None
        if cfg.mode != UB_NORM_NONE:
            scale, shift, mean, rstd = (torch.empty((n, co), dtype=torch.float32, device=y.device) for _ in range(4))
            _lib.check(lib.ub_f32_norm_stats(_p(y), n, co, od * oh * ow, cfg.mode, _p(_f32(gamma)), _p(_f32(beta)),
                                             cfg.eps, cfg.momentum, _p(cfg.running_mean), _p(cfg.running_var), _p(scale),
                                             _p(shift), _p(mean), _p(rstd), st), "ub_f32_norm_stats")
        pooled = None
        if cfg.mode != UB_NORM_NONE or cfg.act or cfg.pool:
            a = torch.empty_like(y)
            if cfg.pool:
                pooled = torch.empty((n, co, od // 2, oh // 2, ow // 2), dtype=torch.float32, device=y.device)
            slope = cfg.slope if (cfg.mode != UB_NORM_NONE or cfg.act) else 1.0
            _lib.check(lib.ub_f32_norm_act_fwd(_p(y), _p(scale), _p(shift), slope, cfg.drop_p, cfg.seed, n, co, od, oh,
                                               ow, _p(a), _p(pooled), st), "ub_f32_norm_act_fwd")
        else:
            a = y
        ctx.cfg, ctx.desc = cfg, desc
        # save_for_backward (also for the outputs): no output -> grad_fn -> ctx -> output reference cycle, autograd's
        # in-place-modification checks apply, and a second backward (retain_graph) works
        ctx.save_for_backward(s0, s1, w, y, a, scale, shift, mean, rstd)
        ctx.has_bias = bias is not None
        if pooled is None:
            return a, None
        return a, pooled

    @staticmethod
    def backward(ctx, dA, dP):
        lib = _lib.load()
        st = _stream()
        cfg, desc = ctx.cfg, ctx.desc
        s0, s1, w, y, a, scale, shift, mean, rstd = ctx.saved_tensors
        n, co, od, oh, ow = y.shape
        dA, dP = _f32(dA), _f32(dP)
        if dA is None and dP is None:
            return (None,) * 7
        dgamma = dbeta = None
        has_norm = cfg.mode != UB_NORM_NONE
        need_gamma = has_norm and (ctx.needs_input_grad[5] or ctx.needs_input_grad[6])
        if has_norm or cfg.act or dP is not None:
            dy = torch.empty_like(y)
            c1c2 = torch.empty((2, n, co), dtype=torch.float32, device=y.device) if has_norm else None
            if need_gamma:
                dgamma = torch.empty((co,), dtype=torch.float32, device=y.device)
                dbeta = torch.empty((co,), dtype=torch.float32, device=y.device)
            slope = cfg.slope if (has_norm or cfg.act) else 1.0
            _lib.check(lib.ub_f32_norm_act_bwd(_p(dA), _p(dP), _p(a), _p(y), cfg.mode, _p(mean), _p(rstd), _p(scale),
                                               _p(shift), slope, cfg.drop_p, cfg.seed, n, co, od, oh, ow, _p(c1c2), _p(dy),
                                               _p(dgamma), _p(dbeta), st), "ub_f32_norm_act_bwd")
        else:
            dy = dA
        need0, need1 = ctx.needs_input_grad[1], ctx.needs_input_grad[2] and s1 is not None
        need_w, need_b = ctx.needs_input_grad[3], ctx.has_bias and ctx.needs_input_grad[4]
        d0 = d1 = dw = db = None
        if need_w:
            dw = torch.empty_like(w)
        if need_b:
            db = torch.empty((co,), dtype=torch.float32, device=y.device)
        if cfg.kind == UB_DECONV_K2S2:
            _, ci, d, h, wd = s0.shape
            if need_w or need_b:
                _lib.check(lib.ub_f32_deconv2_wgrad(n, ci, co, d, h, wd, _p(s0), _p(dy), _p(dw), _p(db), st),
                           "ub_f32_deconv2_wgrad")
            if need0:
                d0 = torch.empty_like(s0)
                _lib.check(lib.ub_f32_deconv2_dgrad(n, ci, co, d, h, wd, _p(dy), _p(w), _p(d0), st), "ub_f32_deconv2_dgrad")
        else:
            if need_w or need_b:
                _lib.check(lib.ub_f32_conv_wgrad(C.byref(desc), _p(s0), _p(s1), _p(dy), _p(dw), _p(db), st),
                           "ub_f32_conv_wgrad")
            if need0 or need1:
                d0 = torch.empty_like(s0)
                d1 = torch.empty_like(s1) if s1 is not None else None
                _lib.check(lib.ub_f32_conv_dgrad(C.byref(desc), _p(dy), _p(w), _p(d0), _p(d1), st), "ub_f32_conv_dgrad")
                if not need0:
                    d0 = None
                if not need1:
                    d1 = None
        return None, d0, d1, dw, db, dgamma if ctx.needs_input_grad[5] else None, dbeta if ctx.needs_input_grad[6] else None


def run_block(blk, src0, src1, seed, pool=False):
    """One ``modules._Block`` in fp32 mode -> (a, pooled). Train / eval behaviour follows the block's own norm and
    dropout modules, as in the bf16 path."""
    from .modules import _bn_momentum
    cfg = _Cfg()
    cfg.kind = blk.spec.kind
    cfg.slope, cfg.act, cfg.pool = float(blk.slope), bool(blk.fused_act), bool(pool)
    cfg.drop_p = blk.dropout_p()
    cfg.seed = int(seed) & 0x7FFFFFFF
    cfg.eps, cfg.momentum, cfg.running_mean, cfg.running_var = 1e-5, 0.1, None, None
    gamma = beta = None
    nm = blk.norm
    if nm is None:
        cfg.mode = UB_NORM_NONE
    else:
        cfg.eps = float(nm.eps)
        gamma, beta = nm.weight, nm.bias
        if blk.norm_kind == "instance":
            cfg.mode = UB_NORM_INSTANCE
        else:
            cfg.mode = UB_NORM_BATCH_TRAIN if (nm.training or not nm.track_running_stats) else UB_NORM_BATCH_EVAL
            cfg.running_mean, cfg.running_var = nm.running_mean, nm.running_var
            if cfg.mode == UB_NORM_BATCH_TRAIN and nm.running_mean is not None:
                cfg.momentum = _bn_momentum(nm)
            if cfg.mode == UB_NORM_BATCH_TRAIN and nm.num_batches_tracked is not None:
                nm.num_batches_tracked.add_(1)
    return _BlockFn.apply(cfg, src0, src1, blk.conv.weight, blk.conv.bias, gamma, beta)


def generator_forward(net, x, fresh_seed):
    """``modules._UNetGraph`` on an NCDHW fp32 input -> (B, 6, D, H, W) fp32 (ref:src/model.py:36-39)."""
    base_seed = fresh_seed() if any(b.dropout_p() > 0.0 for b in net.blocks) else 0
    lid = [0]

    def run(blk, s0, s1=None, pool=False):
        lid[0] += 1
        return run_block(blk, s0, s1, base_seed + 7919 * lid[0], pool=pool)

    cur = x.float()
    if net.head is not None:
        cur, _ = run(net.head, cur)
    skips = []
    for lvl, (c0, c1) in enumerate(net.enc):
        t, _ = run(c0, cur)
        last = lvl == len(net.enc) - 1
        xk, pooled = run(c1, t, pool=not last)
        skips.append(xk)
        cur = xk if last else pooled
    u = skips[-1]
    for j, (dc, c0, c1) in enumerate(net.dec):
        up, _ = run(dc, u)
        t, _ = run(c0, skips[-2 - j], up)          # cat[x_e, up] folded into the two-source conv
        u, _ = run(c1, t)
    out, _ = run(net.final, u)
    return out.to(x.dtype)


def chain_forward(chain, x, y=None):
    """``modules._Chain`` (Discriminator / a standalone DownSampleConv): cat[x, y] is the first conv's two sources."""
    a, s1 = x.float(), (None if y is None else y.float())
    for blk in chain.blocks:
        a, _ = run_block(blk, a, s1, 0)
        s1 = None
    return a.to(x.dtype)
