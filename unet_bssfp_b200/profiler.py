"""Per-op CUDA-event timing of the hot path: wraps the ``ops.*`` entry points the modules call, attributes every
call to the kernel class that serves it (``ub_conv_kernel_class``) and to its algorithmic work -- conv / deconv
FLOPs (2 x MACs) for the tensor-core kernels, tensor bytes (inputs read once + outputs written once) for the
memory-bound ones. Used by ``bench.py`` (the per-class roofline table of the bench line) and ``tools/profile_step.py``.

Timing a step through this profiler adds one event pair per op and runs the weight gradients on the main stream, so
the per-class times are serial times; the step time reported by ``bench.py`` comes from an unprofiled loop.
"""
from __future__ import annotations

import collections
import contextlib
import ctypes as C

import torch

from . import _lib, modules, ops

KIND = {0: "k3", 1: "k1", 2: "k4s2", 3: "dc2", 4: "k4s2d"}
TAPS = {0: 27, 1: 1, 2: 64, 3: 8, 4: 64}
CONV_CLASS = {0: "igemm_fwd_kernel", 1: "igemm_march_kernel", 2: "wgrad_march_kernel", 3: "igemm_wgrad_kernel"}
MEMORY_OPS = ("conv1x1_to_ncdhw", "conv1x1_from_ncdhw_bwd", "pack_ncdhw", "unpack_ncdhw", "norm_finalize", "norm_act_fwd",
              "norm_act_bwd", "maxpool_bwd", "maxpool_bwd_fused", "head_bwd_fused", "colsum", "l1_fwd", "l1_bwd", "bce_logits", "scale_by",
              "pack_conv_weights", "pack_conv_weights_multi")


def _tensors(objs):
    for o in objs:
        if torch.is_tensor(o):
            yield o
        elif isinstance(o, ops.DeferredAct):
            yield o.y
        elif isinstance(o, (tuple, list)):
            yield from _tensors(o)


def _nbytes(*objs):
    return sum(t.numel() * t.element_size() for t in _tensors(objs))


def conv_flops(spec, n, d, h, w):
    v = n * d * h * w
    if spec.kind in (2, 4):
        v //= 8
    return 2.0 * v * (spec.c0 + spec.c1) * spec.co * TAPS[spec.kind]


def _kernel_class(spec, n, d, h, w, direction):
    desc = spec.desc(n, d, h, w)
    return CONV_CLASS[_lib.load().ub_conv_kernel_class(C.byref(desc), direction)]


def _label(spec, n, d, extra=""):
    return f"{KIND[spec.kind]} {spec.c0}+{spec.c1}->{spec.co} n{n} in{d}{extra}"


class Record:
    __slots__ = ("op", "cls", "label", "e0", "e1", "flops", "nbytes")


@contextlib.contextmanager
def profile():
    """``with profile() as records:`` -- every wrapped op called inside appends a Record; call
    ``torch.cuda.synchronize()`` and then ``summarize(records)``."""
    records = []
    saved = {}

    def wrap(name, describe):
        fn = getattr(ops, name)
        saved[name] = fn

        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            r = Record()
            r.op, r.e0, r.e1 = name, e0, e1
            r.cls, r.label, r.flops, r.nbytes = describe(out, *a, **k)
            records.append(r)
            return out

        setattr(ops, name, inner)

    def d_fwd(out, spec, src0, src1, *a, **k):
        n, d, h, w = spec.in_dims(src0)
        tag = " (deferred src0)" if isinstance(src0, ops.DeferredAct) else ""
        return _kernel_class(spec, n, d, h, w, 0), _label(spec, n, d, " fwd" + tag), conv_flops(spec, n, d, h, w), \
            _nbytes(src0, src1, out[0])

    def d_dgrad(out, spec, dy, wp, in_dhw, fuse=None):
        n = dy.shape[0]
        d, h, w = in_dhw
        tag = " +norm-bwd sums" if fuse is not None else ""
        return _kernel_class(spec, n, d, h, w, 1), _label(spec, n, d, " dgrad" + tag), conv_flops(spec, n, d, h, w), \
            _nbytes(dy, out[0], out[1])

    def d_wgrad(out, spec, src0, src1, dy, shape):
        n, d, h, w = spec.in_dims(src0)
        tag = " (deferred src0)" if isinstance(src0, ops.DeferredAct) else ""
        return _kernel_class(spec, n, d, h, w, 2), _label(spec, n, d, " wgrad" + tag), conv_flops(spec, n, d, h, w), \
            _nbytes(src0, src1, dy)

    def d_mem(name):
        def describe(out, *a, **k):
            objs = [getattr(o, "y", o) if isinstance(o, ops.NormBwdFusion) else o for o in list(a) + list(k.values())]
            ts = list(_tensors(objs + [out]))
            big = max(ts, key=lambda t: t.numel()) if ts else None
            return name, f"{tuple(big.shape) if big is not None else ()}", 0.0, _nbytes(*ts)
        return describe

    wrap("conv_fwd", d_fwd)
    wrap("conv_dgrad", d_dgrad)
    wrap("conv_wgrad", d_wgrad)
    for nm in MEMORY_OPS:
        wrap(nm, d_mem(nm))
    side = modules._WGRAD_SIDE
    modules._WGRAD_SIDE = False          # serial per-op times
    try:
        yield records
    finally:
        modules._WGRAD_SIDE = side
        for name, fn in saved.items():
            setattr(ops, name, fn)


def summarize(records, steps=1):
    """-> (per_class, per_op): ordered dicts keyed by kernel class / (op, label) with calls, ms, flops, nbytes
    (totals divided by ``steps``)."""
    per_class = collections.OrderedDict()
    per_op = collections.OrderedDict()
    for r in records:
        ms = r.e0.elapsed_time(r.e1)
        for table, key in ((per_class, r.cls), (per_op, (r.op, r.cls, r.label))):
            a = table.setdefault(key, {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            a["calls"] += 1
            a["ms"] += ms
            a["flops"] += r.flops
            a["bytes"] += r.nbytes
    for table in (per_class, per_op):
        for a in table.values():
            for k in a:
                a[k] = a[k] / steps
    return per_class, per_op
