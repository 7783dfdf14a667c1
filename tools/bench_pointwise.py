"""Micro-benchmark of the memory-bound kernels at the bench shape (8 x 128^3 x 32 channels):
effective GB/s = algorithmic bytes (each tensor once) / time.   python tools/bench_pointwise.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import unet_bssfp_b200 as ub  # noqa: E402
from unet_bssfp_b200 import _lib  # noqa: E402

ops = ub.ops
dev = "cuda"
N, S, CP = 8, 128, 32
if len(sys.argv) > 1:
    N, S, CP = map(int, sys.argv[1:4])
g = torch.Generator(device=dev).manual_seed(0)
y = torch.randn((N, S, S, S, CP), device=dev, generator=g).to(torch.bfloat16)
dA = torch.randn((N, S, S, S, CP), device=dev, generator=g).to(torch.bfloat16)
scale = torch.rand((N, CP), device=dev) + 0.5
shift = torch.randn((N, CP), device=dev) * 0.1
mean = torch.randn((N, CP), device=dev) * 0.1
rstd = torch.rand((N, CP), device=dev) + 0.5
x = torch.rand((N, 24, S, S, S), device=dev)
yy = torch.rand((N, 6, S, S, S), device=dev)
elems = y.numel()


def timeit(name, fn, nbytes, iters=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = min(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    print(f"{name:44s} {ms:7.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s")


a, _ = ops.norm_act_fwd(y, scale, shift, 0.1, 0.05, 123)
I, NONE = _lib.UB_NORM_INSTANCE, _lib.UB_NORM_NONE
timeit("norm_act_fwd p=0.05", lambda: ops.norm_act_fwd(y, scale, shift, 0.1, 0.05, 123), elems * 4)
timeit("norm_act_fwd p=0", lambda: ops.norm_act_fwd(y, scale, shift, 0.1, 0.0, 0), elems * 4)
timeit("norm_act_fwd+pool p=0.05", lambda: ops.norm_act_fwd(y, scale, shift, 0.1, 0.05, 123, pool=True), elems * 4.25)
timeit("norm_act_bwd from a, p=0.05 (14 B/el)", lambda: ops.norm_act_bwd(dA, a, y, I, mean, rstd, scale, 0.1, 0.05, 123, CP), elems * 14)
timeit("norm_act_bwd from y, p=0.05 (10 B/el)", lambda: ops.norm_act_bwd(dA, None, y, I, mean, rstd, scale, 0.1, 0.05, 123, CP, shift=shift), elems * 10)
timeit("norm_act_bwd from y, p=0    (10 B/el)", lambda: ops.norm_act_bwd(dA, None, y, I, mean, rstd, scale, 0.1, 0.0, 0, CP, shift=shift), elems * 10)
timeit("act_bwd (no norm) p=0 (6 B/el)", lambda: ops.norm_act_bwd(dA, a, None, NONE, None, None, None, 0.2, 0.0, 0, CP), elems * 6)
pooled = torch.randn((N, S // 2, S // 2, S // 2, CP), device=dev, generator=g).to(torch.bfloat16)
timeit("maxpool_bwd accumulate", lambda: ops.maxpool_bwd(a, pooled, dA), elems * 6.25)
timeit("colsum", lambda: ops.colsum(y, CP), elems * 2)
if CP == 32:
    timeit("pack_ncdhw x(24)", lambda: ops.pack_ncdhw(x), x.numel() * 4 + elems * 2)
    timeit("pack_ncdhw cat[x, y] s2d", lambda: ops.pack_ncdhw(x, yy, s2d=True), (x.numel() + yy.numel()) * 4 + elems * 2)
    timeit("unpack_ncdhw 6ch", lambda: ops.unpack_ncdhw(y, 6), elems * 2 + yy.numel() * 4)
    timeit("l1_fwd", lambda: ops.l1_fwd(yy, yy), yy.numel() * 8)

# ---- experiment: do power-of-two distances between the streams (each tensor is exactly 2^30 B at the
# bench shape) cost DRAM bandwidth? Re-run with the tensors carved out of one buffer at skewed offsets.
if len(sys.argv) > 4:
    skew = int(sys.argv[4])
    big = torch.empty(4 * elems + 4 * 1024 * 1024, dtype=torch.bfloat16, device=dev)

    def carve(k):
        off = k * (elems + skew // 2)
        return big[off:off + elems].view(N, S, S, S, CP)

    y2, dA2, a2 = carve(0), carve(1), carve(2)
    y2.copy_(y); dA2.copy_(dA); a2.copy_(a)
    print(f"skew {skew} B: ptr deltas {dA2.data_ptr() - y2.data_ptr()}")
    import ctypes as C
    lib = _lib.load()
    out = carve(3)

    def raw_bwd(dAx, yx, outx):
        ws = torch.empty(lib.ub_norm_act_bwd_workspace_bytes(N, CP) // 4, dtype=torch.float32, device=dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(lib.ub_norm_act_bwd(p(dAx), None, p(yx), I, p(mean), p(rstd), p(scale), p(shift), 0.1, 0.05, 123, N,
                                       S * S * S, CP, CP, p(ws), p(outx), None, None, None, st))

    def raw_fwd(yx, outx):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(lib.ub_norm_act_fwd(p(yx), p(scale), p(shift), 0.1, 0.05, 123, N, S, S, S, CP, p(outx), None, None, st))

    timeit("  skewed norm_act_bwd from y p=0.05", lambda: raw_bwd(dA2, y2, out), elems * 10)
    timeit("  skewed norm_act_fwd p=0.05", lambda: raw_fwd(y2, out), elems * 4)
    o2 = torch.empty_like(y)
    timeit("  unskewed raw norm_act_bwd from y p=0.05", lambda: raw_bwd(dA, y, o2), elems * 10)
    timeit("  unskewed raw norm_act_fwd p=0.05", lambda: raw_fwd(y, o2), elems * 4)
if CP == 32 and len(sys.argv) <= 4:
    wt = torch.randn((6, 32, 1, 1, 1), device=dev) * 0.1
    bb = torch.randn((6,), device=dev)
    go = torch.randn((N, 6, S, S, S), device=dev)
    timeit("conv1x1_to_ncdhw (88 B/vox)", lambda: ops.conv1x1_to_ncdhw(y, wt, bb), elems // 32 * 88)
    timeit("conv1x1_from_ncdhw_bwd all (152 B/vox)", lambda: ops.conv1x1_from_ncdhw_bwd(go, y, wt), elems // 32 * 152)
    timeit("conv1x1_from_ncdhw_bwd du only (88 B/vox)", lambda: ops.conv1x1_from_ncdhw_bwd(go, y, wt, need_params=False), elems // 32 * 88)
