"""Run the memory-bound kernels of the step at the bench shape (for ncu / timing):
    python tools/prof_mem.py head_fwd|head_bwd|head_bwd_fused|nab|naf|pool|poolbwd|pool_fused [iters]
head_fwd / head_bwd: the fused output head (1x1x1 conv + layout change) on a deferred activation;
nab: norm_act_bwd (reduce + apply, from y); naf: norm_act_fwd; pool / poolbwd: the pooling passes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import unet_bssfp_b200 as ub  # noqa: E402
from unet_bssfp_b200 import _lib  # noqa: E402

ops = ub.ops
what = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda"
N, S, CP = 8, 128, 32
g = torch.Generator(device=dev).manual_seed(0)
y = torch.randn((N, S, S, S, CP), device=dev, generator=g).to(torch.float16)
dA = torch.randn((N, S, S, S, CP), device=dev, generator=g).to(torch.bfloat16)
scale = torch.rand((N, CP), device=dev) + 0.5
shift = torch.randn((N, CP), device=dev) * 0.1
mean = torch.randn((N, CP), device=dev) * 0.1
rstd = torch.rand((N, CP), device=dev) + 0.5
w = torch.randn((6, 32, 1, 1, 1), device=dev) * 0.1
b = torch.zeros(6, device=dev)
dout = torch.randn((N, 6, S, S, S), device=dev, generator=g)
u = ops.DeferredAct(y, scale, shift, 0.1, 0.05, 123)
I = _lib.UB_NORM_INSTANCE
pooled = torch.randn((N, S // 2, S // 2, S // 2, CP), device=dev, generator=g).to(torch.bfloat16)
a = None
if what == "poolbwd":
    a, _ = ops.norm_act_fwd(y, scale, shift, 0.1, 0.05, 123)

fuse = ops.NormBwdFusion(y, scale, shift, mean, rstd, 0.1, 0.05, 123)
fn = {
    "head_bwd_fused": lambda: ops.head_bwd_fused(dout, fuse, w, I, 32),
    "pool_fused": lambda: ops.maxpool_bwd_fused(pooled, dA, fuse),
    "head_fwd": lambda: ops.conv1x1_to_ncdhw(u, w, b),
    "head_bwd": lambda: ops.conv1x1_from_ncdhw_bwd(dout, u, w, need_input=True, need_params=True),
    "nab": lambda: ops.norm_act_bwd(dA, None, y, I, mean, rstd, scale, 0.1, 0.05, 123, CP, shift=shift),
    "naf": lambda: ops.norm_act_fwd(y, scale, shift, 0.1, 0.05, 123),
    "pool": lambda: ops.norm_act_fwd(y, scale, shift, 0.1, 0.05, 123, pool=True),
    "poolbwd": lambda: ops.maxpool_bwd(a, pooled, dA),
}[what]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
fn()
torch.cuda.synchronize()
ev[0].record()
for i in range(iters):
    fn()
    ev[i + 1].record()
torch.cuda.synchronize()
print(what, ["%.3f" % ev[i].elapsed_time(ev[i + 1]) for i in range(iters)], "ms")
