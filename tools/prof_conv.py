"""Run one convolution launch shape of the hot path (for ncu): python tools/prof_conv.py fwd|dgrad|wgrad kind c0 c1 co n d h w"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_bssfp_b200 as ub
ops = ub.ops
what, kind, c0, c1, co, n, d, h, w = sys.argv[1], *map(int, sys.argv[2:10])
iters = int(sys.argv[10]) if len(sys.argv) > 10 else 3
dev = "cuda"
spec = ops.ConvSpec(kind, c0, co, c1)
g = torch.Generator(device=dev).manual_seed(0)
s0 = torch.randn((n, d, h, w, spec.c0p), device=dev, generator=g).to(torch.bfloat16)
if kind == 4:   # space-to-depth source layout
    s0 = s0.view(n, 8, d // 2, h // 2, w // 2, spec.c0p)
s1 = torch.randn((n, d, h, w, spec.c1p), device=dev, generator=g).to(torch.bfloat16) if c1 else None
k = {0: 3, 1: 1, 2: 4, 3: 2, 4: 4}[kind]
wshape = (c0 + c1, co, k, k, k) if kind == 3 else (co, c0 + c1, k, k, k)
wt = torch.randn(wshape, device=dev, generator=g) * 0.05
od, oh, ow = spec.out_dims(d, h, w)
dy = torch.randn((n, od, oh, ow, spec.cop), device=dev, generator=g).to(torch.bfloat16)
bias = torch.zeros(co, device=dev)
wf, wd = ops.pack_conv_weights(spec, wt, 0), ops.pack_conv_weights(spec, wt, 1)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
torch.cuda.synchronize()
ev[0].record()
for i in range(iters):
    if what == "fwd":
        ops.conv_fwd(spec, s0, s1, wf, bias, want_stats=(kind != 3))
    elif what == "dgrad":
        ops.conv_dgrad(spec, dy, wd, (d, h, w))
    else:
        ops.conv_wgrad(spec, s0, s1, dy, wshape)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
flops = 2.0 * n * d * h * w * (c0 + c1) * co * (k ** 3 if kind != 2 else 8)
print(f"{what} kind={kind} {c0}+{c1}->{co} {n}x{d}x{h}x{w}: ms {['%.3f' % t for t in ts]}  {flops / min(ts) / 1e9:.1f} TFLOP/s")
