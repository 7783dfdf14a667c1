"""Config 5: whole-volume sliding-window inference (160 x 192 x 160, 27 patches of 64^3) + relative-error
evaluation, eager launches vs CUDA-graph replay of the generator forward.   python tools/bench_inference.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import unet_bssfp_b200 as ub  # noqa: E402

dev = "cuda"
shape = (160, 192, 160)
torch.manual_seed(0)
g = ub.Generator("bssfp").to(dev).eval()
vol = torch.rand((24,) + shape, device=dev)
tgt = torch.rand((6,) + shape, device=dev) * 0.95 + 0.05
mask = (torch.rand(shape, device=dev) > 0.3).to(torch.uint8)
probseg = torch.rand(shape + (3,), device=dev)


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


vox = shape[0] * shape[1] * shape[2]
for batch in (8, 9, 27):
    e = timed(lambda: ub.inference.predict_volume(g, vol, patch=64, batch=batch, use_graph=False))
    c = timed(lambda: ub.inference.predict_volume(g, vol, patch=64, batch=batch, use_graph=True))
    print(f"predict_volume batch {batch:2d}: eager {e:7.2f} ms ({vox / e / 1e3:7.1f} Mvox/s)   graph {c:7.2f} ms ({vox / c / 1e3:7.1f} Mvox/s)")
pred = ub.inference.predict_volume(g, vol, patch=64, batch=9)
r = timed(lambda: ub.inference.relative_error(pred, tgt, mask, probseg))
print(f"relative_error (map + ROI means): {r:.3f} ms")
m = timed(lambda: ub.ops.dti_scalar_maps(pred.permute(1, 2, 3, 0).contiguous()))
print(f"dti_scalar_maps: {m:.3f} ms")
n = timed(lambda: ub.nifti.volume_to_nifti_order(pred, (-0.003, 0.004)))
print(f"denorm_to_nifti: {n:.3f} ms")
