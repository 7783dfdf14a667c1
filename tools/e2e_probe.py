"""Where does the end-to-end step lose time against the device-resident step?
  (a) H2D of one batch alone (x bf16 + y fp32, pinned) -> GB/s
  (b) device-resident loop on fp32 x
  (c) device-resident loop on bf16 x (transport dtype, no H2D)
  (d) the bench's e2e loop (double-buffered H2D on a copy stream)
  (e) the same loop, copies issued but from a DEVICE staging tensor (no PCIe traffic)
    python tools/e2e_probe.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import unet_bssfp_b200 as ub  # noqa: E402
from unet_bssfp_b200.train_step import GanTrainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
torch.manual_seed(0)
tr = GanTrainer(ub.Generator("bssfp").to(dev), ub.Discriminator("bssfp").to(dev))
B, S = 8, 128
host = [(torch.rand(B, 24, S, S, S).to(torch.bfloat16).pin_memory(), torch.rand(B, 6, S, S, S).pin_memory()) for _ in range(2)]
dev_f32 = [(xh.to(dev).float(), yh.to(dev)) for xh, yh in host]
dev_b16 = [(xh.to(dev), yh.to(dev)) for xh, yh in host]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, n):
    torch.cuda.synchronize()
    e0.record()
    fn(n)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# (a)
bx, by = torch.empty_like(dev_b16[0][0]), torch.empty_like(dev_b16[0][1])
cs = torch.cuda.Stream()


def h2d(n):
    with torch.cuda.stream(cs):
        for i in range(n):
            bx.copy_(host[i % 2][0], non_blocking=True)
            by.copy_(host[i % 2][1], non_blocking=True)
    torch.cuda.current_stream().wait_stream(cs)


h2d(1)
ms = timed(h2d, 4)
nbytes = bx.numel() * 2 + by.numel() * 4
print(f"(a) H2D alone: {ms:.2f} ms per batch of {nbytes / 1e9:.2f} GB = {nbytes / ms / 1e6:.1f} GB/s")


def resident(batches):
    def run(n):
        for i in range(n):
            tr.step(*batches[i % 2])
    return run


for _ in range(3):
    tr.step(*dev_f32[0])
print(f"(b) resident fp32 x: {timed(resident(dev_f32), steps):.2f} ms/step")
tr.step(*dev_b16[0])
print(f"(c) resident bf16 x: {timed(resident(dev_b16), steps):.2f} ms/step")

bufs = [(torch.empty_like(dev_b16[0][0]), torch.empty_like(dev_b16[0][1])) for _ in range(2)]
ready = [torch.cuda.Event() for _ in range(2)]
done = [torch.cuda.Event() for _ in range(2)]
loss_host = torch.zeros(2, 2).pin_memory()


def make_loop(src):
    def prefetch(i):
        with torch.cuda.stream(cs):
            cs.wait_event(done[i % 2])
            bufs[i % 2][0].copy_(src[i % 2][0], non_blocking=True)
            bufs[i % 2][1].copy_(src[i % 2][1], non_blocking=True)
            ready[i % 2].record(cs)

    def loop(n):
        for ev in done:
            ev.record()
        prefetch(0)
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            gl, dl = tr.step(*bufs[i % 2])
            done[i % 2].record()
            loss_host[i % 2, 0].copy_(gl, non_blocking=True)
            loss_host[i % 2, 1].copy_(dl, non_blocking=True)
    return loop


make_loop(host)(2)
print(f"(d) e2e loop, H2D from pinned host: {timed(make_loop(host), steps):.2f} ms/step")
print(f"(e) same loop, copies from device memory: {timed(make_loop(dev_b16), steps):.2f} ms/step")
print(f"(d) again: {timed(make_loop(host), steps):.2f} ms/step")
print(f"(b) again: {timed(resident(dev_f32), steps):.2f} ms/step")
