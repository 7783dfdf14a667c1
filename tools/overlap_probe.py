"""Experiment: how much does the GPU gain when two independent half-batch GAN steps are in flight on two streams
(tensor-bound and memory-bound kernels of different streams co-scheduled) versus one full-batch step on one
stream?  python tools/overlap_probe.py [steps]      (measurement of overlap potential only: not the bench)"""
import os
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import unet_bssfp_b200 as ub  # noqa: E402
from unet_bssfp_b200.train_step import GanTrainer  # noqa: E402

dev = "cuda"
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
S = 128


def make(batch):
    torch.manual_seed(0)
    g, d = ub.Generator("bssfp").to(dev), ub.Discriminator("bssfp").to(dev)
    tr = GanTrainer(g, d)
    x = torch.rand(batch, 24, S, S, S, device=dev)
    y = torch.rand(batch, 6, S, S, S, device=dev)
    return tr, x, y


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(n)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


tr8, x8, y8 = make(8)
for _ in range(3):
    tr8.step(x8, y8)
one = timed(lambda n: [tr8.step(x8, y8) for _ in range(n)], steps)
print(f"one stream, batch 8: {one:.2f} ms/step")
del tr8, x8, y8
torch.cuda.empty_cache()

pairs = [make(4), make(4)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def worker(i, n):
    tr, x, y = pairs[i]
    with torch.cuda.stream(streams[i]):
        for _ in range(n):
            tr.step(x, y)


def both(n):
    main = torch.cuda.current_stream()
    for s in streams:
        s.wait_stream(main)
    ts = [threading.Thread(target=worker, args=(i, n)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for s in streams:
        main.wait_stream(s)


both(3)
two = timed(both, steps)
print(f"two streams, 2 x batch 4 (two host threads): {two:.2f} ms per pair of half steps")
with torch.cuda.stream(streams[0]):
    for _ in range(2):
        pairs[0][0].step(pairs[0][1], pairs[0][2])
half = timed(lambda n: worker(0, n), steps)
print(f"one stream, batch 4: {half:.2f} ms/step (x2 = {2 * half:.2f})")
