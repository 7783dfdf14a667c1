"""Wall time of the phases of one GAN step (CUDA events on the main stream, steady state):
G forward, D forward on the fake, losses, backward (D dgrad chain + G backward), AdamW(G), then the D phase:
G forward (no grad), D forward x2, backward x2, AdamW(D).   python tools/phase_times.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import unet_bssfp_b200 as ub  # noqa: E402
from unet_bssfp_b200.train_step import GanTrainer, _set_requires_grad  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda:0")
torch.manual_seed(0)
tr = GanTrainer(ub.Generator("bssfp").to(dev), ub.Discriminator("bssfp").to(dev))
B, S = 8, 128
bs = [(torch.rand(B, 24, S, S, S, device=dev), torch.rand(B, 6, S, S, S, device=dev)) for _ in range(2)]
for i in range(3):
    tr.step(*bs[i % 2])
torch.cuda.synchronize()
names = ["G fwd", "D fwd (fake, G phase)", "losses", "backward D dgrad + G", "AdamW G", "G fwd (no grad)", "D fwd fake",
         "D fwd real", "losses D", "backward D x2", "AdamW D"]
acc = [0.0] * len(names)
tot = 0.0
for it in range(steps):
    x, y = bs[it % 2]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    k = [0]

    def mark():
        ev[k[0]].record()
        k[0] += 1

    mark()
    _set_requires_grad(tr.discr, False)
    y_hat = tr.gen(x); mark()
    logits = tr.discr(x, y_hat); mark()
    g_loss = tr.bce(logits, torch.ones_like(logits)) + tr.recon_loss(y_hat, y); mark()
    g_loss.backward(); mark()
    tr._reduce_and_step(tr.reduce_g, tr.opt_g)
    _set_requires_grad(tr.discr, True); mark()
    _set_requires_grad(tr.gen, False)
    with torch.no_grad():
        y_hat = tr.gen(x)
    mark()
    logits_hat = tr.discr(x, y_hat); mark()
    logits = tr.discr(x, y); mark()
    d_loss = (tr.bce(logits, torch.ones_like(logits)) + tr.bce(logits_hat, torch.zeros_like(logits_hat))) / 2; mark()
    d_loss.backward(); mark()
    tr._reduce_and_step(tr.reduce_d, tr.opt_d)
    _set_requires_grad(tr.gen, True); mark()
    torch.cuda.synchronize()
    for i in range(len(names)):
        acc[i] += ev[i].elapsed_time(ev[i + 1])
    tot += ev[0].elapsed_time(ev[-1])
print(f"step {tot / steps:.2f} ms (events between phases)")
for nm, a in zip(names, acc):
    print(f"  {nm:28s} {a / steps:7.2f} ms")

# host enqueue time of a step (no synchronisation inside): how far ahead of the GPU does the CPU run?
import time
torch.cuda.synchronize()
t_enq = []
for it in range(steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tr.step(*bs[it % 2])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    t_enq.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
print("host enqueue ms / step wall ms (from an idle GPU):", ", ".join(f"{a:.1f}/{b:.1f}" for a, b in t_enq))
