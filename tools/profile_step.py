"""Per-op timing of one GAN step (``unet_bssfp_b200.profiler``): a table sorted by time, with TFLOP/s for the
convolutions and GB/s for the memory-bound ops, and the per-kernel-class totals.

    python tools/profile_step.py [--batch 8] [--size 128] [--out gpurun_out/step_profile.txt]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import unet_bssfp_b200 as ub  # noqa: E402
from unet_bssfp_b200 import profiler  # noqa: E402
from unet_bssfp_b200.train_step import GanTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--modality", default="bssfp")
ap.add_argument("--out", default="")
args = ap.parse_args()

dev = torch.device("cuda:0")
torch.manual_seed(0)
gen = ub.Generator(args.modality).to(dev)
dis = ub.Discriminator(args.modality).to(dev)
tr = GanTrainer(gen, dis)
cin = 24 if args.modality in ("bssfp", "pc-bssfp") else 6
B, S = args.batch, args.size
x = torch.rand(B, cin, S, S, S, device=dev)
y = torch.rand(B, 6, S, S, S, device=dev)
for _ in range(2):
    tr.step(x, y)
torch.cuda.synchronize()
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profiler.profile() as records:
    s0.record()
    tr.step(x, y)
    s1.record()
    torch.cuda.synchronize()
total = s0.elapsed_time(s1)
per_class, per_op = profiler.summarize(records)

lines = []
covered = sum(v["ms"] for v in per_class.values())
lines.append(f"step {total:.2f} ms (with event overhead, weight gradients on the main stream); ops covered {covered:.2f} ms; batch {B} size {S}")
for cls, v in sorted(per_class.items(), key=lambda kv: -kv[1]["ms"]):
    tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["flops"] else 0.0
    gb = v["bytes"] / (v["ms"] * 1e-3) / 1e9
    lines.append(f"  {cls:26s} {v['ms']:8.3f} ms {int(v['calls']):4d} calls {tf:8.1f} TFLOP/s {gb:8.0f} GB/s")
lines.append(f"{'op':18s} {'shape':46s} {'n':>3s} {'ms':>8s} {'ms/call':>8s} {'TFLOP/s':>8s} {'GB/s':>8s}")
for (op, cls, label), v in sorted(per_op.items(), key=lambda kv: -kv[1]["ms"]):
    tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["flops"] else 0.0
    gb = v["bytes"] / (v["ms"] * 1e-3) / 1e9
    lines.append(f"{op:18s} {label:46s} {int(v['calls']):3d} {v['ms']:8.3f} {v['ms'] / v['calls']:8.3f} {tf:8.1f} {gb:8.0f}")
txt = "\n".join(lines)
print(txt)
if args.out:
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        f.write(txt + "\n")
