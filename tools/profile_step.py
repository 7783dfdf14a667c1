"""Per-op timing of one GAN step: wraps every ``ops.*`` entry with CUDA events (current stream) and
prints a table sorted by time, with TFLOP/s for the convolutions and GB/s for the memory-bound ops.

    python tools/profile_step.py [--batch 8] [--size 128] [--out gpurun_out/step_profile.txt]
"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import unet_bssfp_b200 as ub  # noqa: E402
from unet_bssfp_b200 import modules, ops  # noqa: E402
from unet_bssfp_b200.train_step import GanTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--modality", default="bssfp")
ap.add_argument("--out", default="")
args = ap.parse_args()

dev = torch.device("cuda:0")
records = []   # (label, ev0, ev1, flops, bytes)
KIND = {0: "k3", 1: "k1", 2: "k4s2", 3: "dc2", 4: "k4s2d"}
TAPS = {0: 27, 1: 1, 2: 64, 3: 8, 4: 64}


def _nbytes(*ts):
    return sum(t.numel() * t.element_size() for t in ts if torch.is_tensor(t))


def wrap(name, labeller):
    fn = getattr(ops, name)

    def inner(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        label, flops, nbytes = labeller(out, *a, **k)
        records.append((name, label, e0, e1, flops, nbytes))
        return out

    setattr(ops, name, inner)


def conv_label(spec, vox_in, n):
    return f"{KIND[spec.kind]} {spec.c0}+{spec.c1}->{spec.co} n{n} in{vox_in}"


def conv_flops(spec, n, d, h, w):
    v = n * d * h * w
    if spec.kind in (2, 4):
        v //= 8
    return 2.0 * v * (spec.c0 + spec.c1) * spec.co * TAPS[spec.kind]


def l_conv_fwd(out, spec, src0, src1, *a, **k):
    n, d, h, w = spec.in_dims(src0)
    return conv_label(spec, d, n), conv_flops(spec, n, d, h, w), _nbytes(src0, src1, out[0])


def l_conv_dgrad(out, spec, dy, wp, in_dhw, fuse=None):
    n = dy.shape[0]
    d, h, w = in_dhw
    tag = " +nbwd" if fuse is not None else ""
    return conv_label(spec, d, n) + tag, conv_flops(spec, n, d, h, w), _nbytes(dy, out[0], out[1])


def l_conv_wgrad(out, spec, src0, src1, dy, shape):
    n, d, h, w = spec.in_dims(src0)
    return conv_label(spec, d, n), conv_flops(spec, n, d, h, w), _nbytes(src0, src1, dy)


def l_generic(out, *a, **k):
    outs = out if isinstance(out, (tuple, list)) else (out,)
    ts = [t for t in list(a) + list(outs) if torch.is_tensor(t)]
    big = max(ts, key=lambda t: t.numel())
    return f"{tuple(big.shape)}", 0.0, _nbytes(*ts)


wrap("conv_fwd", l_conv_fwd)
wrap("conv_dgrad", l_conv_dgrad)
wrap("conv_wgrad", l_conv_wgrad)
for nm in ("conv1x1_to_ncdhw", "conv1x1_from_ncdhw_bwd", "pack_ncdhw", "unpack_ncdhw", "norm_finalize", "norm_act_fwd", "norm_act_bwd", "maxpool_bwd", "colsum",
           "l1_fwd", "l1_bwd", "bce_logits", "scale_by", "pack_conv_weights"):
    wrap(nm, l_generic)

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
records.clear()
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
tr.step(x, y)
s1.record()
torch.cuda.synchronize()
total = s0.elapsed_time(s1)

agg = collections.OrderedDict()
for name, label, e0, e1, fl, nb in records:
    k = (name, label)
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
    a[2] += fl
    a[3] += nb
lines = []
covered = sum(v[1] for v in agg.values())
lines.append(f"step {total:.2f} ms (with event overhead); ops covered {covered:.2f} ms; batch {B} size {S}")
byname = collections.defaultdict(float)
for (name, _), v in agg.items():
    byname[name] += v[1]
for name, ms in sorted(byname.items(), key=lambda kv: -kv[1]):
    lines.append(f"  {name:18s} {ms:8.3f} ms")
lines.append(f"{'op':18s} {'shape':34s} {'n':>3s} {'ms':>8s} {'ms/call':>8s} {'TFLOP/s':>8s} {'GB/s':>8s}")
for (name, label), (cnt, ms, fl, nb) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tf = fl / (ms * 1e-3) / 1e12 if fl else 0.0
    gb = nb / (ms * 1e-3) / 1e9
    lines.append(f"{name:18s} {label:34s} {cnt:3d} {ms:8.3f} {ms / cnt:8.3f} {tf:8.1f} {gb:8.0f}")
txt = "\n".join(lines)
print(txt)
if args.out:
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        f.write(txt + "\n")
