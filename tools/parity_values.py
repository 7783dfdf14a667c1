"""Measured end-to-end parity numbers (generator eval forward rel-L2 against the fp32 oracle, and the bf16-autocast
yardstick) for the shapes the tests and smoke() use.   python tools/parity_values.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_bssfp_b200 as ub
from oracle import model_oracle as O
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
rel = lambda a, b: ((a - b).norm() / b.norm()).item()
for mod, shape, seed in [("bssfp", (1, 32, 32, 32), 0), ("t1w", (2, 32, 48, 32), 0), ("bssfp", (1, 64, 64, 64), 0),
                         ("bssfp", (1, 128, 128, 128), 3), ("t1w", (1, 128, 128, 128), 3), ("bssfp", (2, 96, 96, 96), 1)]:
    torch.manual_seed(seed)
    og, od = O.Generator(mod).to(DEV), O.Discriminator(mod).to(DEV)
    g, d = ub.Generator(mod).to(DEV), ub.Discriminator(mod).to(DEV)
    g.load_state_dict(og.state_dict()); d.load_state_dict(od.state_dict())
    torch.manual_seed(1234)
    n, dd, hh, ww = shape
    x = torch.rand(n, O.in_channels_of(mod), dd, hh, ww, device=DEV)
    y = torch.rand(n, 6, dd, hh, ww, device=DEV)
    og.eval(); g.eval()
    with torch.no_grad():
        ref, got = og(x), g(x)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yard = og(x).float()
    line = f"{mod:6s} {str(shape):20s} G eval rel-L2 {rel(got, ref):.3e}  (bf16 autocast of the oracle {rel(yard, ref):.3e})"
    if dd % 32 == 0 and hh % 32 == 0 and ww % 32 == 0 and n * (dd // 32) * (hh // 32) * (ww // 32) > 1:
        d.eval(); od.eval()
        with torch.no_grad():
            line += f"   D eval logits {rel(d(x, y), od(x, y)):.3e}"
        d.train(); od.train()
        with torch.no_grad():
            line += f"   D train {rel(d(x, y), od(x, y)):.3e}"
    print(line, flush=True)
