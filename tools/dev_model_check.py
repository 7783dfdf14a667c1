"""Dev script (GPU): print parity metrics of the product modules against the torch.nn oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_bssfp_b200 as ub
from oracle import model_oracle as O

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
mod = sys.argv[2] if len(sys.argv) > 2 else "bssfp"
cin = O.in_channels_of(mod)


def rl2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()


torch.manual_seed(0)
og = O.Generator(mod).to(dev)
od = O.Discriminator(mod).to(dev)
g = ub.Generator(mod).to(dev)
d = ub.Discriminator(mod).to(dev)
print(g.load_state_dict(og.state_dict()), d.load_state_dict(od.state_dict()))
torch.manual_seed(1234)
x = torch.rand(1, cin, S, S, S, device=dev)
y = torch.rand(1, 6, S, S, S, device=dev)

# ---- G eval forward
og.eval(); g.eval()
with torch.no_grad():
    ro = og(x); t0 = time.time(); po = g(x); torch.cuda.synchronize(); t1 = time.time()
    ro64 = og.double()(x.double()).float(); og.float()
print(f"G eval fwd: rel-L2 vs fp32 {rl2(po, ro):.3e}  vs fp64 {rl2(po, ro64):.3e}  maxabs {(po-ro).abs().max().item():.3e}  |y|max {ro.abs().max().item():.3f}  time {t1-t0:.3f}s")

# ---- G train-mode (dropout 0) forward + backward
for m in og.modules():
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
for m in g.modules():
    if isinstance(m, torch.nn.Dropout):
        m.p = 0.0
og.train(); g.train()
dY = torch.randn_like(ro)
ro = og(x); ro.backward(dY)
po = g(x); po.backward(dY)
torch.cuda.synchronize()
print(f"G train fwd: rel-L2 {rl2(po, ro):.3e}")
rows = []
for (n1, p1), (n2, p2) in zip(og.named_parameters(), g.named_parameters()):
    assert n1 == n2
    if p1.grad is None and p2.grad is None: continue
    if p1.grad is None or p2.grad is None:
        print("GRAD MISSING", n1, p1.grad is None, p2.grad is None); continue
    rows.append((n1, rl2(p2.grad, p1.grad), cos(p2.grad, p1.grad), p1.grad.abs().max().item()))
for r in rows:
    print(f"  {r[0]:55s} rel-L2 {r[1]:.3e} cos {r[2]:.5f} |g|max {r[3]:.3e}")
# yardstick: the oracle itself under torch bf16 autocast (cuDNN), same inputs
import copy
oa = copy.deepcopy(og)
for p_ in oa.parameters(): p_.grad = None
with torch.autocast("cuda", dtype=torch.bfloat16):
    ra = oa(x)
ra.float().backward(dY)
print(f"[yardstick torch-autocast-bf16] G train fwd rel-L2 {rl2(ra.float(), ro):.3e}")
ya = []
for (n1, p1), (n2, p2), (n3, p3) in zip(og.named_parameters(), g.named_parameters(), oa.named_parameters()):
    if p1.grad is None or p1.ndim < 2: continue
    ya.append((n1, rl2(p2.grad, p1.grad), rl2(p3.grad, p1.grad)))
import statistics
print("  per-tensor weight-grad rel-L2  ours: median %.3e max %.3e | autocast: median %.3e max %.3e" % (
    statistics.median(r[1] for r in ya), max(r[1] for r in ya), statistics.median(r[2] for r in ya), max(r[2] for r in ya)))
worst = max(ya, key=lambda r: r[1] / max(r[2], 1e-9))
print("  worst ratio ours/autocast: %s %.3e vs %.3e" % worst)
fa = torch.cat([p.grad.flatten() for p in oa.parameters() if p.grad is not None and p.ndim > 1])
fo = torch.cat([p.grad.flatten() for p in og.parameters() if p.grad is not None and p.ndim > 1])
print(f"[yardstick] whole weight-grad: rel-L2 {rl2(fa, fo):.3e} cos {cos(fa, fo):.5f}")
fp = torch.cat([p.grad.flatten() for p in g.parameters() if p.grad is not None and p.ndim > 1])
print(f"G whole weight-grad: rel-L2 {rl2(fp, fo):.3e} cos {cos(fp, fo):.5f}")
# running stats
hk = mod
print("head BN running_mean rel", rl2(g.blocks[hk].bn.running_mean, og.blocks[hk].bn.running_mean),
      "running_var rel", rl2(g.blocks[hk].bn.running_var, og.blocks[hk].bn.running_var))

# ---- D forward + backward (train mode, grads wrt params and y)
if S >= 64:
    od.train(); d.train()
    yo = y.clone().requires_grad_(True); yp = y.clone().requires_grad_(True)
    lo = od(x, yo); lp = d(x, yp)
    print(f"D fwd: rel-L2 {rl2(lp, lo):.3e} shape {tuple(lp.shape)}")
    dl = torch.randn_like(lo)
    lo.backward(dl); lp.backward(dl)
    torch.cuda.synchronize()
    print(f"D dy: rel-L2 {rl2(yp.grad, yo.grad):.3e} cos {cos(yp.grad, yo.grad):.5f}")
    for (n1, p1), (n2, p2) in zip(od.named_parameters(), d.named_parameters()):
        if p1.grad is None and p2.grad is None: continue
        if p1.grad is None or p2.grad is None:
            print("GRAD MISSING", n1, p1.grad is None, p2.grad is None); continue
        print(f"  {n1:40s} rel-L2 {rl2(p2.grad, p1.grad):.3e} cos {cos(p2.grad, p1.grad):.5f} |g|max {p1.grad.abs().max().item():.3e}")
    print("d2 BN running_var rel", rl2(d.d2.bn.running_var, od.d2.bn.running_var))

# ---- losses
a = torch.rand(2, 6, 16, 16, 16, device=dev, requires_grad=True); b = torch.rand_like(a)
l1 = ub.L1Loss()(a, b); l1.backward()
a2 = a.detach().clone().requires_grad_(True)
l2 = torch.nn.functional.l1_loss(a2, b); l2.backward()
print("L1", l1.item(), l2.item(), rl2(a.grad, a2.grad))
z = torch.randn(2, 1, 4, 4, 4, device=dev, requires_grad=True)
for tval in (0.0, 1.0):
    t = torch.full_like(z, tval)
    z.grad = None
    lb = ub.BCEWithLogitsLoss()(z, t); (lb * 3).backward()
    z2 = z.detach().clone().requires_grad_(True)
    lb2 = torch.nn.functional.binary_cross_entropy_with_logits(z2, t); (lb2 * 3).backward()
    print("BCE", tval, lb.item(), lb2.item(), rl2(z.grad, z2.grad))
