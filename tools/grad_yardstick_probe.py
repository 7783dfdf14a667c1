"""Print (ours, autocast yardstick) rel-L2 errors of the phase gradients against the reference goldens (the numbers
behind tests/test_reference_golden_gpu.py::test_phase_gradients_match_reference)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import tests.test_reference_golden_gpu as T
from unet_bssfp_b200.train_step import GanTrainer
rel = T.rel_l2
for rep in range(3):
    g, d = T._fresh("bssfp"); g.train(); d.train()
    tr = GanTrainer(g, d)
    xb, yb = T.synth_batch(24); xb, yb = xb.to(T.DEV), yb.to(T.DEV)
    for p in d.parameters(): p.requires_grad_(False)
    gl, _ = tr.gen_loss(xb, yb); gl.backward()
    gp = dict(g.named_parameters(remove_duplicate=False))
    yard = T._autocast_yardstick("bssfp", xb, yb, "gen")
    rows = []
    for k in [k for k in T.REF.files if k.startswith("bssfp_ggrad::")]:
        name, ref = k.split("::")[1], torch.from_numpy(T.REF[k])
        if name.endswith("conv.bias") and "final_conv" not in name and "deconv" not in name: continue
        rows.append((rel(gp[name].grad.cpu(), ref), rel(yard[name].grad.float().cpu(), ref), name))
    worst = max(r[0] / max(r[1], 1e-9) for r in rows)
    print(f"rep {rep}: G-phase tensors {len(rows)}; max ours/yardstick {worst:.2f}; median ours {np.median([r[0] for r in rows]):.3f} yard {np.median([r[1] for r in rows]):.3f}")
    for e, ey, nme in sorted(rows, key=lambda r: -r[0] / max(r[1], 1e-9))[:6]:
        print(f"   {nme:50s} ours {e:.3f} yard {ey:.3f}")
    print("   tensors with yardstick >= 0.5:", [(nme, round(e, 2), round(ey, 2)) for e, ey, nme in rows if ey >= 0.5][:12])
