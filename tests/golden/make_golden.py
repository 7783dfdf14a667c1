"""Generate the committed golden vectors from the oracle (run here, on CPU):

    python tests/golden/make_golden.py

The reference itself cannot be imported (monai / lightning / torchio / nibabel are absent, SURVEY.md
section 8c), so the goldens pin the ORACLE restatement; weights come from the default torch init under
``torch.manual_seed(0)`` and are identified by a checksum stored next to the outputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import eval_oracle as E  # noqa: E402
from oracle import model_oracle as O  # noqa: E402


def state_checksum(module):
    tot = 0.0
    for k, v in sorted(module.state_dict().items()):
        tot += float(v.double().abs().sum()) * (1 + (sum(map(ord, k)) % 97) / 97.0)
    return tot


def synth_eval_volumes(seed=7, shape=(20, 24, 16), c=6):
    rng = np.random.default_rng(seed)
    tgt = rng.uniform(0.05, 1.0, size=shape + (c,)).astype(np.float32)
    pred = (tgt * rng.uniform(0.7, 1.3, size=tgt.shape)).astype(np.float32)
    tgt[0, 0, 0, :] = 0.0                      # |p-t|/0 -> inf -> zeroed by the reduction
    pred[1, 1, 1, 0] = 0.0
    tgt[1, 1, 1, 0] = 0.0                      # 0/0 -> NaN, kept (propagates into channel 0 sums)
    zz, yy, xx = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    ctr = [(s - 1) / 2 for s in shape]
    rad = 0.4 * min(shape)
    mask = (((zz - ctr[0]) ** 2 + (yy - ctr[1]) ** 2 + (xx - ctr[2]) ** 2) <= rad ** 2).astype(np.uint8)
    mask[1, 1, 1] = 1
    probseg = rng.dirichlet([1.0, 1.0, 1.0], size=shape).astype(np.float32)
    return pred, tgt, mask, probseg


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out = {}
    for mod in ("bssfp", "t1w"):
        cin = O.in_channels_of(mod)
        torch.manual_seed(0)
        g, d = O.Generator(mod), O.Discriminator(mod)
        out[f"{mod}_g_checksum"] = np.float64(state_checksum(g))
        out[f"{mod}_d_checksum"] = np.float64(state_checksum(d))
        torch.manual_seed(1234)
        x = torch.rand(1, cin, 32, 32, 32)
        y = torch.rand(1, 6, 32, 32, 32)
        g.eval()
        with torch.no_grad():
            yh = g(x)       # 32^3 is the smallest legal size (InstanceNorm needs > 1 voxel at 1/16 res)
        out[f"{mod}_g_eval_32"] = yh.numpy()
        d.train()
        with torch.no_grad():
            xb, yb = torch.cat([x, x.flip(2)], 0), torch.cat([y, y.flip(3)], 0)
            logits = d(xb, yb)
        out[f"{mod}_d_train_32_b2"] = logits.numpy()
        # losses of one training step before any optimiser update (dropout off, train-mode norms)
        torch.manual_seed(0)
        g, d = O.Generator(mod), O.Discriminator(mod)
        for m in g.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        g.train(); d.train()
        with torch.no_grad():
            gl, _ = O.gen_loss(g, d, xb, yb)
            dl = O.discr_loss(g, d, xb, yb)
        out[f"{mod}_gen_loss"] = np.float64(gl.item())
        out[f"{mod}_discr_loss"] = np.float64(dl.item())
    pred, tgt, mask, probseg = synth_eval_volumes()
    diff = E.rel_error_map(pred, tgt)
    errs, cleaned = E.roi_error_avg(diff, mask, probseg)
    out["eval_diff_rel"] = diff.astype(np.float32)
    out["eval_errs_rel"] = errs
    ang_p = (pred * 720.0 - 180.0).astype(np.float32)
    ang_t = (tgt * 360.0).astype(np.float32)
    diff_a = E.rel_error_map(ang_p[..., :1], ang_t[..., :1], kind="azimuth")
    errs_a, _ = E.roi_error_avg(diff_a, mask, probseg)
    out["eval_diff_ang"] = diff_a.astype(np.float32)
    out["eval_errs_ang"] = errs_a
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
