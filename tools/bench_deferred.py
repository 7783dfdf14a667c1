"""Deferred-activation consumers vs the materialising path at the bench shape (8 x 128^3):
conv forward / weight gradient on a deferred source 0 against (norm_act_fwd + the same conv on the tensor).
    python tools/bench_deferred.py [--batch 8] [--size 128]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import unet_bssfp_b200 as ub  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=128)
args = ap.parse_args()
ops = ub.ops
dev = "cuda"
N, S = args.batch, args.size
g = torch.Generator(device=dev).manual_seed(0)


def timed(fn, iters=5):
    fn(); fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))[iters // 2]


y = (torch.randn((N, S, S, S, 32), device=dev, generator=g) * 1.5).to(torch.float16)
scale = torch.rand((N, 32), device=dev, generator=g) + 0.5
shift = torch.randn((N, 32), device=dev, generator=g) * 0.3
for drop_p in (0.0, 0.05):
    lazy = ops.DeferredAct(y, scale, shift, 0.1, drop_p, 1234)
    t_na = timed(lambda: ops.norm_act_fwd(y, scale, shift, 0.1, drop_p, 1234))
    a, _ = ops.norm_act_fwd(y, scale, shift, 0.1, drop_p, 1234)
    print(f"dropout {drop_p}: norm_act_fwd {t_na:.3f} ms")
    for c0, c1 in ((32, 0), (32, 64)):
        spec = ops.ConvSpec(0, c0, 32, c1)
        s1 = torch.randn((N, S, S, S, c1), device=dev, generator=g).to(torch.bfloat16) if c1 else None
        wt = torch.randn((32, c0 + c1, 3, 3, 3), device=dev, generator=g) * 0.05
        wpk = ops.pack_conv_weights(spec, wt, 0)
        b = torch.zeros(32, device=dev)
        fl = 2.0 * N * S ** 3 * 27 * (c0 + c1) * 32
        t_m = timed(lambda: ops.conv_fwd(spec, a, s1, wpk, b, want_stats=True))
        t_d = timed(lambda: ops.conv_fwd(spec, lazy, s1, wpk, b, want_stats=True))
        lazy16 = ops.DeferredAct(y, scale, shift, 0.1, drop_p, 1234, f16_operand=True)
        wpk16 = ops.pack_conv_weights(spec, wt, ops.UB_PACK_F16_SRC0)
        t_h = timed(lambda: ops.conv_fwd(spec, lazy16, s1, wpk16, b, want_stats=True))
        print(f"  fwd {c0}+{c1}->32: materialised {t_m:.3f} ms ({fl / t_m / 1e9:.0f} TF/s) [+norm_act = {t_m + t_na:.3f}]   "
              f"deferred bf16 form {t_d:.3f} ms ({fl / t_d / 1e9:.0f} TF/s)   deferred fp16 form {t_h:.3f} ms ({fl / t_h / 1e9:.0f} TF/s)")
        dy = torch.randn((N, S, S, S, 32), device=dev, generator=g).to(torch.bfloat16)
        t_m = timed(lambda: ops.conv_wgrad(spec, a, s1, dy, tuple(wt.shape)))
        t_d = timed(lambda: ops.conv_wgrad(spec, lazy, s1, dy, tuple(wt.shape)))
        print(f"  wgrad {c0}+{c1}->32: materialised {t_m:.3f} ms   deferred {t_d:.3f} ms")
        del s1, dy
wt1 = torch.randn((6, 32, 1, 1, 1), device=dev, generator=g)
b1 = torch.zeros(6, device=dev)
print(f"output head: materialised {timed(lambda: ops.conv1x1_to_ncdhw(a, wt1, b1)):.3f} ms   deferred "
      f"{timed(lambda: ops.conv1x1_to_ncdhw(lazy, wt1, b1)):.3f} ms")
print(f"pool: full {timed(lambda: ops.norm_act_fwd(y, scale, shift, 0.1, 0.05, 1234, pool=True)):.3f} ms   pooled only "
      f"{timed(lambda: ops.norm_act_fwd(y, scale, shift, 0.1, 0.05, 1234, pool=True, materialize=False)):.3f} ms")
