#!/usr/bin/env python
"""bench.py -- G+D train-step voxels/s of the sm_100a hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference          # the reference arithmetic (oracle) on the host cores

One "step" = one ``training_step`` of ref:src/model.py:259-281 (G phase + AdamW, D phase on a fresh G
forward + AdamW; L1 + adversarial losses, perceptual term out of scope) on one synthetic batch:
config[1] of BASELINE.json -- ``bssfp`` (24 input channels), batch 8 per GPU, 128^3 patches, bf16
tensor-core math with fp32 accumulate. Data-parallel over ranks (weak scaling, per-rank batch fixed),
gradient all-reduce on NCCL. Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "G+D train-step voxels/sec"
UNIT = "voxels/s"
# conv/deconv MACs x 2 per input voxel per train step, reference semantics
FLOP_PER_VOXEL = {"bssfp": 2369152, "pc-bssfp": 2369152, "t1w": 2311264, "dwi-tensor": 2311264}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm": p["hbm_gbs"], "src": "measured"}
    except Exception:
        return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                try:
                    pw.append(float(r[2]))
                except Exception:
                    pass
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        # median under load: ignore idle samples far below the busiest
        hot = [v for v in sm if v >= 0.5 * max(sm)] if sm else []
        hot_pw = [p for p, v in zip(pw, sm) if v >= 0.5 * max(sm)] if (sm and len(pw) == len(sm)) else pw
        return {"sm_mhz": statistics.median(hot) if hot else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": statistics.median(hot_pw) if hot_pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(steps, warmup, size=64, modality="bssfp"):
    """The reference arithmetic (oracle restatement of ref:src/model.py) on the host cores: full GAN
    step at batch 1, size^3 (a bounded sample of the workload). -> (voxels/s, ms/step, cores)."""
    from oracle import model_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    og, od = O.Generator(modality), O.Discriminator(modality)
    opt_g, opt_d = O.make_optimizers(og, od)
    cin = O.in_channels_of(modality)
    torch.manual_seed(1234)
    x, y = torch.rand(1, cin, size, size, size), torch.rand(1, 6, size, size, size)
    for _ in range(warmup):
        O.gan_step(og, od, opt_g, opt_d, x, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.gan_step(og, od, opt_g, opt_d, x, y)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return size ** 3 / dt, dt * 1e3, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    v, ms, cores = cpu_reference_run(steps, warmup, size=64, modality=args.modality)
    sample = f"full GAN step, batch 1 x 64^3 ({args.modality}), fp32 torch CPU, {steps} timed steps after {warmup} warm-up"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": f"full GAN step (U-Net G + PatchGAN D, L1+adversarial), {args.modality}, "
                               f"batch {args.batch} x {args.size}^3 per GPU", "sampled_as": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def dominant_kernel_roofline(dev, batch, size, peaks, iters=5):
    """Time the dominant launch in isolation: igemm_fwd_kernel on upcat_1.conv_0 (skip-concat
    [32 | 64] -> 32 channels, 3x3x3) at the bench shape; 29.9 % of the generator FLOPs."""
    import unet_bssfp_b200 as ub
    ops = ub.ops
    spec = ops.ConvSpec(0, 32, 32, 64)
    g = torch.Generator(device=dev).manual_seed(1)
    s0 = torch.randn((batch, size, size, size, 32), device=dev, generator=g).to(torch.bfloat16)
    s1 = torch.randn((batch, size, size, size, 64), device=dev, generator=g).to(torch.bfloat16)
    w = torch.randn((32, 96, 3, 3, 3), device=dev, generator=g) * 0.02
    wpk = ops.pack_conv_weights(spec, w, 0)
    bias = torch.zeros(32, device=dev)
    for _ in range(2):
        ops.conv_fwd(spec, s0, s1, wpk, bias, want_stats=True)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        ops.conv_fwd(spec, s0, s1, wpk, bias, want_stats=True)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = statistics.median(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    flops = 2.0 * batch * size ** 3 * 27 * 96 * 32
    ach = flops / (ms * 1e-3) / 1e12
    # DRAM traffic of this launch from the committed `ncu --set full` capture of the CTA-pair kernel
    # (profiles/r01w_march96_pair_ncu_summary.txt): dram__bytes_read.sum 3.267 GB + dram__bytes_write.sum 1.053 GB;
    # algorithmic bytes = both sources once + the output once = (32 + 64 + 32) ch * 2 B * 8 * 128^3 voxels = 4.295 GB
    traffic = 3.267318e9 + 1.052962e9 if (batch, size) == (8, 128) else None
    return {"bound": "tensor", "kernel": "igemm_march_kernel[upcat_1.conv_0 fwd, cat[32|64]->32 k3, 8x128^3]", "achieved": ach,
            "peak": peaks["bf16_burst"], "peak_src": peaks["src"] + " (burst: kernel timed alone)", "unit": "TFLOP/s",
            "frac": ach / peaks["bf16_burst"], "ms_per_launch": ms, "flops_per_launch": flops, "traffic": traffic,
            "traffic_unit": "bytes per launch (ncu dram read + write)", "algorithmic_bytes": 2.0 * batch * size ** 3 * 128}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="per-GPU batch (BASELINE config 2: 8)")
    ap.add_argument("--size", type=int, default=128, help="patch edge (BASELINE config 2: 128)")
    ap.add_argument("--modality", default="bssfp")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries ONE JSON line (rank 0): everything else that writes to fd 1 while the bench runs (NCCL's
    # version banner, library chatter) is sent to stderr; the line itself goes to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import unet_bssfp_b200 as ub
    from unet_bssfp_b200 import _lib, hostmem
    from unet_bssfp_b200.train_step import GanTrainer
    lib = _lib.load()
    peaks = load_peaks()
    # one process per GPU: stage the pinned input batch on the NUMA node of this rank's GPU (before allocating it)
    numa_cpus = hostmem.bind_to_gpu(local) if (world > 1 and os.environ.get("UB_BIND_NUMA", "1") != "0") else None

    torch.manual_seed(0)  # same initial weights on every rank (DDP would broadcast rank 0's)
    gen = ub.Generator(args.modality).to(dev)
    dis = ub.Discriminator(args.modality).to(dev)
    trainer = GanTrainer(gen, dis)
    cin = 24 if args.modality in ("bssfp", "pc-bssfp") else 6
    B, S = args.batch, args.size
    torch.manual_seed(1234 + rank)
    x_host = torch.rand(B, cin, S, S, S).pin_memory()
    y_host = torch.rand(B, 6, S, S, S).pin_memory()
    x, y = x_host.to(dev), y_host.to(dev)
    voxels_per_step = B * S ** 3 * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ("value") ----------------
    for _ in range(args.warmup):
        trainer.step(x, y)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.ub_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        g_loss, d_loss = trainer.step(x, y)
    e1.record()
    barrier()
    launches = lib.ub_launch_count() - n0
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = ms_total.item() / args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = voxels_per_step / (ms_step * 1e-3)

    # ---------------- end-to-end: host buffers, H2D + D2H inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream(device=dev)
        bufs = [(torch.empty_like(x), torch.empty_like(y)) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        loss_host = torch.zeros(2, 2).pin_memory()

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[i % 2])
                bufs[i % 2][0].copy_(x_host, non_blocking=True)
                bufs[i % 2][1].copy_(y_host, non_blocking=True)
                ready[i % 2].record(copy_stream)

        def e2e_loop(nsteps):
            for ev in done:
                ev.record()
            prefetch(0)
            for i in range(nsteps):
                if i + 1 < nsteps:
                    prefetch(i + 1)
                torch.cuda.current_stream().wait_event(ready[i % 2])
                gl, dl = trainer.step(*bufs[i % 2])
                done[i % 2].record()
                loss_host[i % 2, 0].copy_(gl, non_blocking=True)
                loss_host[i % 2, 1].copy_(dl, non_blocking=True)
            torch.cuda.synchronize()
            return float(loss_host[(nsteps - 1) % 2, 0])

        e2e_loop(1)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        e2e_loop(args.steps)
        e1.record()
        barrier()
        ms_e = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
        ms_e_step = ms_e.item() / args.steps
        e2e = {"value": voxels_per_step / (ms_e_step * 1e-3), "unit": UNIT, "ms_per_step": ms_e_step,
               "h2d_bytes_per_step": int((x_host.numel() + y_host.numel()) * 4), "d2h_bytes_per_step": 8,
               "how": "pinned host batch -> double-buffered H2D on a copy stream -> GanTrainer.step -> losses D2H",
               "numa_bound_cpus": len(numa_cpus) if numa_cpus else None}

    roof = cpu = None
    if rank == 0:
        roof = dominant_kernel_roofline(dev, B, S, peaks) if not args.no_roofline else {"bound": "tensor"}
        flop_step = FLOP_PER_VOXEL[args.modality] * B * S ** 3
        roof["step_tflops"] = flop_step / (ms_step * 1e-3) / 1e12
        roof["step_frac_of_sustained_peak"] = roof["step_tflops"] / peaks["bf16_sustained"]
        if not args.no_cpu_baseline and world == 1:   # reported at N = 1 only (torchrun pins OMP_NUM_THREADS=1)
            v, ms, cores = cpu_reference_run(steps=2, warmup=1, size=64, modality=args.modality)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": ms,
                   "sample": f"full GAN step, batch 1 x 64^3 ({args.modality}), fp32 torch CPU oracle, 2 timed steps after 1 warm-up"}
    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"full GAN step (U-Net G + PatchGAN D, L1+adversarial, 2x AdamW), {args.modality}, "
                                   f"batch {B} x {S}^3 per GPU", "global_batch": B * world, "patch": S,
                       "parallelism": f"dp{world}", "l2": "inputs larger than L2 (activations are GBs per step)",
                       "losses": [float(g_loss), float(d_loss)]},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
