"""Per-step wall time (CUDA events) of a long run with the caching allocator's counters: are slow steps caused by
device allocations (cudaMalloc / cudaFree synchronise)?   python tools/step_jitter.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_bssfp_b200 as ub
from unet_bssfp_b200.train_step import GanTrainer
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda:0")
torch.manual_seed(0)
tr = GanTrainer(ub.Generator("bssfp").to(dev), ub.Discriminator("bssfp").to(dev))
B, S = 8, 128
bs = [(torch.rand(B, 24, S, S, S, device=dev), torch.rand(B, 6, S, S, S, device=dev)) for _ in range(2)]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
stats = []
ev[0].record()
for i in range(steps):
    tr.step(*bs[i % 2])
    ev[i + 1].record()
    st = torch.cuda.memory_stats()
    stats.append((st["num_device_alloc"], st["num_device_free"], st["num_alloc_retries"], st["reserved_bytes.all.current"] / 2**30,
                  st["allocated_bytes.all.peak"] / 2**30))
torch.cuda.synchronize()
for i in range(steps):
    print(f"step {i:3d} {ev[i].elapsed_time(ev[i + 1]):8.2f} ms  device allocs {stats[i][0]:4d} frees {stats[i][1]:3d} retries {stats[i][2]} "
          f"reserved {stats[i][3]:6.1f} GiB peak allocated {stats[i][4]:6.1f} GiB")
