"""Host-side placement for the data-parallel input path (one process per GPU).

With eight ranks each streaming its 2 GB batch per step from pinned host memory, the copies are bound by host DRAM
and the socket interconnect unless a rank's staging buffers live on the NUMA node its GPU hangs off. ``bind_to_gpu``
pins the calling process to the CPUs NVML reports as local to the device -- call it BEFORE allocating pinned
memory (pages are placed on first touch). The reference leaves this to SLURM (`run.sh:12-14`: one task per GPU,
16 CPUs per task)."""
from __future__ import annotations

import os

__all__ = ["bind_to_gpu"]


def bind_to_gpu(device_index: int):
    """-> sorted list of CPU ids the process is now bound to, or None when NVML / affinity control is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        # CUDA_VISIBLE_DEVICES remaps CUDA ordinals; resolve through the PCI bus id of the CUDA device
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev:02x}.0")
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
