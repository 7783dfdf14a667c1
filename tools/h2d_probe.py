"""H2D bandwidth of the e2e input path: pinned fp32 batch -> device, alone."""
import torch, time
x_host = torch.rand(8, 24, 128, 128, 128).pin_memory()
x = torch.empty_like(x_host, device="cuda")
s = torch.cuda.Stream()
for chunks in (1, 8):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.cuda.stream(s):
            if chunks == 1:
                x.copy_(x_host, non_blocking=True)
            else:
                for i in range(chunks):
                    x[i].copy_(x_host[i], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"chunks={chunks}: {x_host.numel()*4/dt/1e9:.1f} GB/s ({dt*1e3:.1f} ms)")
