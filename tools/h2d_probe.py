# ad-hoc probe: pinned host -> device copy bandwidth of this box (the ceiling of bench.py's e2e figure)
import time
import torch
x = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for n in (1 << 30, 1 << 28, 1 << 26):
    for _ in range(2):
        y[:n].copy_(x[:n], non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    reps = (1 << 31) // n
    for _ in range(reps):
        y[:n].copy_(x[:n], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"H2D {n >> 20} MiB chunks: {reps * n / dt / 1e9:.1f} GB/s")
