import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
t = torch.empty(2 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(2 << 30, dtype=torch.uint8, device='cuda')
for _ in range(2): d.copy_(t, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4): d.copy_(t, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"pinned H2D: {4 * 2.147 / dt:.1f} GB/s")
