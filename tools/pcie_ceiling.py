#!/usr/bin/env python3
"""Measured PCIe ceiling of the box for the e2e leg of bench.py: pinned-host <-> device copies of
4 GiB, each direction alone and both at once (two streams), GB/s and GiB/s per direction."""
import torch, time
n = 4 << 30
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if h2d:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
run(True, True, 1)
for name, a, b in (("H2D alone", True, False), ("D2H alone", False, True), ("both at once", True, True)):
    t = run(a, b)
    print(f"{name:14s} {n / t / 1e9:7.2f} GB/s = {n / t / 2**30:6.2f} GiB/s per direction")
