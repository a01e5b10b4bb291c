#!/usr/bin/env python3
"""PCIe / host-memory ceiling of the BOX for N GPUs at once (e2e went 44 -> 62 GiB/s
from 1 to 8 GPUs and nobody knew whether the box or the library capped it).

One process, one thread per GPU, pinned host buffers, both directions at once on two streams per GPU:
pure cudaMemcpyAsync, no kernels, no library code.  Prints per-direction GiB/s summed over the GPUs for
N = 1, 2, 4, 8 (as many as are visible), plus the host-side NUMA / affinity facts.

    python tools/pcie_ceiling_multi.py [GiB per GPU per direction]
"""
import os
import sys
import threading
import time

import torch

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
n = int(gib * (1 << 30))
ndev = torch.cuda.device_count()
print(f"GPUs visible {ndev}, host cores {os.cpu_count()}, affinity {sorted(os.sched_getaffinity(0))[:4]}..{max(os.sched_getaffinity(0))}")
try:
    nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
    print("NUMA nodes:", sorted(nodes))
except OSError:
    pass
bufs = []
for d in range(ndev):
    torch.cuda.set_device(d)
    bufs.append((torch.empty(n, dtype=torch.uint8, pin_memory=True), torch.empty(n, dtype=torch.uint8, pin_memory=True),
                 torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}"), torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}"),
                 torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)))


def one(d, h2d, d2h, reps, out):
    torch.cuda.set_device(d)
    hin, hout, din, dout, s1, s2 = bufs[d]
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
    s1.synchronize(); s2.synchronize()
    out[d] = time.perf_counter()


def run(k, h2d, d2h, reps=4):
    out = [0.0] * ndev
    ts = [threading.Thread(target=one, args=(d, h2d, d2h, reps, out)) for d in range(k)]
    for d in range(k):
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    [t.start() for t in ts]
    [t.join() for t in ts]
    dt = max(out[:k]) - t0
    return k * reps * gib / dt


for k in [x for x in (1, 2, 4, 8) if x <= ndev]:
    run(k, True, True, 1)
    print(f"{k} GPU(s): H2D alone {run(k, True, False):7.1f}  D2H alone {run(k, False, True):7.1f}  "
          f"both at once {run(k, True, True):7.1f} GiB/s per direction (sum over GPUs)", flush=True)
