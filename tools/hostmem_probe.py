#!/usr/bin/env python3
"""What a drop-in caller's malloc'd (pageable) buffer costs on this box, measured so that the
pageable path of uaes_host.c is designed on numbers:

  1. cudaMemcpy of pageable memory, H2D and D2H (the driver's own staging);
  2. cudaHostRegister / cudaHostUnregister of touched pageable memory (pinning in place);
  3. plain memcpy into pinned memory with 1..16 threads (the library's bounce chunks + helper threads);
  4. the library itself: AES_CTR_encrypt in place on pageable vs pinned vs registered memory,
     for several helper-thread counts.

    python tools/hostmem_probe.py [GiB]
"""
import ctypes
import importlib
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GIB = 1 << 30


def best(f, reps=3):
    b = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        f()
        torch.cuda.synchronize()
        b = min(b, time.perf_counter() - t)
    return b


def main():
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    n = int(gib * GIB)
    print(f"host cores {os.cpu_count()}, buffer {gib:g} GiB")
    page = np.empty(n, dtype=np.uint8)
    page[:] = 1                                              # touched
    tp = torch.from_numpy(page)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    pin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    rt = torch.cuda.cudart()

    print(f"cudaMemcpy pageable H2D   {gib / best(lambda: dev.copy_(tp)):7.2f} GiB/s")
    print(f"cudaMemcpy pageable D2H   {gib / best(lambda: tp.copy_(dev)):7.2f} GiB/s")
    print(f"cudaMemcpy pinned   H2D   {gib / best(lambda: dev.copy_(pin, non_blocking=True)):7.2f} GiB/s")
    print(f"cudaMemcpy pinned   D2H   {gib / best(lambda: pin.copy_(dev, non_blocking=True)):7.2f} GiB/s")

    t = time.perf_counter()
    rc = rt.cudaHostRegister(page.ctypes.data, n, 0)
    t_reg = time.perf_counter() - t
    t = time.perf_counter()
    rt.cudaHostUnregister(page.ctypes.data)
    t_unreg = time.perf_counter() - t
    print(f"cudaHostRegister          {gib / t_reg:7.2f} GiB/s ({t_reg * 1e3:.0f} ms, rc {int(rc)}), unregister {gib / t_unreg:7.2f} GiB/s")
    t = time.perf_counter()
    for off in range(0, n, 64 << 20):
        rt.cudaHostRegister(page.ctypes.data + off, min(64 << 20, n - off), 0)
    t_reg = time.perf_counter() - t
    for off in range(0, n, 64 << 20):
        rt.cudaHostUnregister(page.ctypes.data + off)
    print(f"cudaHostRegister, 64 MiB pieces {gib / t_reg:7.2f} GiB/s")

    pin_np = pin.numpy()
    for threads in (1, 2, 4, 8, 16):
        per = n // threads

        def cp(i, to_pinned=True):
            a, b = i * per, n if i == threads - 1 else (i + 1) * per
            if to_pinned:
                np.copyto(pin_np[a:b], page[a:b])
            else:
                np.copyto(page[a:b], pin_np[a:b])

        def run(to_pinned):
            ts = [threading.Thread(target=cp, args=(i, to_pinned)) for i in range(threads)]
            [x.start() for x in ts]
            [x.join() for x in ts]
        print(f"memcpy {threads:2d} threads  pageable->pinned {gib / best(lambda: run(True)):7.2f} GiB/s   pinned->pageable {gib / best(lambda: run(False)):7.2f} GiB/s")

    uaes = importlib.import_module("micro-aes_b200")
    shim = uaes.shim(128)
    key, iv = bytes(range(16)), bytes(12)
    hp = ctypes.c_void_p(pin.data_ptr())
    shim.AES_CTR_encrypt(key, iv, hp, n, hp)
    print(f"AES_CTR_encrypt in place, pinned              {gib / best(lambda: shim.AES_CTR_encrypt(key, iv, hp, n, hp), 2):7.2f} GiB/s")
    pp = ctypes.c_void_p(page.ctypes.data)
    for threads in (1, 2, 4, 8, 12, 16):
        uaes.set_copy_threads(threads)
        shim.AES_CTR_encrypt(key, iv, pp, n, pp)
        print(f"AES_CTR_encrypt in place, pageable, {threads:2d} helpers {gib / best(lambda: shim.AES_CTR_encrypt(key, iv, pp, n, pp), 2):7.2f} GiB/s")
    for chunk in (8, 16, 32):
        uaes.set_staging(chunk << 20, 4)
        uaes.set_copy_threads(8)
        shim.AES_CTR_encrypt(key, iv, pp, n, pp)
        print(f"AES_CTR_encrypt pageable, 8 helpers, {chunk} MiB x 4 chunks {gib / best(lambda: shim.AES_CTR_encrypt(key, iv, pp, n, pp), 2):7.2f} GiB/s")
    uaes.set_staging(64 << 20, 3)
    t = time.perf_counter()
    uaes.core().uaes_host_register(page.ctypes.data, n)
    shim.AES_CTR_encrypt(key, iv, pp, n, pp)
    uaes.core().uaes_host_unregister(page.ctypes.data)
    dt = time.perf_counter() - t
    print(f"register + AES_CTR_encrypt + unregister        {gib / dt:7.2f} GiB/s")
    print("error latch:", uaes.core().uaes_last_error())


if __name__ == "__main__":
    main()
