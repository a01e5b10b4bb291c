#!/usr/bin/env python3
"""Latency of ONE reference-API call as a function of its size, GPU library vs the reference on one host
core: where does the GPU start to win?  (document the minimum useful size.)
Every GPU call pays a fixed cost: two copies, the kernel launch and the 227 KB table fill of each CTA.

    python tools/small_calls.py
"""
import ctypes
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
uaes = importlib.import_module("micro-aes_b200")
shim = uaes.shim(128)
ref_path = os.path.join(ROOT, "oracle", "_ref", "libref128.so")
ref = ctypes.CDLL(ref_path) if os.path.exists(ref_path) else None
key, iv = bytes(range(16)), bytes(12)
print(f"{'bytes':>10s} {'GPU CTR us':>11s} {'GPU GCM us':>11s} {'GPU XTS us':>11s} {'CPU CTR us':>11s} {'CPU GCM us':>11s}   (host buffers, median of 30)")
for n in (16, 256, 4096, 65536, 1 << 20, 4 << 20, 16 << 20, 64 << 20, 256 << 20):
    src = ctypes.create_string_buffer(os.urandom(min(n, 4096)) * max(1, n // 4096), n)
    dst = ctypes.create_string_buffer(n + 16)

    def med(f, reps=30):
        f()
        ts = []
        for _ in range(reps if n <= (16 << 20) else 5):
            t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
        return sorted(ts)[len(ts) // 2] * 1e6

    g_ctr = med(lambda: shim.AES_CTR_encrypt(key, iv, src, n, dst))
    g_gcm = med(lambda: shim.AES_GCM_encrypt(key, iv, None, 0, src, n, dst))
    g_xts = med(lambda: shim.AES_XTS_encrypt(key + key, iv + bytes(4), src, n, dst))
    c_ctr = c_gcm = float("nan")
    if ref and n <= (16 << 20):
        c_ctr = med(lambda: ref.AES_CTR_encrypt(key, iv, src, ctypes.c_size_t(n), dst), 5)
        c_gcm = med(lambda: ref.AES_GCM_encrypt(key, iv, None, ctypes.c_size_t(0), src, ctypes.c_size_t(n), dst), 3)
    print(f"{n:10d} {g_ctr:11.1f} {g_gcm:11.1f} {g_xts:11.1f} {c_ctr:11.1f} {c_gcm:11.1f}", flush=True)
print("error latch:", uaes.core().uaes_last_error())
