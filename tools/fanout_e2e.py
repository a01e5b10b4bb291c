#!/usr/bin/env python3
"""ONE process, ONE reference-API call on a host buffer, spread over 1 / 2 / 4 / 8 GPUs by the library
(uaes_set_devices): GiB/s end to end (H2D + kernel + D2H inside the call).  VERDICT r1 item 2.

    python tools/fanout_e2e.py [GiB]        (default 32)
"""
import ctypes
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import Oracle  # noqa: E402

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 32.0
n = int(gib * (1 << 30))
uaes = importlib.import_module("micro-aes_b200")
shim = uaes.shim(128)
orc = Oracle()
key, iv = bytes(range(16)), bytes(range(12))
ndev = torch.cuda.device_count()
print(f"GPUs {ndev}, host buffer {gib:g} GiB pinned")
h = torch.empty(n + 16, dtype=torch.uint8, pin_memory=True)
h[:n] = 0
hp = ctypes.c_void_p(h.data_ptr())
W = 1 << 16


def check_ctr():
    """the buffer was zero: it now holds the keystream; windows against the oracle"""
    ok = True
    for off in (0, n // 2 - W, n // 2, n - W, (n // 3) // 16 * 16):
        ok &= bytes(h[off:off + W].numpy()) == orc.ctr(key, iv, bytes(W), first_block=off // 16)
    return ok


for k in [x for x in (1, 2, 4, 8) if x <= ndev]:
    uaes.set_devices(k)
    h[:n] = 0
    shim.AES_CTR_encrypt(key, iv, hp, n, hp)                 # warm-up: contexts, staging chunks
    ok = check_ctr()
    t = time.perf_counter()
    shim.AES_CTR_encrypt(key, iv, hp, n, hp)
    shim.AES_CTR_encrypt(key, iv, hp, n, hp)
    dt = (time.perf_counter() - t) / 2
    g = min(n, 8 << 30)
    shim.AES_GCM_encrypt(key, iv, None, 0, hp, g, hp)
    t = time.perf_counter()
    shim.AES_GCM_encrypt(key, iv, None, 0, hp, g, hp)
    dg = time.perf_counter() - t
    print(f"{k} GPU(s): AES_CTR_encrypt {gib / dt:7.1f} GiB/s (oracle windows {'ok' if ok else 'MISMATCH'})   "
          f"AES_GCM_encrypt ({g / 2**30:g} GiB) {g / 2**30 / dg:7.1f} GiB/s   error latch {uaes.core().uaes_last_error()}", flush=True)
uaes.set_devices(0)
page = np.empty(min(n, 8 << 30), dtype=np.uint8)
page[:] = 0
pp = ctypes.c_void_p(page.ctypes.data)
shim.AES_CTR_encrypt(key, iv, pp, page.size, pp)
ok = page[:W].tobytes() == orc.ctr(key, iv, bytes(W)) and page[-W:].tobytes() == orc.ctr(key, iv, bytes(W), first_block=(page.size - W) // 16)
t = time.perf_counter()
shim.AES_CTR_encrypt(key, iv, pp, page.size, pp)
dt = time.perf_counter() - t
print(f"all GPUs, pageable {page.size / 2**30:g} GiB: AES_CTR_encrypt {page.size / 2**30 / dt:7.1f} GiB/s ({'ok' if ok else 'MISMATCH'})")
