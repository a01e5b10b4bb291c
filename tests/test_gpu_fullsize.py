"""Parity at BASELINE.json's full sizes, where the 30-60 MiB/s/core oracle cannot cover every byte.

  config 2  AES-128-CTR 1 GiB      every byte against the oracle (all host cores)
  config 4  AES-128-CTR 16 GiB     oracle windows (shard edges, the 2^32 counter carry, random),
  (+5)                             launch-geometry invariance (one call == 16 ranged calls) and the
                                   involution E(E(x)) == x, both as XOR-folds over all 16 GiB
  config 3  AES-256-XTS 16 GiB     oracle windows (sector edges, sector numbers across 2^32),
            512-byte sectors       encrypt -> decrypt round trip and piecewise == whole as XOR-folds
  config 4  AES-128-GCM 4 GiB+tag  ciphertext == the (already pinned) CTR stream over all 4 GiB;
                                   the tag against GHASH recomputed by the oracle in 16 parallel
                                   pieces recombined with powers of H; decrypt accepts it and
                                   rejects a flipped bit

All integer work: bit-exact, no tolerance.
"""
import ctypes
import importlib
import os
from concurrent.futures import ThreadPoolExecutor

import pytest

from util import Oracle, rnd

pytestmark = pytest.mark.gpu
GIB = 1 << 30
CORES = os.cpu_count() or 4


@pytest.fixture(scope="module")
def uaes():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return importlib.import_module("micro-aes_b200")


@pytest.fixture(scope="module")
def torch():
    return pytest.importorskip("torch")


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def host(t, a, b):
    return bytes(t[a:b].cpu().numpy())


def windows(nbytes, extra=(), w=65536, count=12, tag="w"):
    """window offsets: both ends, given points of interest, and seeded random ones"""
    offs = {0, nbytes - w}
    for e in extra:
        offs.add(max(0, min(nbytes - w, e - w // 2)) // 16 * 16)
    for i in range(count):
        offs.add(int.from_bytes(rnd(f"{tag}{i}", 8), "little") % (nbytes - w) // 4096 * 4096)
    return sorted(offs)


def test_ctr128_1gib_every_byte(uaes, orc, torch):
    n, seed = GIB, 0x5EED0002
    key, iv = rnd("c2-key", 16), rnd("c2-iv", 12)
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    dst = torch.empty(n, dtype=torch.uint8, device="cuda")
    uaes.fill_splitmix64(seed, 0, src, n // 8)
    uaes.ctr_crypt_range(128, key, iv, 0, src, n, dst)
    got = dst.cpu().numpy()
    piece = 4 << 20

    def check(off):
        pt = orc.splitmix(seed, off // 8, piece // 8)
        return bytes(got[off:off + piece]) == orc.ctr(key, iv, pt, first_block=off // 16)

    with ThreadPoolExecutor(CORES) as ex:
        assert all(ex.map(check, range(0, n, piece)))


def test_ctr128_16gib_windows_and_invariants(uaes, orc, torch):
    n, seed = 16 * GIB, 0x5EED0004
    key, iv = rnd("c4-key", 16), rnd("c4-iv", 12)
    first = 3 << 30                  # rank 3's shard of config 5: contains the 2^32 counter carry
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    dst = torch.empty(n, dtype=torch.uint8, device="cuda")
    uaes.fill_splitmix64(seed, 0, src, n // 8)
    uaes.ctr_crypt_range(128, key, iv, first, src, n, dst)
    carry_at = ((1 << 32) - 1 - first) * 16      # byte offset of the block whose counter is 2^32
    for off in windows(n, extra=(carry_at, carry_at - 65536, GIB, 8 * GIB)):
        pt = orc.splitmix(seed, off // 8, 65536 // 8)
        assert host(dst, off, off + 65536) == orc.ctr(key, iv, pt, first_block=first + off // 16), off
    whole = uaes.xor_fold64(dst, n // 8)
    # (1) launch geometry: 16 ranged calls of 1 GiB (different grid-to-data mapping and counter
    #     bases) must produce the same 16 GiB
    pieces = 0
    tmp = torch.empty(GIB, dtype=torch.uint8, device="cuda")
    for i in range(16):
        uaes.ctr_crypt_range(128, key, iv, first + i * (GIB // 16), src[i * GIB:], GIB, tmp)
        pieces ^= uaes.xor_fold64(tmp, GIB // 8)
        assert torch.equal(tmp[:4096], dst[i * GIB:i * GIB + 4096])
    assert pieces == whole
    # (2) CTR is an involution: a second pass, in place, gives the plaintext back
    uaes.ctr_crypt_range(128, key, iv, first, dst, n, dst)
    assert uaes.xor_fold64(dst, n // 8) == uaes.xor_fold64(src, n // 8)
    assert torch.equal(dst[-(1 << 20):], src[-(1 << 20):]) and torch.equal(dst[:1 << 20], src[:1 << 20])


def test_xts256_16gib_sectors(uaes, orc, torch):
    n, seed, sb = 16 * GIB, 0x5EED0003, 512
    keys = rnd("c3-keys", 64)
    first = (1 << 32) - (n // sb) // 2           # sector numbers straddle 2^32
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    dst = torch.empty(n, dtype=torch.uint8, device="cuda")
    uaes.fill_splitmix64(seed, 0, src, n // 8)
    uaes.xts_sectors(256, keys, first, sb, src, n, dst, True)
    mid = ((1 << 32) - first) * sb
    for off in windows(n, extra=(mid, mid - 65536, 5 * GIB), tag="x"):
        pt = orc.splitmix(seed, off // 8, 65536 // 8)
        want = orc.xts_sectors(keys, first + off // sb, sb, pt)[1]
        assert host(dst, off, off + 65536) == want, off
    whole = uaes.xor_fold64(dst, n // 8)
    # piecewise (4 calls with shifted first_sector) == whole
    acc = 0
    tmp = torch.empty(4 * GIB, dtype=torch.uint8, device="cuda")
    for i in range(4):
        uaes.xts_sectors(256, keys, first + i * (4 * GIB // sb), sb, src[i * 4 * GIB:], 4 * GIB, tmp, True)
        acc ^= uaes.xor_fold64(tmp, 4 * GIB // 8)
    assert acc == whole
    del tmp
    # decrypt (the inverse-cipher kernel) restores the plaintext, in place
    uaes.xts_sectors(256, keys, first, sb, dst, n, dst, False)
    assert uaes.xor_fold64(dst, n // 8) == uaes.xor_fold64(src, n // 8)
    assert torch.equal(dst[:1 << 20], src[:1 << 20]) and torch.equal(dst[-(1 << 20):], src[-(1 << 20):])
    # one 1 GiB data unit through the reference-shaped API (tweak jump-ahead across 2^26 blocks)
    tw = rnd("c3-tweak", 16)
    m = GIB + 16 * 5 + 9                          # ragged: ciphertext stealing at the end
    uaes.xts_unit(256, keys, tw, src, m, dst, True)
    head = orc.xts(keys, tw, orc.splitmix(seed, 0, (1 << 20) // 8))[1]
    assert host(dst, 0, 1 << 20) == head
    # far offsets against the oracle: a wrong entry of the jump-ahead ladder
    # x^(128 * 2^i) above bit 18 of the block index would survive a round trip but not this.  The oracle
    # walks the reference's own chain T_(j+1) = alpha * T_j (micro_aes.c:1030-1036) from T_0 to the window.
    nblk = m // 16
    for k in (1 << 20, (1 << 20) - 1, 1 << 24, (1 << 24) + 12345, (1 << 25) + (1 << 22) + 7, (1 << 26) - 4096, nblk - 8192):
        pt = orc.splitmix(seed, k * 2, 4096 * 2)
        assert host(dst, 16 * k, 16 * (k + 4096)) == orc.xts_range(keys, tw, k, pt)[1], k
    # the last blocks with the stolen pair (micro_aes.c:1037-1053): window [nblk - 64, end)
    k = nblk - 64
    tail = orc.splitmix(seed, k * 2, (m - 16 * k + 7) // 8)[:m - 16 * k]
    assert host(dst, 16 * k, m) == orc.xts_range(keys, tw, k, tail)[1]
    whole = uaes.xor_fold64(dst, GIB // 8)
    # the same unit as 3 ranges (the multi-GPU / staged decomposition, uaes_xts_crypt_range) == one call
    tmp = torch.zeros(GIB + 4096, dtype=torch.uint8, device="cuda")
    cut1, cut2 = 16 * ((1 << 24) + 1000), 16 * ((1 << 25) + (1 << 23))
    for a, b in ((cut2, m), (0, cut1), (cut1, cut2)):
        uaes.xts_crypt_range(256, keys, tw, a // 16, src[a:], b - a, tmp[a:], True)
    assert uaes.xor_fold64(tmp, GIB // 8) == whole and torch.equal(tmp[GIB - 4096:m], dst[GIB - 4096:m])
    assert torch.equal(tmp[cut1 - 4096:cut1 + 4096], dst[cut1 - 4096:cut1 + 4096])
    del tmp
    uaes.xts_unit(256, keys, tw, dst, m, dst, False)
    assert torch.equal(dst[:m], src[:m])


def test_gcm128_4gib_tag(uaes, orc, torch):
    n, seed = 4 * GIB, 0x5EED0005
    key, nonce, aad = rnd("c5-key", 16), rnd("c5-nonce", 12), rnd("c5-aad", 20)
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    dst = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
    uaes.fill_splitmix64(seed, 0, src, n // 8)
    uaes.gcm_encrypt(128, key, nonce, aad, src, n, dst)
    # ciphertext == CTR stream from J0 + 1 (CTR counter(k) = iv||1 + k, so first_block = 1)
    ref = torch.empty(n, dtype=torch.uint8, device="cuda")
    uaes.ctr_crypt_range(128, key, nonce, 1, src, n, ref)
    assert uaes.xor_fold64(ref, n // 8) == uaes.xor_fold64(dst, n // 8)
    assert torch.equal(ref, dst[:n])
    for off in windows(n, count=4, tag="g"):
        pt = orc.splitmix(seed, off // 8, 65536 // 8)
        assert host(dst, off, off + 65536) == orc.ctr(key, nonce, pt, first_block=1 + off // 16)
    del ref
    # tag: GHASH by the oracle over 16*k pieces in parallel, recombined with H^(piece blocks)
    ct = dst[:n].cpu().numpy()
    base = ct.ctypes.data
    H = orc.encrypt_block(key, bytes(16))
    piece = n // (4 * CORES) // 16 * 16                      # whole blocks; the last piece is shorter
    parts = [(o, min(piece, n - o)) for o in range(0, n, piece)]
    with ThreadPoolExecutor(CORES) as ex:
        zs = list(ex.map(lambda p: orc.ghash_absorb(H, base + p[0], p[1]), parts))
    state = orc.ghash_absorb(H, aad, len(aad))
    pows = {}
    for (o, ln), z in zip(parts, zs):
        hp = pows.setdefault(ln, orc.gf128_pow(H, ln // 16))
        state = bytes(a ^ b for a, b in zip(orc.gf128_mul(hp, state), z))
    lens = (len(aad) * 8).to_bytes(8, "big") + (n * 8).to_bytes(8, "big")
    state = orc.ghash_absorb(H, lens, 16, state)
    ej0 = orc.encrypt_block(key, nonce + b"\0\0\0\1")
    want_tag = bytes(a ^ b for a, b in zip(state, ej0))
    assert host(dst, n, n + 16) == want_tag
    # decrypt: accepts, restores; rejects a flipped ciphertext bit without touching the output
    out = torch.full((n,), 0xAA, dtype=torch.uint8, device="cuda")
    dst[n // 2 + 5] ^= 4
    assert uaes.gcm_decrypt(128, key, nonce, aad, dst, n, out) == 0x1A
    assert int(out[:4096].min()) == 0xAA and uaes.xor_fold64(out, n // 8) == 0
    dst[n // 2 + 5] ^= 4
    assert uaes.gcm_decrypt(128, key, nonce, aad, dst, n, out) == 0
    assert torch.equal(out, src)


def test_ctr128_4gib_pageable_host_buffer(uaes, orc, torch):
    """what a drop-in caller really passes: a malloc'd (pageable) 4 GiB buffer, in place (VERDICT r1,
    weak #4).  Every byte against the device-resident result (same kernels, no staging), oracle
    windows at both ends and around staging-chunk borders."""
    n, seed = 4 * GIB + 16 * 3 + 5, 0x5EED0007
    key, iv = rnd("pg-key", 16), rnd("pg-iv", 12)
    src = torch.empty(n + 11, dtype=torch.uint8, device="cuda")
    uaes.fill_splitmix64(seed, 0, src, (n + 11) // 8)
    import numpy as np
    hostbuf = np.empty(n + 64, dtype=np.uint8)                      # pageable
    hostbuf[:n] = src[:n].cpu().numpy()
    hostbuf[n:] = 0xCC
    uaes.ctr_crypt_range(128, key, iv, 0, hostbuf.ctypes.data, n, hostbuf.ctypes.data)
    assert uaes.core().uaes_last_error() == 0
    assert hostbuf[n:].tobytes() == b"\xcc" * 64
    dst = torch.empty(n + 11, dtype=torch.uint8, device="cuda")
    uaes.ctr_crypt_range(128, key, iv, 0, src, n, dst)
    got = torch.from_numpy(hostbuf[:n])
    step = 1 << 30
    for a in range(0, n, step):
        b = min(n, a + step)
        assert torch.equal(got[a:b].cuda(), dst[a:b]), a
    chunk = 64 << 20
    for off in windows(n - 5, extra=(chunk, 2 * chunk, 37 * chunk, n - 65536), count=4, tag="pg"):
        pt = orc.splitmix(seed, off // 8, 65536 // 8)
        assert hostbuf[off:off + 65536].tobytes() == orc.ctr(key, iv, pt, first_block=off // 16), off


def test_gcm128_4gib_host_buffer_pipelined(uaes, orc, torch):
    """BASELINE config 4 through the reference-facing call on a HOST buffer: the message runs through
    the chunk pipeline as 64 shards; the tag must equal the one-launch device result (itself checked
    against the oracle's GHASH in test_gcm128_4gib_tag)"""
    n, seed = 4 * GIB, 0x5EED0005
    key, nonce, aad = rnd("c5-key", 16), rnd("c5-nonce", 12), rnd("c5-aad", 20)
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    dst = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
    uaes.fill_splitmix64(seed, 0, src, n // 8)
    uaes.gcm_encrypt(128, key, nonce, aad, src, n, dst)
    h = torch.empty(n + 16, dtype=torch.uint8, pin_memory=True)
    h[:n].copy_(src)
    shim = uaes.shim(128)
    shim.AES_GCM_encrypt(key, nonce, aad, len(aad), ctypes.c_void_p(h.data_ptr()), n, ctypes.c_void_p(h.data_ptr()))
    assert uaes.core().uaes_last_error() == 0
    assert bytes(h[n:].numpy()) == host(dst, n, n + 16)
    for a in range(0, n, GIB):
        assert torch.equal(h[a:a + GIB].cuda(), dst[a:a + GIB]), a
    # decrypt in place: verified on the device first, then copied back
    rc = shim.AES_GCM_decrypt(key, nonce, aad, len(aad), ctypes.c_void_p(h.data_ptr()), n, ctypes.c_void_p(h.data_ptr()))
    assert ord(rc) == 0
    for a in range(0, n, GIB):
        assert torch.equal(h[a:a + GIB].cuda(), src[a:a + GIB]), a
    # a forged tag: M_AUTHENTICATION_ERROR and the caller's buffer keeps the ciphertext
    shim.AES_GCM_encrypt(key, nonce, aad, len(aad), ctypes.c_void_p(h.data_ptr()), n, ctypes.c_void_p(h.data_ptr()))
    h[n + 3] ^= 1
    rc = shim.AES_GCM_decrypt(key, nonce, aad, len(aad), ctypes.c_void_p(h.data_ptr()), n, ctypes.c_void_p(h.data_ptr()))
    assert ord(rc) == 0x1A
    assert torch.equal(h[:GIB].cuda(), dst[:GIB]) and torch.equal(h[n - GIB:n].cuda(), dst[n - GIB:n])
