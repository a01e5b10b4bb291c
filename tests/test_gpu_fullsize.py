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
