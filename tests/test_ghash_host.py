"""The GCM kernel's multiply-by-constant (ghash_mul_table / ghash_table_entry in
micro-aes_b200/csrc/uaes_gf128.cuh: 16 independent table lookups, unreduced 248-bit accumulator,
one fold) compiled for the HOST and compared with the oracle's mulGF128 (micro_aes.c:476-493)."""
import ctypes
import os
import subprocess

import pytest

from util import ROOT, Oracle, rnd

HARNESS_DIR = os.path.join(ROOT, "tests", "host_harness")
SO = os.path.join(HARNESS_DIR, "libghash_host.so")
SRC = os.path.join(HARNESS_DIR, "ghash_host.cu")
CSRC = os.path.join(ROOT, "micro-aes_b200", "csrc")


@pytest.fixture(scope="module")
def harness():
    deps = [SRC, os.path.join(CSRC, "uaes_gf128.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["nvcc", "-O1", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
                               "-I", CSRC, "-o", SO, SRC])
    lib = ctypes.CDLL(SO)
    lib.ghash_host_mul.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p]
    return lib


def _mul(harness, C, ys, want_table=False):
    out = ctypes.create_string_buffer(len(ys))
    table = ctypes.create_string_buffer(256 * 16) if want_table else None
    assert harness.ghash_host_mul(C, ys, len(ys) // 16, out, table) == 0
    return out.raw, (table.raw if want_table else None)


def _edge_blocks():
    blocks = [bytes(16), b"\xff" * 16, b"\x80" + bytes(15), bytes(15) + b"\x01"]
    for bit in range(0, 128, 7):                       # single coefficients x^bit
        b = bytearray(16)
        b[bit // 8] = 0x80 >> (bit % 8)
        blocks.append(bytes(b))
    for i in range(16):                                # one full byte at each of the 16 positions
        b = bytearray(16)
        b[i] = 0xFF
        blocks.append(bytes(b))
    return blocks


@pytest.mark.parametrize("seed", range(6))
def test_product_matches_oracle(harness, seed):
    orc = Oracle()
    C = rnd(f"ghash-C-{seed}", 16)
    blocks = _edge_blocks() + [rnd(f"ghash-y-{seed}-{i}", 16) for i in range(200)]
    got, _ = _mul(harness, C, b"".join(blocks))
    for i, y in enumerate(blocks):
        assert got[16 * i:16 * i + 16] == orc.gf128_mul(y, C), (seed, i, y.hex())


@pytest.mark.parametrize("C", [bytes(16), b"\x80" + bytes(15), bytes(15) + b"\x01", b"\xff" * 16,
                               b"\xe1" + bytes(15)], ids=["zero", "one", "x127", "ones", "x128"])
def test_special_constants(harness, C):
    orc = Oracle()
    blocks = _edge_blocks() + [rnd(f"ghash-ys-{i}", 16) for i in range(32)]
    got, _ = _mul(harness, C, b"".join(blocks))
    for i, y in enumerate(blocks):
        assert got[16 * i:16 * i + 16] == orc.gf128_mul(y, C), (C.hex(), y.hex())


def test_table_entries(harness):
    """M[b] = b(x) * C with bit 7 of b the coefficient of x^0."""
    orc = Oracle()
    C = rnd("ghash-table-C", 16)
    _, table = _mul(harness, C, bytes(16), want_table=True)
    for b in range(256):
        assert table[16 * b:16 * b + 16] == orc.gf128_mul(bytes([b]) + bytes(15), C), b


def test_horner_chain_equals_ghash(harness):
    """y <- (y ^ X_i) * H over a message = the oracle's GHASH absorb loop (micro_aes.c:551-570)."""
    orc = Oracle()
    H = rnd("ghash-chain-H", 16)
    data = rnd("ghash-chain-data", 16 * 40)
    y = bytes(16)
    for i in range(0, len(data), 16):
        x = bytes(a ^ b for a, b in zip(y, data[i:i + 16]))
        y, _ = _mul(harness, H, x)
    assert y == orc.ghash_absorb(H, data, len(data))
