"""The bitsliced AES core of the CTR kernel's ALU co-runner (micro-aes_b200/csrc/uaes_bitslice.cuh,
generated S-box uaes_sbox_lut3.cuh) compiled for the HOST and compared with the oracle: one
1024-counter pass = 1024 keystream blocks, for all key sizes and several counter bases."""
import ctypes
import os
import struct
import subprocess
import sys

import pytest

from util import ROOT, Oracle, rnd

HARNESS_DIR = os.path.join(ROOT, "tests", "host_harness")
SO = os.path.join(HARNESS_DIR, "libbitslice_host.so")
SRC = os.path.join(HARNESS_DIR, "bitslice_host.cu")
CSRC = os.path.join(ROOT, "micro-aes_b200", "csrc")


@pytest.fixture(scope="module")
def harness():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("uaes_bitslice.cuh", "uaes_bitslice8.cuh", "uaes_sbox_lut3.cuh", "uaes_tables.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["nvcc", "-O1", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
                               "-I", CSRC, "-o", SO, SRC])
    lib = ctypes.CDLL(SO)
    lib.bs_host_pass.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p]
    return lib


@pytest.mark.parametrize("bits", [128, 192, 256])
@pytest.mark.parametrize("base", [0, 1024, 0x3FC00, 0xFFFFFC00, 0x123456789ABC00, 0xFFFFFFFFFFFC00])
def test_pass_matches_oracle(harness, bits, base):
    orc = Oracle()
    key = rnd(f"bs-key-{bits}", bits // 8)
    rk = orc.key_expansion(key)
    rounds = len(rk) // 16 - 1
    iv9 = rnd(f"bs-iv-{base}", 9)
    block0 = iv9 + base.to_bytes(7, "big")
    out = ctypes.create_string_buffer(1024 * 16)
    assert harness.bs_host_pass(rk, rounds, block0, out) == 0
    for idx in list(range(0, 1024, 37)) + [1, 31, 32, 255, 256, 1023]:
        blk = iv9 + ((base + idx) & ((1 << 56) - 1)).to_bytes(7, "big")
        assert out.raw[16 * idx:16 * idx + 16] == orc.encrypt_block(key, blk), (bits, hex(base), idx)


@pytest.mark.parametrize("bits", [128, 192, 256])
@pytest.mark.parametrize("base", [0, 256, 0xFF00, 0x3FF00, 0xFFFFFF00, 0x123456789ABC00, 0xFFFFFFFFFFFF00])
def test_narrow_form_group_matches_oracle(harness, bits, base):
    """uaes_bitslice8.cuh (8 blocks per thread, 32 state registers): one group = 256 counters, every block"""
    orc = Oracle()
    key = rnd(f"bs8-key-{bits}", bits // 8)
    rk = orc.key_expansion(key)
    rounds = len(rk) // 16 - 1
    iv9 = rnd(f"bs8-iv-{base}", 9)
    harness.bs8_host_group.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p]
    out = ctypes.create_string_buffer(256 * 16)
    assert harness.bs8_host_group(rk, rounds, iv9 + base.to_bytes(7, "big"), out) == 0
    want = b"".join(orc.encrypt_block(key, iv9 + (base + i).to_bytes(7, "big")) for i in range(256))
    assert out.raw == want, (bits, hex(base))


def test_whole_pass_every_block(harness):
    orc = Oracle()
    key = rnd("bs-key-all", 16)
    rk = orc.key_expansion(key)
    iv9 = rnd("bs-iv-all", 9)
    base = 0xABCDEF00 & ~1023
    out = ctypes.create_string_buffer(1024 * 16)
    assert harness.bs_host_pass(rk, 10, iv9 + base.to_bytes(7, "big"), out) == 0
    want = b"".join(orc.encrypt_block(key, iv9 + (base + i).to_bytes(7, "big")) for i in range(1024))
    assert out.raw == want


@pytest.mark.parametrize("bits", [128, 192, 256])
def test_general_32_block_path_matches_oracle(harness, bits):
    """the data-dependent form used by the XTS co-runner: 32 arbitrary blocks in, ECB out"""
    orc = Oracle()
    harness.bs_host_ecb32.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p]
    for trial in range(4):
        key, data = rnd(f"bs-ecb-k{bits}{trial}", bits // 8), rnd(f"bs-ecb-d{bits}{trial}", 512)
        rk = orc.key_expansion(key)
        out = ctypes.create_string_buffer(512)
        assert harness.bs_host_ecb32(rk, len(rk) // 16 - 1, data, out) == 0
        assert out.raw == orc.ecb_encrypt(key, data), (bits, trial)
        # and back through the bitsliced inverse cipher (derived inverse S-box, equivalent inverse schedule)
        harness.bs_host_ecb32_decrypt.argtypes = harness.bs_host_ecb32.argtypes
        back = ctypes.create_string_buffer(512)
        assert harness.bs_host_ecb32_decrypt(rk, len(rk) // 16 - 1, data, back) == 0
        assert (0, back.raw) == orc.ecb_decrypt(key, data), (bits, trial)


@pytest.mark.parametrize("bits", [128, 192, 256])
def test_narrow_general_form_matches_oracle(harness, bits):
    """uaes_bitslice8.cuh, data-dependent form: 8 arbitrary blocks in, ECB out, both directions"""
    orc = Oracle()
    harness.bs8_host_ecb8.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    for trial in range(4):
        key, data = rnd(f"bs8-ecb-k{bits}{trial}", bits // 8), rnd(f"bs8-ecb-d{bits}{trial}", 128)
        rk = orc.key_expansion(key)
        out = ctypes.create_string_buffer(128)
        assert harness.bs8_host_ecb8(rk, len(rk) // 16 - 1, data, out, 0) == 0
        assert out.raw == orc.ecb_encrypt(key, data), (bits, trial)
        back = ctypes.create_string_buffer(128)
        assert harness.bs8_host_ecb8(rk, len(rk) // 16 - 1, data, back, 1) == 0
        assert (0, back.raw) == orc.ecb_decrypt(key, data), (bits, trial)


def test_generated_sbox_is_current(harness, tmp_path):
    """the committed header is exactly what tools/gen_sbox_lut3.py emits (the generator checks both
    networks against the S-box tables on all 256 inputs before it writes anything)"""
    assert harness.bs_host_sbox_lut3_count() <= 80
    out = tmp_path / "uaes_sbox_lut3.cuh"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_sbox_lut3.py"), "4", str(out)])
    assert out.read_text() == open(os.path.join(CSRC, "uaes_sbox_lut3.cuh")).read()
