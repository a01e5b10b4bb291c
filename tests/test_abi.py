"""CPU-only checks of the drop-in boundary: the shared libraries load, export every symbol the
public headers declare, and refuse to compute without a CUDA device (no CPU fallback)."""
import ctypes
import importlib
import os
import re
import subprocess

import pytest

from util import ROOT

LIB = os.path.join(ROOT, "micro-aes_b200", "lib")


@pytest.fixture(scope="module")
def uaes():
    if not os.path.exists(os.path.join(LIB, "libuaes_b200.so")):
        import __graft_entry__
        __graft_entry__.build()
    return importlib.import_module("micro-aes_b200")


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:AES|uaes|GCM_SIV)_\w+)\s*\(", src)))


def exported(lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(LIB, lib)], text=True)
    return {line.split()[-1] for line in out.splitlines() if " T " in line}


def test_headers_and_exports_agree(uaes):
    ext = declared_functions("uaes_b200.h")
    assert len(ext) == 71 and set(ext) == set(uaes.UAES_ABI), sorted(set(ext) ^ set(uaes.UAES_ABI))
    assert set(ext) <= exported("libuaes_b200.so")
    ref = declared_functions("micro_aes.h")
    assert ref == sorted(uaes.MICRO_AES_ABI) and len(ref) == 20
    for bits in (128, 192, 256):
        assert set(ref) <= exported(f"libmicro_aes_{bits}.so")
    # the shim exports the reference's names and nothing else of ours
    assert not any(s.startswith("uaes_") for s in exported("libmicro_aes_128.so"))


def test_libraries_load_and_bind(uaes):
    uaes.core()
    for bits in (128, 192, 256):
        uaes.shim(bits)
    assert uaes.core().uaes_kernel_launches() >= 0


def test_product_does_not_link_the_oracle():
    for lib in os.listdir(LIB):
        if lib.endswith(".so"):
            deps = subprocess.check_output(["ldd", os.path.join(LIB, lib)], text=True)
            assert "oracle" not in deps and "libref" not in deps, (lib, deps)
    srcs = os.path.join(ROOT, "micro-aes_b200")
    for dirpath, _, files in os.walk(srcs):
        for f in files:
            if f.endswith((".c", ".cu", ".cuh", ".h", ".py")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "aes_oracle" not in text, f


def test_header_compiles_as_c90_and_cpp(tmp_path):
    prog = tmp_path / "t.c"
    prog.write_text('#include "micro_aes.h"\n#include "uaes_b200.h"\n'
                    "int main(void){ return (int)AES_KEYLENGTH - 16 + M_RESULT_SUCCESS; }\n")
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c90", "-pedantic", "-Wall", "-Wno-long-long", "-Werror", "-I", inc,
                           "-c", str(prog), "-o", str(tmp_path / "t.o")])
    subprocess.check_call(["g++", "-x", "c++", "-Wall", "-Werror", "-DAES___=256", "-I", inc,
                           "-c", str(prog), "-o", str(tmp_path / "t2.o")])


def test_no_device_is_a_loud_failure(uaes):
    """without a GPU nothing is computed: -1 (UAES_E_NO_DEVICE) and a latched message"""
    core = uaes.core()
    if core.uaes_device_count() > 0:
        pytest.skip("a CUDA device is present")
    out = ctypes.create_string_buffer(b"\xcc" * 32, 32)
    rc = core.uaes_ctr_crypt(128, bytes(16), bytes(12), b"x" * 32, 32, out)
    assert rc == -1 and out.raw == b"\xcc" * 32
    assert core.uaes_last_error() == -1 and b"no CPU fallback" in core.uaes_last_error_string()
    core.uaes_clear_error()
    with pytest.raises(uaes.UaesError):
        uaes.MicroAES(128).AES_CTR_encrypt(bytes(16), bytes(12), b"abc")
    # argument errors keep the reference's codes even without a device
    assert core.uaes_xts_encrypt(128, bytes(32), None, b"short", 5, out) == 1      # M_DATALENGTH_ERROR
    assert core.uaes_ctr_crypt(100, bytes(16), bytes(12), b"x", 1, out) == -3
    # XTS-192 is accepted like the reference built with AES___ = 192: only the device is missing
    assert core.uaes_xts_encrypt(192, bytes(48), None, bytes(32), 32, out) == -1
    assert core.uaes_gcm_encrypt_ex(128, bytes(16), bytes(12), 12, None, 0, b"x", 1, out, 17) == -3   # tag length
    assert core.uaes_ecb_encrypt_padded(128, bytes(16), b"x", 1, out, 3) == -3                       # padding mode
    assert core.uaes_cbc_decrypt_ex(128, bytes(16), bytes(16), bytes(17), 17, out, 0) == 1          # CTS = 0: whole blocks
    # settings work without a device
    assert core.uaes_set_devices(0) == 1 and core.uaes_get_devices() == 1
    core.uaes_set_burn(1); core.uaes_set_burn(0); core.uaes_trim(); core.uaes_shutdown()
