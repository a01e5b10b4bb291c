"""GPU parity tests for what round 2 added, all through the C ABI:

  * the reference's compile-time variants: PRESET_COUNTER, GCM nonces that are
    not 12 bytes and truncated tags (3 x 1000 cases of GcmEncryptExtIV*.rsp the default build skips),
    PKCS#7 / ISO 7816 padding, CTS = 0 -- through the run-time entry points AND through shim
    libraries compiled with the same macros (libmicro_aes_128_<variant>.so);
  * a range of one XTS data unit (uaes_xts_crypt_range) and XTS-192;
  * the staged host-buffer pipelines: single-unit XTS and GCM cut into chunks, pageable memory
    through the bounce chunks, calls spread over several device parts (fan-out), the per-call
    scratch pool under asynchronous multi-stream use, trim / shutdown / burn.

Bit-exact against the pinned oracle; no tolerance.
"""
import ctypes
import gzip
import importlib
import json
import os

import numpy as np
import pytest

from util import GOLDEN, Oracle, golden, rnd, sha256

pytestmark = pytest.mark.gpu
H = bytes.fromhex
MIB = 1 << 20


@pytest.fixture(scope="module")
def uaes():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    # more parts than devices are allowed so that the fan-out path runs on a one-GPU box
    os.environ.setdefault("UAES_FANOUT_OVERSUBSCRIBE", "1")
    mod = importlib.import_module("micro-aes_b200")
    assert mod.core().uaes_device_count() >= 1
    yield mod
    mod.set_devices(1)
    mod.set_staging(64 * MIB, 3)


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.fixture(autouse=True)
def _fresh_error_latch(uaes):
    """the latch keeps the most recent failure of the thread: earlier tests provoke failures on purpose"""
    uaes.core().uaes_clear_error()
    yield


@pytest.fixture(scope="module")
def torch():
    return pytest.importorskip("torch")


@pytest.fixture()
def small_chunks(uaes):
    """1 MiB staging chunks: a few MiB already run many chunks through the pipeline"""
    uaes.set_staging(1 * MIB, 3)
    yield
    uaes.set_devices(1)
    uaes.set_fanout_min(256 * MIB)
    uaes.set_staging(64 * MIB, 3)


def dev(torch, data, pad=16):
    t = torch.zeros(len(data) + pad, dtype=torch.uint8, device="cuda")
    if data:
        t[:len(data)] = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
    return t


def host(t, a, b):
    return bytes(t[a:b].cpu().numpy())


def hbuf(n, fill=0xCC):
    return ctypes.create_string_buffer(bytes([fill]) * max(n, 1), max(n, 1))


# ---------------------------------------------------------------- row b2: compile-time variants

@pytest.mark.parametrize("bits", [128, 192, 256])
def test_gcm_rsp_other_nonce_and_tag_lengths(uaes, bits):
    """GcmEncryptExtIV*.rsp groups with IVlen = 8 / 1024 bits or Taglen < 128 (micro_aes.c:1145-1149, 1178)"""
    with gzip.open(os.path.join(GOLDEN, "gcm_variants.json.gz")) as f:
        cases = json.load(f)[str(bits)]
    assert len(cases) == 1000
    for c in cases:
        key, iv, pt, aad, tag = H(c["key"]), H(c["iv"]), H(c["pt"]), H(c["aad"]), H(c["tag"])
        out = hbuf(len(pt) + 16)
        uaes.gcm_encrypt_ex(bits, key, iv, aad, pt, len(pt), out, taglen=len(tag))
        assert out.raw[:len(pt) + len(tag)] == H(c["ct"]) + tag, c
        assert out.raw[len(pt) + len(tag):len(pt) + 16] == b"\xcc" * (16 - len(tag))     # nothing beyond the tag
        back = hbuf(len(pt))
        assert uaes.gcm_decrypt_ex(bits, key, iv, aad, out.raw[:len(pt) + len(tag)], len(pt), back, taglen=len(tag)) == 0
        assert back.raw[:len(pt)] == pt


def test_gcm_variants_device_buffers_and_forgery(uaes, orc, torch):
    for ivlen, taglen, n, a in ((1, 16, 4096 + 5, 20), (128, 12, 1 << 20, 0), (60, 4, 33, 5000), (16, 15, 0, 7)):
        key, iv, aad, pt = rnd("gv-k", 16), rnd(f"gv-n{ivlen}", ivlen), rnd(f"gv-a{a}", a), rnd(f"gv-p{n}", n)
        want = orc.gcm_encrypt_ex(key, iv, aad, pt, taglen)
        src, dst = dev(torch, pt), dev(torch, b"", pad=n + 48)
        dst[:] = 0xCC
        uaes.gcm_encrypt_ex(128, key, iv, aad, src, n, dst, taglen)
        assert host(dst, 0, n + taglen) == want
        assert host(dst, n + taglen, n + 32) == b"\xcc" * (32 - taglen)
        back = dev(torch, b"", pad=n + 16)
        assert uaes.gcm_decrypt_ex(128, key, iv, aad, dst, n, back, taglen) == 0 and host(back, 0, n) == pt
        forged = bytearray(want); forged[-1] ^= 0x80
        out = hbuf(n)
        assert uaes.gcm_decrypt_ex(128, key, iv, aad, bytes(forged), n, out, taglen) == 0x1A
        assert out.raw[:n] == b"\xcc" * n                                   # untouched, micro_aes.c:1204-1208


def test_shim_variants_match_reference_builds(uaes, orc):
    """libmicro_aes_128_<variant>.so (same macros as oracle/_ref/libref128<variant>.so) on the recorded
    outputs of those reference builds"""
    v = golden("oracle_ref_variant_samples.json")
    pc = uaes.shim("128_pc")
    m = golden("main_c.json")
    key, iv16, pt = H(m["key_pool"]), H(m["iv16"]), H(m["plaintext"])
    out = hbuf(len(pt))
    pc.AES_CTR_encrypt(key[:16], iv16, pt, len(pt), out)                     # main.c:46-47 (PRESET_COUNTER vector)
    assert out.raw[:len(pt)] == H(m["ctr128_preset_counter"])
    for c in v["ctr_preset_counter"]:
        p = rnd(c["pt_tag"], c["n"])
        out = hbuf(c["n"])
        pc.AES_CTR_encrypt(H(c["key"]), H(c["counter0"]), p, c["n"], out)
        assert sha256(out.raw[:c["n"]]) == c["ct_sha256"], c
        back = hbuf(c["n"])
        pc.AES_CTR_decrypt(H(c["key"]), H(c["counter0"]), out.raw[:c["n"]], c["n"], back)
        assert back.raw[:c["n"]] == p
    civ = uaes.shim("128_civ8")                # CTR_IV_LENGTH = 8, CTR_START_VALUE = 0x01020304
    for c in v["ctr_iv8_start"]:
        p = rnd(c["pt_tag"], c["n"])
        out = hbuf(c["n"])
        civ.AES_CTR_encrypt(H(c["key"]), H(c["iv"]), p, c["n"], out)
        assert sha256(out.raw[:c["n"]]) == c["ct_sha256"], c
    for c in v["gcm_nonce"]:
        lib = uaes.shim(f"128_iv{c['noncelen']}")
        aad, p = rnd(c["aad_tag"], c["aadlen"]), rnd(c["pt_tag"], c["n"])
        out = hbuf(c["n"] + 16)
        lib.AES_GCM_encrypt(H(c["key"]), H(c["nonce"]), aad, len(aad), p, c["n"], out)
        assert sha256(out.raw[:c["n"]]) == c["ct_sha256"] and out.raw[c["n"]:c["n"] + 16].hex() == c["tag"], c
        back = hbuf(c["n"])
        assert ord(lib.AES_GCM_decrypt(H(c["key"]), H(c["nonce"]), aad, len(aad), out.raw[:c["n"] + 16], c["n"], back)) == 0
        assert back.raw[:c["n"]] == p
    t12 = uaes.shim("128_tag12")
    for c in v["gcm_tag12"]:
        aad, p = rnd(c["aad_tag"], c["aadlen"]), rnd(c["pt_tag"], c["n"])
        out = hbuf(c["n"] + 16, fill=0xEE)
        t12.AES_GCM_encrypt(H(c["key"]), H(c["nonce"]), aad, len(aad), p, c["n"], out)
        assert out.raw[c["n"]:c["n"] + 12].hex() == c["tag"] and out.raw[c["n"] + 12:c["n"] + 16] == b"\xee" * 4
        assert sha256(out.raw[:c["n"]]) == c["ct_sha256"]
        bad = bytearray(out.raw[:c["n"] + 12]); bad[-1] ^= 1
        assert ord(t12.AES_GCM_decrypt(H(c["key"]), H(c["nonce"]), aad, len(aad), bytes(bad), c["n"], hbuf(c["n"]))) == c["rc_forged"]
    at = uaes.shim("128_atag")                 # CCM_TAG_LEN = 8, EAX_TAG_LEN = 10, OCB_TAG_LEN = 12
    for c in v["aead_tags"]:
        enc = getattr(at, f"AES_{c['mode'].upper()}_encrypt")
        dec = getattr(at, f"AES_{c['mode'].upper()}_decrypt")
        aad, p, tl = rnd(c["aad_tag"], c["aadlen"]), rnd(c["pt_tag"], c["n"]), c["taglen"]
        out = hbuf(c["n"] + 16, fill=0xEE)
        enc(H(c["key"]), H(c["nonce"]), aad, len(aad), p, c["n"], out)
        assert sha256(out.raw[:c["n"]]) == c["ct_sha256"] and out.raw[c["n"]:c["n"] + tl].hex() == c["tag"], c
        assert out.raw[c["n"] + tl:] == b"\xee" * (16 - tl), c
        back = hbuf(c["n"])
        assert ord(dec(H(c["key"]), H(c["nonce"]), aad, len(aad), out.raw[:c["n"] + tl], c["n"], back)) == 0
        assert back.raw[:c["n"]] == p
        bad = bytearray(out.raw[:c["n"] + tl]); bad[-1] ^= 2
        assert ord(dec(H(c["key"]), H(c["nonce"]), aad, len(aad), bytes(bad), c["n"], hbuf(c["n"]))) == c["rc_forged"]
    for c in v["ecb_padding"]:
        lib = uaes.shim(f"128_pad{c['padding']}")
        p = rnd(c["pt_tag"], c["n"])
        m16 = (c["n"] // 16 + 1) * 16
        out = hbuf(m16 + 16, fill=0xEE)
        lib.AES_ECB_encrypt(H(c["key"]), p, c["n"], out)
        assert sha256(out.raw[:m16]) == c["ct_sha256"] and out.raw[m16:m16 + 16] == b"\xee" * 16, c
    c0 = uaes.shim("128_cts0")
    for c in v["cbc_nocts"]:
        ct = rnd(c["ct_tag"], c["n"])
        out = hbuf(c["n"])
        rc = ord(c0.AES_CBC_decrypt(H(c["key"]), H(c["iv"]), ct, c["n"], out))
        assert rc == c["rc"] and (rc or sha256(out.raw[:c["n"]]) == c["pt_sha256"]), c
    assert uaes.core().uaes_last_error() == 0


def test_padding_and_preset_counter_large_and_device(uaes, orc, torch):
    key = rnd("pp-k", 32)
    for mode in (1, 2):
        for n in (0, 16, 1 << 20, (1 << 20) + 7, 3 * MIB + 16):
            pt = rnd(f"pp-p{n}", n)
            want = orc.ecb_encrypt_padded(key, pt, mode)
            src, dst = dev(torch, pt), dev(torch, b"", pad=len(want) + 32)
            dst[:] = 0xCC
            uaes.ecb_encrypt_padded(256, key, src, n, dst, mode)
            assert host(dst, 0, len(want)) == want and host(dst, len(want), len(want) + 16) == b"\xcc" * 16
            out = hbuf(len(want) + 16)
            uaes.ecb_encrypt_padded(256, key, pt, n, out, mode)               # host buffers (staged)
            assert out.raw[:len(want)] == want and out.raw[len(want):] == b"\xcc" * 16
    # PRESET_COUNTER with the carries of micro_aes.c:421-427 and a counter range
    for ctr_hex in ("000102030405060708090a0bfffffffe", "a0a1a2a3a4a5a6a7a8fffffffffffffd", "00112233445566778899aabbccddeeff"):
        ctr, pt = H(ctr_hex), rnd("pp-c", 5 * MIB + 3)
        want = orc.ctr_block(key[:16], ctr, pt)
        src, dst = dev(torch, pt), dev(torch, b"", pad=len(pt) + 16)
        uaes.ctr_crypt_block(128, key[:16], ctr, 0, src, len(pt), dst)
        assert host(dst, 0, len(pt)) == want
        off = 16 * 70001
        uaes.ctr_crypt_block(128, key[:16], ctr, 70001, src[off:], len(pt) - off, dst)
        assert host(dst, 0, len(pt) - off) == orc.ctr_block(key[:16], ctr, pt[off:], first_block=70001)


def test_ccm_eax_ocb_tag_lengths(uaes, orc, torch):
    """CCM_TAG_LEN / EAX_TAG_LEN / OCB_TAG_LEN as run-time arguments: every legal length, device buffers for OCB"""
    key = rnd("tl-k", 16)
    for mode, nl, lens in (("ccm", 11, (4, 6, 8, 10, 12, 14, 16)), ("eax", 16, (1, 5, 8, 15, 16)), ("ocb", 12, (1, 4, 8, 12, 15, 16))):
        for tl in lens:
            for n, alen in ((0, 3), (57, 31), (16 * 300 + 5, 0)):
                nonce, aad, pt = rnd(f"tl-n{mode}", nl), rnd(f"tl-a{alen}", alen), rnd(f"tl-p{n}", n)
                want = orc.aead_ex(mode, key, nonce, aad, pt, tl)
                out = hbuf(n + 16)
                uaes.aead_ex(mode, 128, key, nonce, aad, pt, n, out, tl)
                assert out.raw[:n + tl] == want and out.raw[n + tl:] == b"\xcc" * (16 - tl), (mode, tl, n)
                back = hbuf(n)
                assert uaes.aead_ex(mode, 128, key, nonce, aad, want, n, back, tl, encrypt=False) == 0 and back.raw[:n] == pt
                bad = bytearray(want); bad[-1] ^= 0x10
                assert uaes.aead_ex(mode, 128, key, nonce, aad, bytes(bad), n, hbuf(n), tl, encrypt=False) == 0x1A
    nonce, pt = rnd("tl-on", 12), rnd("tl-op", 3 * MIB + 9)
    src, dst = dev(torch, pt), dev(torch, b"", pad=len(pt) + 32)
    dst[:] = 0xCC
    uaes.aead_ex("ocb", 128, key, nonce, b"hdr", src, len(pt), dst, 12)
    assert host(dst, 0, len(pt) + 12) == orc.aead_ex("ocb", key, nonce, b"hdr", pt, 12) and host(dst, len(pt) + 12, len(pt) + 16) == b"\xcc" * 4
    # CCM tags are even and at least 4 bytes (micro_aes.h:105); anything else is an argument error
    assert uaes.core().uaes_ccm_encrypt_ex(128, key, rnd("x", 11), None, 0, pt[:10], 10, hbuf(26), 7) == -3
    assert uaes.core().uaes_ocb_encrypt_ex(128, key, nonce, None, 0, pt[:10], 10, hbuf(26), 17) == -3


def test_cbc_without_cts(uaes, orc, torch):
    key, iv = rnd("nc-k", 24), rnd("nc-i", 16)
    for n in (0, 16, 32, 16 * 1000, 2 * MIB):
        ct = rnd(f"nc-c{n}", n)
        want = orc.cbc_nocts(key, iv, ct)
        out = hbuf(n)
        assert uaes.cbc_decrypt_ex(192, key, iv, ct, n, out, cts=False) == want[0] == 0
        assert out.raw[:n] == want[1]
        if n:
            src, dst = dev(torch, ct), dev(torch, b"", pad=n + 16)
            assert uaes.cbc_decrypt_ex(192, key, iv, src, n, dst, cts=False) == 0 and host(dst, 0, n) == want[1]
    assert uaes.cbc_decrypt_ex(192, key, iv, bytes(17), 17, hbuf(17), cts=False) == 1     # M_DATALENGTH_ERROR
    # cts = True stays the CS3 path
    ct = rnd("nc-cs3", 100)
    out = hbuf(100)
    assert uaes.cbc_decrypt_ex(192, key, iv, ct, 100, out, cts=True) == 0 and out.raw[:100] == orc.cbc(key, iv, ct)[1]


@pytest.mark.parametrize("narrow", [1, 0])
@pytest.mark.parametrize("bits", [128, 256])
def test_gcm_bitsliced_corunner(uaes, orc, torch, bits, narrow, monkeypatch):
    """gcm_bulk_hybrid_kernel forced on for small messages: the last part of the message is encrypted by
    bitsliced warps that also run their share of the GHASH; every split, ragged ends, AAD, shards whose
    first block is not a multiple of 1024, both hash directions (encrypt / decrypting shard)"""
    # narrow = 1: gcm_bulk_hybrid8_kernel (8 blocks per bitsliced thread, uaes_bitslice8.cuh), 0: the wide form
    monkeypatch.setenv("UAES_GCM_NARROW", str(narrow))
    key, nonce = rnd(f"gb-k{bits}", bits // 8), rnd("gb-n", 12)
    try:
        for share, n, alen in ((200, 16 * 9000 + 3, 20), (512, 16 * 70001, 0), (900, (1 << 21) + 9, 4500), (100, 16 * 4096, 7),
                               (1000, 16 * 5000 + 15, 1), (300, 3 * MIB, 0)):
            uaes.ctr_tuning(-1, share, 0)
            aad, pt = rnd(f"gb-a{alen}", alen), rnd(f"gb-p{bits}{n}", n)
            want = orc.gcm_encrypt(key, nonce, aad, pt)
            src, dst = dev(torch, pt), dev(torch, b"", pad=n + 32)
            uaes.gcm_encrypt(bits, key, nonce, aad, src, n, dst)
            assert host(dst, n, n + 16) == want[n:], (share, n, alen)
            assert host(dst, 0, n) == want[:n], (share, n, alen)
            back = dev(torch, b"", pad=n + 16)
            assert uaes.gcm_decrypt(bits, key, nonce, aad, dst, n, back) == 0 and host(back, 0, n) == pt
            # as two shards (the second starts at a block that is not a multiple of 1024), encrypting and
            # decrypting (MODE 2: CTR + GHASH of the input)
            cut = (n // 3) // 16 * 16 + 16 * 33                 # a block that is not a multiple of 1024
            for decrypt, data, ref in ((False, pt, want[:n]), (True, want[:n], pt)):
                s_in, s_out = dev(torch, data), dev(torch, b"", pad=n + 16)
                z0 = uaes.gcm_shard(bits, key, nonce, 0, s_in, cut, s_out, decrypt=decrypt)
                z1 = uaes.gcm_shard(bits, key, nonce, cut // 16, s_in[cut:], n - cut, s_out[cut:], decrypt=decrypt)
                assert host(s_out, 0, n) == ref, (share, n, decrypt)
                tag = uaes.gcm_combine(bits, key, nonce, aad, [z0, z1], [(n + 15) // 16 - cut // 16, 0], n)
                assert tag == want[n:], (share, n, decrypt)
    finally:
        uaes.ctr_tuning(388, 195, 1 << 23)


# ---------------------------------------------------------------- XTS: ranges of a unit, XTS-192

def test_xts_192(uaes, orc, torch):
    """the reference built with AES___ = 192 runs XTS with two 24-byte keys"""
    keys, tw = rnd("x192-k", 48), rnd("x192-t", 16)
    for n in (16, 17, 57, 4096 + 9, 2 * MIB + 1):
        pt = rnd(f"x192-p{n}", n)
        want = orc.xts(keys, tw, pt)
        assert uaes.MicroAES(192).AES_XTS_encrypt(keys, tw, pt) == want
        assert uaes.MicroAES(192).AES_XTS_decrypt(keys, tw, want[1]) == (0, pt)
    pt = rnd("x192-s", 512 * 100)
    src, dst = dev(torch, pt), dev(torch, b"", pad=len(pt))
    uaes.xts_sectors(192, keys, (1 << 32) - 7, 512, src, len(pt), dst, True)
    assert host(dst, 0, len(pt)) == orc.xts_sectors(keys, (1 << 32) - 7, 512, pt)[1]


@pytest.mark.parametrize("bits", [128, 256])
def test_xts_ranges_equal_the_unit(uaes, orc, torch, bits):
    """cut one data unit at block boundaries, run the ranges separately (any order, device or host
    memory): the concatenation is AES_XTS_encrypt of the whole, stealing included"""
    keys, tw = rnd(f"xr-k{bits}", bits // 4), rnd("xr-t", 16)
    n = 16 * 200000 + 11
    pt = rnd(f"xr-p{bits}", n)
    rc, want = orc.xts(keys, tw, pt)
    assert rc == 0
    cuts = [0, 16, 16 * 1023, 16 * 1024, 16 * 77777, 16 * 199999, n]
    src = dev(torch, pt)
    got = bytearray(n)
    for a, b in reversed(list(zip(cuts, cuts[1:]))):
        dst = dev(torch, b"", pad=b - a + 16)
        tmp = dev(torch, pt[a:b])                                            # 16-byte aligned start
        uaes.xts_crypt_range(bits, keys, tw, a // 16, tmp, b - a, dst, True)
        got[a:b] = host(dst, 0, b - a)
        out = hbuf(b - a)
        uaes.xts_crypt_range(bits, keys, tw, a // 16, pt[a:b], b - a, out, True)   # host memory
        assert out.raw[:b - a] == bytes(got[a:b])
    assert bytes(got) == want
    a = cuts[4]
    out = hbuf(n - a)
    uaes.xts_crypt_range(bits, keys, tw, a // 16, want[a:], n - a, out, False)
    assert out.raw[:n - a] == pt[a:]
    del src


# ---------------------------------------------------------------- staged pipelines

def test_single_unit_xts_pipelined_host(uaes, orc, small_chunks):
    """one AES_XTS_encrypt on a host buffer of many staging chunks (was: whole-buffer staging)"""
    keys, tw = rnd("xp-k", 64), rnd("xp-t", 16)
    for n in (5 * MIB + 9, 3 * MIB, MIB + 16, MIB + 17, 2 * MIB + 31, MIB - 3):
        pt = rnd(f"xp-p{n}", n)
        want = orc.xts(keys, tw, pt)
        assert uaes.MicroAES(256).AES_XTS_encrypt(keys, tw, pt) == want, n
        assert uaes.MicroAES(256).AES_XTS_decrypt(keys, tw, want[1]) == (0, pt), n


def test_gcm_pipelined_host(uaes, orc, small_chunks):
    """AES_GCM_* on host buffers of many chunks: one shard per chunk, contributions folded at the end"""
    a = uaes.MicroAES(128)
    key, nonce = rnd("gp-k", 16), rnd("gp-n", 12)
    for n, alen in ((5 * MIB + 9, 31), (3 * MIB, 0), (MIB + 1, 5000), (8 * MIB + 15, 70000), (12 * MIB, 3)):
        aad, pt = rnd(f"gp-a{alen}", alen), rnd(f"gp-p{n}", n)
        want = orc.gcm_encrypt(key, nonce, aad, pt)
        got = a.AES_GCM_encrypt(key, nonce, aad, pt)
        assert got[-16:] == want[-16:], (n, alen)
        assert got == want
        assert a.AES_GCM_decrypt(key, nonce, aad, want) == (0, pt)
        bad = bytearray(want); bad[n // 2] ^= 4
        rc, out = a.AES_GCM_decrypt(key, nonce, aad, bytes(bad))
        assert rc == 0x1A and out == b"\xcc" * n                            # untouched on failure
    # in place on one host buffer, AES-256, odd nonce length and short tag
    key = rnd("gp-k256", 32)
    n = 6 * MIB + 5
    pt, iv = rnd("gp-ip", n), rnd("gp-iv", 40)
    buf = ctypes.create_string_buffer(pt + bytes(16), n + 16)
    uaes.gcm_encrypt_ex(256, key, iv, b"hdr", buf, n, buf, taglen=13)
    assert buf.raw[:n + 13] == orc.gcm_encrypt_ex(key, iv, b"hdr", pt, 13)
    assert uaes.gcm_decrypt_ex(256, key, iv, b"hdr", buf, n, buf, taglen=13) == 0 and buf.raw[:n] == pt


def test_pageable_and_pinned_host_buffers(uaes, orc, torch, small_chunks):
    """pageable memory goes through the pinned bounce chunks and helper threads; pinned memory is
    DMA'd directly; mixed in/out classes; in place"""
    key, iv = rnd("pg-k", 16), rnd("pg-i", 12)
    n = 9 * MIB + 5
    pt = rnd("pg-p", n)
    want = orc.ctr(key, iv, pt)
    for threads in (1, 3):
        uaes.set_copy_threads(threads)
        pg_in = np.frombuffer(pt, dtype=np.uint8).copy()                     # pageable
        pg_out = np.zeros(n + 8, dtype=np.uint8)
        uaes.ctr_crypt_range(128, key, iv, 0, pg_in.ctypes.data, n, pg_out.ctypes.data)
        assert pg_out[:n].tobytes() == want and pg_out[n:].tobytes() == bytes(8)
        uaes.ctr_crypt_range(128, key, iv, 0, pg_in.ctypes.data, n, pg_in.ctypes.data)     # in place
        assert pg_in.tobytes() == want
    uaes.set_copy_threads(4)
    pin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    pin.copy_(torch.from_numpy(np.frombuffer(pt, dtype=np.uint8).copy()))
    pg_out = np.zeros(n, dtype=np.uint8)
    uaes.ctr_crypt_range(128, key, iv, 0, pin.data_ptr(), n, pg_out.ctypes.data)          # pinned -> pageable
    assert pg_out.tobytes() == want
    pin2 = torch.zeros(n, dtype=torch.uint8, pin_memory=True)
    uaes.ctr_crypt_range(128, key, iv, 0, pg_out.ctypes.data, n, pin2.data_ptr())         # pageable -> pinned
    assert bytes(pin2.numpy()) == pt
    d = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
    uaes.ctr_crypt_range(128, key, iv, 0, pg_out.ctypes.data, n, d)                       # pageable -> device
    assert host(d, 0, n) == pt
    # registered memory counts as pinned
    reg = np.frombuffer(pt, dtype=np.uint8).copy()
    assert uaes.core().uaes_host_register(reg.ctypes.data, n) == 0
    uaes.ctr_crypt_range(128, key, iv, 0, reg.ctypes.data, n, reg.ctypes.data)
    assert uaes.core().uaes_host_unregister(reg.ctypes.data) == 0
    assert reg.tobytes() == want
    # GCM on pageable memory with the DEFAULT staging geometry: 8 MiB ring pieces, one shard each
    uaes.set_staging(64 * MIB, 3)
    gp = rnd("pg-g", 1000) * 12600                                      # ~12 MiB
    gbuf = np.frombuffer(gp + bytes(16), dtype=np.uint8).copy()
    uaes.gcm_encrypt(128, key, iv, b"aad", gbuf.ctypes.data, len(gp), gbuf.ctypes.data)
    assert gbuf.tobytes() == orc.gcm_encrypt(key, iv, b"aad", gp)
    assert uaes.gcm_decrypt(128, key, iv, b"aad", gbuf.ctypes.data, len(gp), gbuf.ctypes.data) == 0
    assert gbuf[:len(gp)].tobytes() == gp
    uaes.set_staging(1 * MIB, 3)
    # the other staged modes on pageable memory
    keys = rnd("pg-x", 64)
    sec = np.frombuffer(rnd("pg-s", 4 * MIB), dtype=np.uint8).copy()
    out = np.zeros_like(sec)
    uaes.xts_sectors(256, keys, 77, 4096, sec.ctypes.data, sec.size, out.ctypes.data, True)
    assert out.tobytes() == orc.xts_sectors(keys, 77, 4096, sec.tobytes())[1]
    e = np.zeros(3 * MIB + 16, dtype=np.uint8)
    uaes.ecb(128, key, pg_in.ctypes.data, 3 * MIB + 7, e.ctypes.data, True)
    assert e.tobytes() == orc.ecb_encrypt(key, want[:3 * MIB + 7])


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_fanout_over_device_parts(uaes, orc, torch, small_chunks, parts):
    """a host-buffer call cut into `parts` device parts, one host thread each (on a one-GPU box the parts
    share the device: same code path, no speed-up); every mode that shards"""
    have = uaes.set_devices(parts)
    assert have == parts or have == uaes.core().uaes_device_count()
    uaes.set_fanout_min(MIB)
    key, iv = rnd("fo-k", 16), rnd("fo-i", 12)
    n = 11 * MIB + 13
    pt = rnd("fo-p", n)
    out = hbuf(n + 16)
    uaes.ctr_crypt_range(128, key, iv, (1 << 32) - 100000, pt, n, out)       # crosses the 2^32 carry inside a part
    assert out.raw[:n] == orc.ctr(key, iv, pt, first_block=(1 << 32) - 100000) and out.raw[n:] == b"\xcc" * 16
    e = hbuf(n + 16)
    uaes.ecb(128, key, pt, n, e, True)
    assert e.raw[:(n + 15) // 16 * 16] == orc.ecb_encrypt(key, pt)
    d = hbuf(n)
    assert uaes.ecb(128, key, e.raw[:n - 13], n - 13, d, False) == 0 and d.raw[:n - 13] == pt[:n - 13]
    keys, tw = rnd("fo-x", 64), rnd("fo-t", 16)
    assert uaes.MicroAES(256).AES_XTS_encrypt(keys, tw, pt) == orc.xts(keys, tw, pt)
    x = hbuf(8 * MIB)
    uaes.xts_sectors(256, keys, (1 << 32) - 3, 512, pt[:8 * MIB], 8 * MIB, x, True)
    assert x.raw[:8 * MIB] == orc.xts_sectors(keys, (1 << 32) - 3, 512, pt[:8 * MIB])[1]
    aad = rnd("fo-a", 4500)
    a = uaes.MicroAES(128)
    want = orc.gcm_encrypt(key, iv, aad, pt)
    assert a.AES_GCM_encrypt(key, iv, aad, pt) == want
    assert a.AES_GCM_decrypt(key, iv, aad, want) == (0, pt)
    bad = bytearray(want); bad[-1] ^= 1
    rc, o = a.AES_GCM_decrypt(key, iv, aad, bytes(bad))
    assert rc == 0x1A and o == b"\xcc" * n
    # pageable memory under fan-out (the helpers are shared by the parts)
    pg = np.frombuffer(pt, dtype=np.uint8).copy()
    uaes.ctr_crypt_range(128, key, iv, 0, pg.ctypes.data, n, pg.ctypes.data)
    assert pg.tobytes() == orc.ctr(key, iv, pt)
    assert uaes.core().uaes_last_error() == 0


def test_threads_drive_the_library_concurrently(uaes, orc, torch):
    """no global lock any more: several host threads issue staged and direct calls at once"""
    import threading
    key, iv = rnd("th-k", 16), rnd("th-i", 12)
    errs = []

    def work(i):
        try:
            n = 3 * MIB + 17 * i
            pt = rnd(f"th-p{i}", n)
            for _ in range(3):
                out = hbuf(n)
                uaes.ctr_crypt_range(128, key, iv, i, pt, n, out)
                assert out.raw[:n] == orc.ctr(key, iv, pt, first_block=i)
                enc = uaes.MicroAES(128).AES_GCM_encrypt(key, iv, b"t%d" % i, pt[:100000 + i])
                assert enc == orc.gcm_encrypt(key, iv, b"t%d" % i, pt[:100000 + i])
        except Exception as e:                                             # pragma: no cover
            errs.append((i, repr(e)))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs


def test_async_calls_on_two_streams_do_not_share_scratch(uaes, orc, torch):
    """GCM / OCB / GCM-SIV calls enqueued on different user streams used to share one work
    area; every call now owns a block of the pool until its stream has passed it"""
    key = rnd("as-k", 16)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    n = 6 * MIB
    jobs = []
    for i in range(6):
        nonce, pt = rnd(f"as-n{i}", 12), rnd(f"as-p{i}", 1000) * (n // 1000)
        jobs.append((nonce, pt, dev(torch, pt), dev(torch, b"", pad=len(pt) + 32)))
    torch.cuda.synchronize()
    uaes.set_async(True)
    try:
        for i, (nonce, pt, src, dst) in enumerate(jobs):
            uaes.set_stream((s1 if i % 2 == 0 else s2).cuda_stream)
            if i % 3 == 2:
                uaes.ocb(128, key, nonce, b"aad", src, len(pt), dst, True)
            else:
                uaes.gcm_encrypt(128, key, nonce, b"aad", src, len(pt), dst)
    finally:
        uaes.set_async(False)
        uaes.set_stream(0)
    torch.cuda.synchronize()
    for i, (nonce, pt, src, dst) in enumerate(jobs):
        want = orc.ocb_encrypt(key, nonce, b"aad", pt) if i % 3 == 2 else orc.gcm_encrypt(key, nonce, b"aad", pt)
        assert host(dst, len(pt), len(pt) + 16) == want[-16:], i
        assert sha256(host(dst, 0, len(pt))) == sha256(want[:-16]), i


def test_gcm_shard_contributions_stay_on_the_device(uaes, orc, torch):
    """uaes_gcm_shard with a DEVICE `partial` + uaes_gcm_combine on device arrays (the multi-GPU bench
    gathers them device to device); any number of shards"""
    key, nonce, aad = rnd("sd-k", 16), rnd("sd-n", 12), rnd("sd-a", 77)
    n = 40 * 65536 + 5
    pt = rnd("sd-p", n)
    src, dst = dev(torch, pt), dev(torch, b"", pad=n + 16)
    nsh = 40
    parts = torch.zeros(16 * nsh, dtype=torch.uint8, device="cuda")
    after = []
    uaes.set_async(True)
    for r in range(nsh):
        off = r * 65536
        ln = 65536 if r < nsh - 1 else n - off
        uaes.gcm_shard(128, key, nonce, off // 16, src[off:], ln, dst[off:], partial_dev=parts[16 * r:])
        after.append((n + 15) // 16 - (off + ln + 15) // 16)
    tag_dev = torch.zeros(16, dtype=torch.uint8, device="cuda")
    uaes.gcm_combine(128, key, nonce, aad, None, after, n, partials_dev=parts, tag_dev=tag_dev)     # enqueued only
    uaes.set_async(False)
    tag = uaes.gcm_combine(128, key, nonce, aad, None, after, n, partials_dev=parts)
    want = orc.gcm_encrypt(key, nonce, aad, pt)
    assert host(dst, 0, n) == want[:n] and tag == want[n:] and host(tag_dev, 0, 16) == want[n:]


def test_trim_shutdown_burn(uaes, orc, torch):
    key, iv = rnd("lc-k", 16), rnd("lc-i", 12)
    pt = rnd("lc-p", 5 * MIB + 3)
    a = uaes.MicroAES(128)
    want = orc.gcm_encrypt(key, iv, b"x", pt)
    assert a.AES_GCM_encrypt(key, iv, b"x", pt) == want
    free0 = torch.cuda.mem_get_info()[0]
    uaes.trim()
    assert torch.cuda.mem_get_info()[0] >= free0
    assert a.AES_GCM_encrypt(key, iv, b"x", pt) == want
    uaes.shutdown()
    free1 = torch.cuda.mem_get_info()[0]
    assert free1 >= free0 + 3 * 60 * MIB or free1 >= free0                   # the staging chunks are gone
    assert a.AES_CTR_encrypt(key, iv, pt) == orc.ctr(key, iv, pt)            # initialises again
    uaes.set_burn(True)
    try:
        assert a.AES_GCM_encrypt(key, iv, b"x", pt) == want
        assert a.AES_GCM_decrypt(key, iv, b"x", want) == (0, pt)
        assert a.AES_CTR_encrypt(key, iv, pt) == orc.ctr(key, iv, pt)
        keys = rnd("lc-x", 32)
        assert a.AES_XTS_encrypt(keys, None, pt) == orc.xts(keys, None, pt)
        pg = np.frombuffer(pt, dtype=np.uint8).copy()
        uaes.ctr_crypt_range(128, key, iv, 0, pg.ctypes.data, pg.size, pg.ctypes.data)
        assert pg.tobytes() == orc.ctr(key, iv, pt)
    finally:
        uaes.set_burn(False)
    assert uaes.core().uaes_last_error() == 0
