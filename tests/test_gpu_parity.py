"""GPU parity tests: the CUDA path, called through the C ABI (libmicro_aes_<bits>.so with the
reference's symbol names, and the uaes_* extensions of libuaes_b200.so), against

  * the reference's own known-answer vectors (tests/golden/*.json), bit-exact;
  * the pinned CPU oracle on seeded inputs, including the edge cases the reference exercises
    (empty / ragged inputs, stealing, auth failure) and the ones it cannot (counter carries,
    device pointers, in-place, misaligned pointers);
  * size-independent properties at BASELINE.json's full sizes (test_gpu_fullsize.py).

Everything here is integer/byte work: the bar is bit-exact, no tolerance.
"""
import ctypes
import importlib

import pytest

from util import Oracle, golden, rnd, sha256

pytestmark = pytest.mark.gpu
H = bytes.fromhex


@pytest.fixture(scope="module")
def uaes():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    mod = importlib.import_module("micro-aes_b200")
    assert mod.core().uaes_device_count() >= 1
    return mod


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.fixture(scope="module")
def torch():
    return pytest.importorskip("torch")


def dev(torch, data, pad=0, offset=0):
    """device uint8 tensor holding `data` at byte offset `offset` of an over-allocated buffer"""
    import numpy as np
    t = torch.zeros(offset + len(data) + pad + 16, dtype=torch.uint8, device="cuda")
    if data:
        t[offset:offset + len(data)] = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
    return t


def host(t, a, b):
    return bytes(t[a:b].cpu().numpy())


# ---------------------------------------------------------------- the reference's own vectors

def test_fips197_and_main_c(uaes):
    m = golden("main_c.json")
    a = uaes.MicroAES(128)
    v = m["fips197_c1"]
    assert a.AES_ECB_encrypt(H(v["key"]), H(v["pt"])) == H(v["ct"])          # BASELINE config 1
    assert a.AES_ECB_decrypt(H(v["key"]), H(v["ct"])) == (0, H(v["pt"]))
    key, iv16, pt = H(m["key_pool"]), H(m["iv16"]), H(m["plaintext"])
    assert a.AES_ECB_encrypt(key[:16], pt) == H(m["ecb128"])                  # main.c:139-145
    rc, back = a.AES_ECB_decrypt(key[:16], H(m["ecb128"]))
    assert rc == 0 and back == pt + bytes(7)
    assert a.AES_CTR_encrypt(key[:16], iv16[:12], pt) == H(m["ctr128"])       # main.c:167-173
    assert a.AES_CTR_decrypt(key[:16], iv16[:12], H(m["ctr128"])) == pt
    for bits in (128, 256):
        b = uaes.MicroAES(bits)
        assert b.AES_XTS_encrypt(key[:bits // 4], iv16, pt) == (0, H(m[f"xts{bits}"]))   # main.c:174-180
        assert b.AES_XTS_decrypt(key[:bits // 4], iv16, H(m[f"xts{bits}"])) == (0, pt)
        out = b.AES_GCM_encrypt(key[:bits // 8], iv16[:12], H(m["aad"]), pt)  # main.c:191-197
        assert out == H(m[f"gcm{bits}"])
        assert b.AES_GCM_decrypt(key[:bits // 8], iv16[:12], H(m["aad"]), out) == (0, pt)


@pytest.mark.parametrize("bits,expected", [(128, 800), (256, 600)])
def test_xts_rsp(uaes, bits, expected):
    a = uaes.MicroAES(bits)
    cases = golden(f"xts{bits}.json")["cases"]
    assert len(cases) == expected
    for c in cases:
        assert a.AES_XTS_encrypt(H(c["key"]), H(c["i"]), H(c["pt"])) == (0, H(c["ct"])), c
        assert a.AES_XTS_decrypt(H(c["key"]), H(c["i"]), H(c["ct"])) == (0, H(c["pt"])), c


@pytest.mark.parametrize("bits", [128, 192, 256])
def test_gcm_rsp(uaes, bits):
    a = uaes.MicroAES(bits)
    cases = golden(f"gcm{bits}.json")["cases"]
    assert len(cases) == 375
    for c in cases:
        out = a.AES_GCM_encrypt(H(c["key"]), H(c["iv"]), H(c["aad"]), H(c["pt"]))
        assert out == H(c["ct"]) + H(c["tag"]), c
        assert a.AES_GCM_decrypt(H(c["key"]), H(c["iv"]), H(c["aad"]), out) == (0, H(c["pt"])), c


def test_recorded_reference_outputs(uaes):
    """outputs of the unmodified reference recorded by tests/golden/make_golden.py"""
    s = golden("oracle_ref_samples.json")
    lib = {b: uaes.MicroAES(b) for b in (128, 192, 256)}
    for c in s["ctr"]:
        ct = lib[c["bits"]].AES_CTR_encrypt(H(c["key"]), H(c["iv"]), rnd(c["pt_tag"], c["n"]))
        assert sha256(ct) == c["ct_sha256"], c
    for c in s["ecb"]:
        assert sha256(lib[c["bits"]].AES_ECB_encrypt(H(c["key"]), rnd(c["pt_tag"], c["n"]))) == c["ct_sha256"], c
    for c in s["xts"]:
        rc, ct = lib[c["bits"]].AES_XTS_encrypt(H(c["keys"]), H(c["tweak"]), rnd(c["pt_tag"], c["n"]))
        assert rc == 0 and sha256(ct) == c["ct_sha256"], c
        assert lib[c["bits"]].AES_XTS_decrypt(H(c["keys"]), H(c["tweak"]), ct) == (0, rnd(c["pt_tag"], c["n"]))
    for c in s["gcm"]:
        out = lib[c["bits"]].AES_GCM_encrypt(H(c["key"]), H(c["nonce"]), rnd(c["aad_tag"], c["aadlen"]),
                                             rnd(c["pt_tag"], c["n"]))
        assert sha256(out[:-16]) == c["ct_sha256"] and out[-16:].hex() == c["tag"], c
    for c in s["xts_sectors"]:
        n = c["sector_bytes"] * c["sectors"]
        pt = rnd(c["pt_tag"], n)
        out = ctypes.create_string_buffer(n)
        uaes.xts_sectors(c["bits"], H(c["keys"]), c["first_sector"], c["sector_bytes"], pt, n, out)
        assert sha256(out.raw) == c["ct_sha256"], c


def test_counter_carries(uaes, orc):
    """56-bit big-endian counter (micro_aes.c:421-427): carry out of byte 12 into the nonce bytes
    and the wrap at 2^56 that leaves byte 8 alone; pinned by the PRESET_COUNTER reference run"""
    s = golden("oracle_ref_samples.json")
    c = next(x for x in s["ctr_preset_counter"] if x["name"] == "into_nonce_byte11")
    key, pt, iv = H(c["key"]), rnd(c["pt_tag"], 128), H(c["counter0"])[:12]
    out = ctypes.create_string_buffer(128)
    uaes.ctr_crypt_range(128, key, iv, 0xFFFFFFFE - 1, pt, 128, out)
    assert out.raw.hex() == c["ct"]
    c = next(x for x in s["ctr_preset_counter"] if x["name"] == "wrap56")
    key, pt, blk = H(c["key"]), rnd(c["pt_tag"], 128), H(c["counter0"])
    iv = blk[:9] + b"\xff\xff\xff"                     # field = ffffff|00000001 after the ^1
    first = (int.from_bytes(blk[9:], "big") - ((0xFFFFFF << 32) | 1)) % (1 << 56)
    uaes.ctr_crypt_range(128, key, iv, first, pt, 128, out)
    assert out.raw.hex() == c["ct"]
    # and against the oracle across a group (256-counter) boundary and the 2^32 boundary
    key, iv = rnd("cc-k", 32), rnd("cc-i", 12)
    for first in (0, 200, 254, 255, 256, 0xFFFFFF00, 0xFFFFFFFF - 40, (1 << 40) - 3, (1 << 56) - 300):
        data = rnd(f"cc-d{first}", 16 * 600 + 7)
        o = ctypes.create_string_buffer(len(data))
        uaes.ctr_crypt_range(256, key, iv, first, data, len(data), o)
        assert o.raw == orc.ctr(key, iv, data, first_block=first), first


# ---------------------------------------------------------------- seeded differential vs oracle

SIZES = [0, 1, 15, 16, 17, 31, 32, 33, 255, 256, 257, 511, 512, 513, 4095, 4096, 4097,
         16 * 255, 16 * 256 + 9, 65536 - 1, 65536, (1 << 20) + 5]


@pytest.mark.parametrize("bits", [128, 192, 256])
def test_ctr_ecb_sizes(uaes, orc, bits):
    a = uaes.MicroAES(bits)
    for n in SIZES:
        key, iv, data = rnd(f"s-k{bits}{n}", bits // 8), rnd(f"s-i{bits}{n}", 12), rnd(f"s-d{bits}{n}", n)
        assert a.AES_CTR_encrypt(key, iv, data) == orc.ctr(key, iv, data), n
        assert a.AES_ECB_encrypt(key, data) == orc.ecb_encrypt(key, data), n
        assert a.AES_ECB_decrypt(key, data) == orc.ecb_decrypt(key, data), n


@pytest.mark.parametrize("bits", [128, 192, 256])
def test_ctr_bitsliced_corunner(uaes, orc, torch, bits):
    """the CTR kernels' ALU co-runner warps (bitsliced AES, csrc/uaes_bitslice.cuh) forced on for
    small calls, at every static split between the two kinds of warps and through the work queue, with
    ragged ends, counter offsets that are not multiples of 1024 (or of a queue unit) and the carries
    into counter bytes 13, 12 and 9"""
    key, iv = rnd(f"bs-k{bits}", bits // 8), rnd(f"bs-i{bits}", 12)
    try:
        for threads in (384, 385, 386, 388):     # 385 = two blocks per table-driven thread in flight; 386 = the work-queue kernel;
                                                 # 388 = the work queue with narrow bitsliced warps (uaes_bitslice8.cuh)
            for share, n, first in ((1024, 16 * 5000 + 3, 0), (512, 16 * 70001, 1), (300, (1 << 21) + 9, 1000),
                                    (1024, 16 * 3000, (1 << 16) - 1500), (700, 16 * 4100 + 15, (1 << 24) - 2050),
                                    (1024, 16 * 2500, (1 << 32) - 1200), (900, 16 * 2048, (1 << 56) - 1024 - 2),
                                    (1, 16 * 9000, 7), (1024, 16 * 1023, 1), (1024, 16, 0)):
                uaes.ctr_tuning(threads, share, 0)
                data = rnd(f"bs-d{bits}{n}", n)
                src, dst = dev(torch, data), dev(torch, b"", pad=n)
                uaes.ctr_crypt_range(bits, key, iv, first, src, n, dst)
                torch.cuda.synchronize()
                assert host(dst, 0, n) == orc.ctr(key, iv, data, first_block=first), (threads, share, n, first)
                assert host(dst, n, n + 16) == bytes(16)
    finally:
        uaes.ctr_tuning(388, 195, 1 << 23)      # library defaults (uaes_kernels.cu kCtrDefaultGeometry / kCtrDefaultShare)


@pytest.mark.parametrize("bits", [128, 256])
def test_xts_sizes(uaes, orc, bits):
    a = uaes.MicroAES(bits)
    for n in [s for s in SIZES if s >= 16] + [16 * 1024 * 40 + 3, (1 << 22) + 15]:
        keys, tw, data = rnd(f"x-k{bits}{n}", bits // 4), rnd(f"x-t{bits}{n}", 16), rnd(f"x-d{bits}{n}", n)
        want = orc.xts(keys, tw, data)
        assert a.AES_XTS_encrypt(keys, tw, data) == want, n
        assert a.AES_XTS_decrypt(keys, tw, want[1]) == (0, data), n
    assert a.AES_XTS_encrypt(rnd("x-k", bits // 4), bytes(16), b"short") == (1, b"\0" * 5)   # M_DATALENGTH_ERROR
    keys, data = rnd("x-k0", bits // 4), rnd("x-d0", 80)
    assert a.AES_XTS_encrypt(keys, None, data) == orc.xts(keys, None, data)    # NULL tweak = sector 0


@pytest.mark.parametrize("bits", [128, 256])
def test_xts_sectors(uaes, orc, bits):
    for first, sb, ns in ((0, 512, 100), ((1 << 32) - 5, 512, 70), (3, 4096, 33), (99, 16, 1000),
                          (1 << 60, 528, 65), (7, 8192, 5)):
        keys, data = rnd(f"xs-k{bits}{first}", bits // 4), rnd(f"xs-d{bits}{first}", sb * ns)
        out = ctypes.create_string_buffer(len(data))
        uaes.xts_sectors(bits, keys, first, sb, data, len(data), out, True)
        assert (0, out.raw) == orc.xts_sectors(keys, first, sb, data), (first, sb, ns)
        back = ctypes.create_string_buffer(len(data))
        uaes.xts_sectors(bits, keys, first, sb, out.raw, len(data), back, False)
        assert back.raw == data


@pytest.mark.parametrize("bits", [128, 192, 256])
def test_ecb_bitsliced_corunner(uaes, orc, bits):
    """ECB encryption with the co-runner forced on for small calls: all splits, ragged tiles, padded tail"""
    a = uaes.MicroAES(bits)
    try:
        for share, n in ((1024, 16 * 2048), (512, 16 * 5000 + 7), (300, 16 * 70001), (1024, 16 * 3071 + 15), (1, 16 * 4096)):
            uaes.ctr_tuning(-1, share, 0)
            key, data = rnd(f"eh-k{bits}{n}", bits // 8), rnd(f"eh-d{bits}{n}", n)
            assert a.AES_ECB_encrypt(key, data) == orc.ecb_encrypt(key, data), (share, n)
            # decryption: ecb_dec_hybrid_kernel (table-driven Td rounds + the bitsliced equivalent inverse cipher)
            assert a.AES_ECB_decrypt(key, data) == orc.ecb_decrypt(key, data), (share, n)
            iv = rnd(f"eh-i{bits}{n}", 16)                       # CFB decryption rides the same kernel
            o = ctypes.create_string_buffer(max(n, 1))
            orc.lib.oracle_cfb_decrypt(bits, key, iv, data, n, o)
            assert a.AES_CFB_decrypt(key, iv, data) == o.raw[:n], (share, n)
    finally:
        uaes.ctr_tuning(388, 195, 1 << 23)      # library defaults (uaes_kernels.cu kCtrDefaultGeometry / kCtrDefaultShare)


@pytest.mark.parametrize("bits", [128, 192, 256])
def test_cbc_decrypt_bitsliced_corunner(uaes, orc, bits):
    """CBC decryption through ecb_dec_hybrid_kernel<NR, true> (table-driven Td rounds + bitsliced inverse cipher, XOR with
    the previous ciphertext block) forced on for small calls: all splits, whole blocks and CS3 stealing pairs"""
    a = uaes.MicroAES(bits)
    try:
        for share, n in ((1024, 16 * 2048), (512, 16 * 5000 + 7), (300, 16 * 70001), (1024, 16 * 3071 + 15), (1, 16 * 4096 + 1),
                         (700, 16 * 2049)):
            uaes.ctr_tuning(-1, share, 0)
            key, iv, ct = rnd(f"ch-k{bits}{n}", bits // 8), rnd(f"ch-i{bits}{n}", 16), rnd(f"ch-d{bits}{n}", n)
            assert a.AES_CBC_decrypt(key, iv, ct) == orc.cbc(key, iv, ct), (share, n)
    finally:
        uaes.ctr_tuning(388, 195, 1 << 23)      # library defaults (uaes_kernels.cu kCtrDefaultGeometry / kCtrDefaultShare)


@pytest.mark.parametrize("bits", [128, 256])
def test_ocb_bitsliced_corunner(uaes, orc, bits):
    """OCB encryption with the co-runner forced on for small calls: offsets, checksum and tag must not
    depend on which kind of warp handled a block"""
    a = uaes.MicroAES(bits)
    try:
        for share, n, alen in ((1024, 16 * 2048, 0), (512, 16 * 5000 + 7, 20), (300, 16 * 70001, 5), (1024, 16 * 3071 + 15, 0),
                               (1, 16 * 4096, 33)):
            uaes.ctr_tuning(-1, share, 0)
            key, nonce = rnd(f"oh-k{bits}{n}", bits // 8), rnd(f"oh-n{bits}{n}", 12)
            aad, data = rnd(f"oh-a{bits}{n}", alen), rnd(f"oh-d{bits}{n}", n)
            want = orc.ocb_encrypt(key, nonce, aad, data)
            got = a.AES_OCB_encrypt(key, nonce, aad, data)
            assert got[-16:] == want[-16:] and got == want, (share, n)
            assert a.AES_OCB_decrypt(key, nonce, aad, want) == (0, data)
    finally:
        uaes.ctr_tuning(388, 195, 1 << 23)      # library defaults (uaes_kernels.cu kCtrDefaultGeometry / kCtrDefaultShare)


@pytest.mark.parametrize("narrow", [1, 0])
@pytest.mark.parametrize("bits", [128, 192, 256])
def test_xts_sectors_bitsliced_corunner(uaes, orc, bits, narrow, monkeypatch):
    """512-byte sector encryption with the ALU co-runner warps forced on for small calls, at several
    splits between table-driven and bitsliced tiles, ragged last tiles, sector numbers across 2^32;
    narrow = 1: xts_sectors_hybrid8_kernel (8 sectors per bitsliced pass, uaes_bitslice8.cuh, both directions
    through the work queue), narrow = 0: the wide form"""
    monkeypatch.setenv("UAES_XTS_NARROW", str(narrow))
    try:
        for share, first, ns in ((1024, 0, 33), (512, 5, 100), (300, (1 << 32) - 40, 1000), (1024, 1 << 40, 64),
                                 (700, 9, 4096 + 7), (1, 3, 2048), (1024, 0, 1)):
            uaes.ctr_tuning(-1, share, 0)
            keys, data = rnd(f"xh-k{bits}{first}", bits // 4), rnd(f"xh-d{bits}{first}{ns}", 512 * ns)
            out = ctypes.create_string_buffer(len(data))
            uaes.xts_sectors(bits, keys, first, 512, data, len(data), out, True)
            assert (0, out.raw) == orc.xts_sectors(keys, first, 512, data), (share, first, ns)
            back = ctypes.create_string_buffer(len(data))
            uaes.xts_sectors(bits, keys, first, 512, out.raw, len(data), back, False)
            assert back.raw == data
            # the decrypting co-runner (bitsliced inverse cipher) against the oracle on unrelated input
            uaes.xts_sectors(bits, keys, first, 512, data, len(data), back, False)
            assert (0, back.raw) == orc.xts_sectors(keys, first, 512, data, encrypt=False), (share, first, ns)
    finally:
        uaes.ctr_tuning(388, 195, 1 << 23)      # library defaults (uaes_kernels.cu kCtrDefaultGeometry / kCtrDefaultShare)


@pytest.mark.parametrize("bits", [128, 256])
def test_xts_unit_bitsliced_corunner(uaes, orc, bits):
    """one large data unit (the reference's AES_XTS_encrypt) with the co-runner forced on: all splits,
    ragged last tile, ciphertext stealing"""
    a = uaes.MicroAES(bits)
    try:
        for share, n in ((1024, 16 * 2048), (512, 16 * 5000 + 7), (300, 16 * 70001 + 15), (1024, 16 * 3071 + 1), (1, 16 * 4096)):
            uaes.ctr_tuning(-1, share, 0)
            keys, tw, data = rnd(f"xu-k{bits}{n}", bits // 4), rnd(f"xu-t{bits}{n}", 16), rnd(f"xu-d{bits}{n}", n)
            want = orc.xts(keys, tw, data)
            assert a.AES_XTS_encrypt(keys, tw, data) == want, (share, n)
            assert a.AES_XTS_decrypt(keys, tw, want[1]) == (0, data)
    finally:
        uaes.ctr_tuning(388, 195, 1 << 23)      # library defaults (uaes_kernels.cu kCtrDefaultGeometry / kCtrDefaultShare)


@pytest.mark.parametrize("bits", [128, 192, 256])
def test_gcm_sizes(uaes, orc, bits):
    a = uaes.MicroAES(bits)
    for n, alen in [(0, 0), (0, 13), (1, 0), (15, 1), (16, 16), (17, 90), (511, 20), (512, 0), (513, 31),
                    (4096, 7), (16 * 255, 0), (16 * 257 + 3, 129), (65536 + 3, 20), ((1 << 20) + 5, 20),
                    (3 * (1 << 20) + 1, 0), (1000, 70000 + 5), (0, 4096), (100, 1 << 20)]:   # last three: bulk-hashed AAD
        key, nonce = rnd(f"g-k{bits}{n}", bits // 8), rnd(f"g-n{bits}{n}", 12)
        aad, data = rnd(f"g-a{bits}{n}", alen), rnd(f"g-d{bits}{n}", n)
        want = orc.gcm_encrypt(key, nonce, aad, data)
        got = a.AES_GCM_encrypt(key, nonce, aad, data)
        assert got[-16:] == want[-16:], (n, alen)
        assert got == want, (n, alen)
        assert a.AES_GCM_decrypt(key, nonce, aad, want) == (0, data), (n, alen)


def test_gcm_multirow_chunks_ragged(uaes, orc):
    """The bulk kernel cuts GHASH into one chunk per warp of the grid (R rows of 32 blocks each) and
    absorbs the block of row r while the rounds of row r+1 run: messages with 2..5 rows per chunk,
    a first row and a last row that are only partly inside the message, and a ragged tail."""
    a = uaes.MicroAES(128)
    for rows_total, extra_blocks, tail in [(2 * 2960, 17, 5), (3 * 2960 - 1, 31, 0), (5 * 2960 + 7, 1, 15),
                                           (2 * 3552 + 3, 9, 1), (4 * 2960, 0, 0)]:
        n = 16 * (32 * rows_total + extra_blocks) + tail
        key, nonce = rnd(f"mr-k{n}", 16), rnd(f"mr-n{n}", 12)
        aad, data = rnd(f"mr-a{n}", 21), rnd(f"mr-d{n}", n)
        want = orc.gcm_encrypt(key, nonce, aad, data)
        got = a.AES_GCM_encrypt(key, nonce, aad, data)
        assert got[-16:] == want[-16:], n
        assert got == want, n
        assert a.AES_GCM_decrypt(key, nonce, aad, want) == (0, data), n
        wsiv = orc.gcmsiv_encrypt(key, nonce, aad, data)
        assert a.GCM_SIV_encrypt(key, nonce, aad, data) == wsiv, n
        assert a.GCM_SIV_decrypt(key, nonce, aad, wsiv) == (0, data), n


def test_gcm_auth_failure_leaves_output_untouched(uaes, orc):
    a = uaes.MicroAES(128)
    key, nonce, aad, data = rnd("af-k", 16), rnd("af-n", 12), rnd("af-a", 20), rnd("af-d", 1000)
    enc = bytearray(orc.gcm_encrypt(key, nonce, aad, data))
    for pos in (0, 500, len(enc) - 1):
        bad = bytearray(enc)
        bad[pos] ^= 0x10
        rc, out = a.AES_GCM_decrypt(key, nonce, aad, bytes(bad))
        assert rc == 0x1A and out == b"\xcc" * 1000            # micro_aes.c:1204-1208
    rc, _ = a.AES_GCM_decrypt(key, nonce, aad + b"x", bytes(enc))
    assert rc == 0x1A


# ---------------------------------------------------------------- device pointers

def test_device_pointers_inplace_and_misaligned(uaes, orc, torch):
    key, iv = rnd("dp-k", 16), rnd("dp-i", 12)
    n = 100000 + 7
    data = rnd("dp-d", n)
    want = orc.ctr(key, iv, data, first_block=5)
    # aligned, out of place; bytes after the end must stay zero
    src, dst = dev(torch, data), dev(torch, b"", pad=n)
    uaes.ctr_crypt_range(128, key, iv, 5, src, n, dst)
    torch.cuda.synchronize()
    assert host(dst, 0, n) == want and host(dst, n, n + 16) == bytes(16)
    # in place (micro_aes.h:520-526: in == out must work)
    uaes.ctr_crypt_range(128, key, iv, 5, src, n, src)
    assert host(src, 0, n) == want
    # misaligned device pointers go through the staging path and still match
    src = dev(torch, data, offset=3)
    dst = dev(torch, b"", pad=n, offset=9)
    uaes.ctr_crypt_range(128, key, iv, 5, src.data_ptr() + 3, n, dst.data_ptr() + 9)
    assert host(dst, 9, 9 + n) == want and host(dst, 0, 9) == bytes(9)
    # mixed: host in, device out
    dst = dev(torch, b"", pad=n)
    uaes.ctr_crypt_range(128, key, iv, 5, data, n, dst)
    assert host(dst, 0, n) == want


def test_device_gcm_xts_ecb(uaes, orc, torch):
    key, nonce, aad = rnd("dg-k", 32), rnd("dg-n", 12), rnd("dg-a", 33)
    n = (1 << 21) + 11
    data = rnd("dg-d", n)
    want = orc.gcm_encrypt(key, nonce, aad, data)
    src, dst = dev(torch, data), dev(torch, b"", pad=n + 16)
    uaes.gcm_encrypt(256, key, nonce, aad, src, n, dst)
    assert host(dst, 0, n + 16) == want
    back = dev(torch, b"", pad=n)
    assert uaes.gcm_decrypt(256, key, nonce, aad, dst, n, back) == 0
    assert host(back, 0, n) == data
    dst[n + 3] ^= 1                                            # corrupt the tag on the device
    untouched = dev(torch, b"\xaa" * n)
    assert uaes.gcm_decrypt(256, key, nonce, aad, dst, n, untouched) == 0x1A
    assert host(untouched, 0, n) == b"\xaa" * n

    keys, tw = rnd("dx-k", 64), rnd("dx-t", 16)
    want = orc.xts(keys, tw, data)[1]
    src = dev(torch, data)
    uaes.xts_unit(256, keys, tw, src, n, src, True)            # in place, with stealing
    assert host(src, 0, n) == want
    uaes.xts_unit(256, keys, tw, src, n, src, False)
    assert host(src, 0, n) == data

    m = n - n % 512
    want = orc.xts_sectors(keys, 77, 512, data[:m])[1]
    uaes.xts_sectors(256, keys, 77, 512, src, m, src, True)
    assert host(src, 0, m) == want

    src = dev(torch, data)
    dst = dev(torch, b"", pad=n + 16)
    uaes.ecb(192, key[:24], src, n, dst, True)
    assert host(dst, 0, (n + 15) // 16 * 16) == orc.ecb_encrypt(key[:24], data)


def test_large_host_buffers_use_chunked_pipeline(uaes, orc):
    """> 3 staging chunks of 32 MiB, ragged: exercises slot reuse and the counter offset per chunk"""
    from concurrent.futures import ThreadPoolExecutor
    import os
    n = 100 * (1 << 20) + 13
    data = orc.splitmix(0xABCDEF, 0, (n + 7) // 8)[:n]
    key, iv = rnd("lh-k", 16), rnd("lh-i", 12)
    got = uaes.MicroAES(128).AES_CTR_encrypt(key, iv, data)
    piece = 1 << 20
    with ThreadPoolExecutor(os.cpu_count() or 4) as ex:
        parts = list(ex.map(lambda o: orc.ctr(key, iv, data[o:o + piece], first_block=o // 16), range(0, n, piece)))
    assert got == b"".join(parts)


def test_async_stream_api(uaes, orc, torch):
    key, iv, data = rnd("as-k", 16), rnd("as-i", 12), rnd("as-d", 1 << 16)
    s = torch.cuda.Stream()
    src, dst = dev(torch, data), dev(torch, b"", pad=len(data))
    uaes.set_stream(s.cuda_stream)
    uaes.set_async(True)
    try:
        uaes.ctr_crypt_range(128, key, iv, 0, src, len(data), dst)
        s.synchronize()
    finally:
        uaes.set_async(False)
        uaes.set_stream(None)
    assert host(dst, 0, len(data)) == orc.ctr(key, iv, data)


def test_gcm_sharded_message(uaes, orc, torch):
    """one GCM message cut into shards (as over several GPUs), each shard one fused
    pass returning 16 bytes, folded by uaes_gcm_combine: ciphertext and tag == the unsharded oracle"""
    for bits, n, alen, cuts in [(128, 3 * (1 << 20) + 7, 20, [1 << 20, (2 << 20) + 4096]),
                                (256, 100000, 0, [16, 32, 48, 99968]),
                                (192, 4096, 33, []),
                                (128, 5, 7, []),
                                (128, 0, 13, [])]:
        key, nonce = rnd(f"sh-k{bits}{n}", bits // 8), rnd(f"sh-n{bits}{n}", 12)
        aad, data = rnd(f"sh-a{bits}{n}", alen), rnd(f"sh-d{bits}{n}", n)
        want = orc.gcm_encrypt(key, nonce, aad, data)
        edges = [0] + cuts + [n]
        total_blocks = (n + 15) // 16
        for device in (False, True):
            parts, after, out = [], [], b""
            for a, b in zip(edges[:-1], edges[1:]):
                if device:
                    src, dst = dev(torch, data[a:b]), dev(torch, b"", pad=b - a)
                    parts.append(uaes.gcm_shard(bits, key, nonce, a // 16, src, b - a, dst))
                    out += host(dst, 0, b - a)
                else:
                    dst = ctypes.create_string_buffer(max(b - a, 1))
                    parts.append(uaes.gcm_shard(bits, key, nonce, a // 16, data[a:b], b - a, dst))
                    out += dst.raw[:b - a]
                after.append(total_blocks - (b + 15) // 16)
            tag = uaes.gcm_combine(bits, key, nonce, aad, parts, after, n)
            assert out == want[:-16], (bits, n, device)
            assert tag == want[-16:], (bits, n, device)
        # decrypting shards: plaintext back, same tag from the INPUT ciphertext
        parts, after, back = [], [], b""
        for a, b in zip(edges[:-1], edges[1:]):
            dst = ctypes.create_string_buffer(max(b - a, 1))
            parts.append(uaes.gcm_shard(bits, key, nonce, a // 16, want[a:b], b - a, dst, decrypt=True))
            back += dst.raw[:b - a]
            after.append(total_blocks - (b + 15) // 16)
        assert back == data and uaes.gcm_combine(bits, key, nonce, aad, parts, after, n) == want[-16:]


# ---------------------------------------------------------------- GCM-SIV

def test_gcmsiv_reference_vectors(uaes):
    m = golden("main_c.json")
    a = uaes.MicroAES(128)
    key, nonce, aad, pt = H(m["key_pool"])[:16], H(m["iv16"])[:12], H(m["aad"]), H(m["plaintext"])
    out = a.GCM_SIV_encrypt(key, nonce, aad, pt)               # main.c:219-224
    assert out == H(m["gcmsiv128"])
    assert a.GCM_SIV_decrypt(key, nonce, aad, out) == (0, pt)
    for v in m["gcmsiv_rfc8452"]:                              # main.c:275-299
        assert a.GCM_SIV_encrypt(H(v["key"]), H(v["iv"]), H(v["aad"]), H(v["pt"])) == H(v["ct"])
    cases = golden("gcmsiv128.json")["cases"]                  # testvectors/SIV_GCM_ACVP.tv
    assert len(cases) == 102
    for c in cases:
        assert a.GCM_SIV_encrypt(H(c["key"]), H(c["iv"]), H(c["aad"]), H(c["pt"])) == H(c["ct"]), c
        assert a.GCM_SIV_decrypt(H(c["key"]), H(c["iv"]), H(c["aad"]), H(c["ct"])) == (0, H(c["pt"])), c


def test_gcmsiv_recorded_reference_and_oracle(uaes, orc, torch):
    s = golden("oracle_ref_samples.json")
    lib = {b: uaes.MicroAES(b) for b in (128, 192, 256)}
    for c in s["gcmsiv"]:
        out = lib[c["bits"]].GCM_SIV_encrypt(H(c["key"]), H(c["nonce"]), rnd(c["aad_tag"], c["aadlen"]),
                                             rnd(c["pt_tag"], c["n"]))
        assert sha256(out[:-16]) == c["ct_sha256"] and out[-16:].hex() == c["tag"], c
    for c in s["gcmsiv_forged_tag_decrypt"]:                   # 32-bit LE counter wrap; auth must fail
        rc, out = lib[c["bits"]].GCM_SIV_decrypt(H(c["key"]), H(c["nonce"]), b"", rnd(c["ct_tag"], c["n"]) + H(c["tag"]))
        assert rc == 0x1A and sha256(out) == c["out_sha256"], c
    for bits in (128, 256):
        for n, alen in [(0, 0), (0, 5), (15, 0), (16, 1), (33, 100), (4095, 7), ((1 << 20) + 9, 20), (3 * (1 << 20), 0),
                        (500, 70000 + 5), (0, 4096 + 16)]:
            key, nonce = rnd(f"gv-k{bits}{n}", bits // 8), rnd(f"gv-n{bits}{n}", 12)
            aad, data = rnd(f"gv-a{bits}{n}", alen), rnd(f"gv-d{bits}{n}", n)
            want = orc.gcmsiv_encrypt(key, nonce, aad, data)
            assert lib[bits].GCM_SIV_encrypt(key, nonce, aad, data) == want, (bits, n)
            assert lib[bits].GCM_SIV_decrypt(key, nonce, aad, want) == (0, data), (bits, n)
            if n:
                bad = bytearray(want)
                bad[n // 2] ^= 1
                assert lib[bits].GCM_SIV_decrypt(key, nonce, aad, bytes(bad))[0] == 0x1A
    # device pointers, in place
    key, nonce, aad = rnd("gv-dk", 16), rnd("gv-dn", 12), rnd("gv-da", 33)
    n = (1 << 21) + 5
    data = rnd("gv-dd", n)
    want = orc.gcmsiv_encrypt(key, nonce, aad, data)
    buf = dev(torch, data, pad=16)
    uaes.gcmsiv(128, key, nonce, aad, buf, n, buf, True)
    assert host(buf, 0, n + 16) == want
    assert uaes.gcmsiv(128, key, nonce, aad, buf, n, buf, False) == 0
    assert host(buf, 0, n) == data


# ---------------------------------------------------------------- CBC / CFB decrypt

def test_cbc_cfb_decrypt(uaes, orc, torch):
    m = golden("main_c.json")
    a = uaes.MicroAES(128)
    key, iv, pt = H(m["key_pool"])[:16], H(m["iv16"]), H(m["plaintext"])
    assert a.AES_CBC_decrypt(key, iv, H(m["cbc128_cts"])) == (0, pt)         # main.c:146-152 (CTS)
    assert a.AES_CFB_decrypt(key, iv, H(m["cfb128"])) == pt                  # main.c:153-159
    s = golden("oracle_ref_samples.json")
    lib = {b: uaes.MicroAES(b) for b in (128, 192, 256)}
    for c in s["cbc_decrypt"]:                                               # unmodified reference runs
        rc, out = lib[c["bits"]].AES_CBC_decrypt(H(c["key"]), H(c["iv"]), rnd(c["ct_tag"], c["n"]))
        assert rc == c["rc"], c
        if rc == 0:
            assert sha256(out) == c["pt_sha256"], c
        else:
            assert out == b"\xcc" * c["n"]                                   # M_DATALENGTH_ERROR: nothing written
    for c in s["cfb_decrypt"]:
        assert sha256(lib[c["bits"]].AES_CFB_decrypt(H(c["key"]), H(c["iv"]), rnd(c["ct_tag"], c["n"]))) == c["pt_sha256"], c
    for bits in (128, 192, 256):
        for n in [16, 17, 31, 32, 33, 47, 48, 49, 511, 512, 513, 4096, 4097, 65536 + 15, (1 << 20) + 16, (1 << 20) + 7]:
            key, iv, ct = rnd(f"cb-k{bits}{n}", bits // 8), rnd(f"cb-i{bits}{n}", 16), rnd(f"cb-c{bits}{n}", n)
            assert lib[bits].AES_CBC_decrypt(key, iv, ct) == orc.cbc(key, iv, ct), (bits, n)
            assert lib[bits].AES_CFB_decrypt(key, iv, ct) == orc.cfb(key, iv, ct), (bits, n)
    # valid ciphertexts made by the oracle's serial encrypt come back as the plaintext
    key, iv, pt = rnd("cb-rk", 32), rnd("cb-ri", 16), rnd("cb-rp", 100000 + 3)
    assert lib[256].AES_CBC_decrypt(key, iv, orc.cbc(key, iv, pt, encrypt=True)[1]) == (0, pt)
    assert lib[256].AES_CFB_decrypt(key, iv, orc.cfb(key, iv, pt, encrypt=True)) == pt
    # device pointers (out of place) and the in == out request (staged copy)
    n = (1 << 21) + 9
    ct = rnd("cb-dc", n)
    src, dst = dev(torch, ct), dev(torch, b"", pad=n)
    uaes.chain_decrypt(128, key[:16], iv, src, n, dst, cbc=True)
    assert host(dst, 0, n) == orc.cbc(key[:16], iv, ct)[1] and host(dst, n, n + 16) == bytes(16)
    uaes.chain_decrypt(128, key[:16], iv, src, n, dst, cbc=False)
    assert host(dst, 0, n) == orc.cfb(key[:16], iv, ct)
    uaes.chain_decrypt(128, key[:16], iv, src, n, src, cbc=True)
    assert host(src, 0, n) == orc.cbc(key[:16], iv, ct)[1]


# ---------------------------------------------------------------- OCB

def test_ocb(uaes, orc, torch):
    m = golden("main_c.json")
    a = uaes.MicroAES(128)
    key, nonce, aad, pt = H(m["key_pool"])[:16], H(m["iv16"])[:12], H(m["aad"]), H(m["plaintext"])
    out = a.AES_OCB_encrypt(key, nonce, aad, pt)                  # main.c:204-210
    assert out == H(m["ocb128"])
    assert a.AES_OCB_decrypt(key, nonce, aad, out) == (0, pt)
    v = m["ocb_rfc7253"]                                          # main.c:262-274
    assert a.AES_OCB_encrypt(H(v["key"]), H(v["iv"]), H(v["aad"]), H(v["pt"])) == H(v["ct"])
    cases = golden("ocb128.json")["cases"]                        # testvectors/OCB_AES128.tv
    assert len(cases) == 16
    for c in cases:
        assert a.AES_OCB_encrypt(H(c["key"]), H(c["iv"]), H(c["aad"]), H(c["pt"])) == H(c["ct"]), c
        assert a.AES_OCB_decrypt(H(c["key"]), H(c["iv"]), H(c["aad"]), H(c["ct"])) == (0, H(c["pt"])), c
    lib = {b: uaes.MicroAES(b) for b in (128, 192, 256)}
    for c in golden("oracle_ref_samples.json")["ocb"]:            # unmodified reference runs
        out = lib[c["bits"]].AES_OCB_encrypt(H(c["key"]), H(c["nonce"]), rnd(c["aad_tag"], c["aadlen"]), rnd(c["pt_tag"], c["n"]))
        assert sha256(out[:-16]) == c["ct_sha256"] and out[-16:].hex() == c["tag"], c
    for bits in (128, 256):
        for n, alen in [(0, 0), (0, 5), (1, 0), (15, 16), (16, 0), (17, 1), (511, 33), (512, 0), (513, 100), (4096, 7),
                        (16 * 1024 + 3, 20), ((1 << 20) + 9, 20), (3 * (1 << 20), 5000 + 7)]:
            key, nonce = rnd(f"oc-k{bits}{n}", bits // 8), rnd(f"oc-n{bits}{n}", 12)
            aad, data = rnd(f"oc-a{bits}{n}", alen), rnd(f"oc-d{bits}{n}", n)
            want = orc.ocb_encrypt(key, nonce, aad, data)
            assert lib[bits].AES_OCB_encrypt(key, nonce, aad, data) == want, (bits, n, alen)
            assert lib[bits].AES_OCB_decrypt(key, nonce, aad, want) == (0, data), (bits, n, alen)
            if n:
                bad = bytearray(want)
                bad[n // 2] ^= 1
                assert lib[bits].AES_OCB_decrypt(key, nonce, aad, bytes(bad)) == orc.ocb_decrypt(key, nonce, aad, bytes(bad))
    # every value of "bottom" (the last six nonce bits select the Stretch window)
    key, data = rnd("oc-bk", 16), rnd("oc-bd", 100)
    for b in range(64):
        nonce = rnd("oc-bn", 11) + bytes([0x40 | b])
        assert a.AES_OCB_encrypt(key, nonce, b"", data) == orc.ocb_encrypt(key, nonce, b"", data), b
    # device pointers, in place
    key, nonce, aad = rnd("oc-dk", 16), rnd("oc-dn", 12), rnd("oc-da", 33)
    n = (1 << 21) + 5
    data = rnd("oc-dd", n)
    want = orc.ocb_encrypt(key, nonce, aad, data)
    buf = dev(torch, data, pad=16)
    uaes.ocb(128, key, nonce, aad, buf, n, buf, True)
    assert host(buf, 0, n + 16) == want
    assert uaes.ocb(128, key, nonce, aad, buf, n, buf, False) == 0
    assert host(buf, 0, n) == data


# ---------------------------------------------------------------- CCM, batched

def test_ccm_reference_vectors(uaes, orc):
    m = golden("main_c.json")
    a = uaes.MicroAES(128)
    key, nonce, aad, pt = H(m["key_pool"])[:16], H(m["iv16"])[:11], H(m["aad"]), H(m["plaintext"])
    assert a.AES_CCM_encrypt(key, nonce, aad, pt) == H(m["ccm128"])                  # main.c:198-204
    assert a.AES_CCM_decrypt(key, nonce, aad, H(m["ccm128"])) == (0, pt)
    for bits in (128, 192, 256):
        a = uaes.MicroAES(bits)
        for c in golden(f"ccm{bits}.json")["cases"]:                                 # testvectors/VNT*.rsp
            assert a.AES_CCM_encrypt(H(c["key"]), H(c["nonce"]), H(c["aad"]), H(c["pt"])) == H(c["ct"]), c
            assert a.AES_CCM_decrypt(H(c["key"]), H(c["nonce"]), H(c["aad"]), H(c["ct"])) == (0, H(c["pt"])), c
        for c in [x for x in golden("oracle_ref_samples_row4.json")["ccm"] if x["bits"] == bits]:
            out = a.AES_CCM_encrypt(H(c["key"]), H(c["nonce"]), rnd(c["aad_tag"], c["aadlen"]), rnd(c["pt_tag"], c["n"]))
            assert sha256(out[:-16]) == c["ct_sha256"] and out[-16:].hex() == c["tag"], c


def _ccm_batch_case(uaes, seed, n, max_len, max_aad, align):
    """n messages with seeded lengths packed at `align`-byte aligned offsets (align = 1: none)"""
    import random
    r = random.Random(seed)
    msgs = (uaes.Msg * n)()
    in_pos = out_pos = aad_pos = 0
    up = lambda v: (v + align - 1) // align * align
    for i in range(n):
        ln = r.choice([0, 1, 15, 16, 17, 31, 32, 33, max_len]) if r.random() < 0.3 else r.randrange(max_len + 1)
        al = r.choice([0, 0, 1, 13, 14, 15, 16, 30, max_aad]) if r.random() < 0.5 else r.randrange(max_aad + 1)
        in_pos, out_pos, aad_pos = up(in_pos + r.randrange(3)), up(out_pos + r.randrange(5)), up(aad_pos + r.randrange(2))
        msgs[i].in_off, msgs[i].out_off, msgs[i].aad_off = in_pos, out_pos, aad_pos
        msgs[i].len, msgs[i].aad_len = ln, al
        nonce = rnd(f"ccmb-n{seed}-{i}", 11)
        for j in range(11):
            msgs[i].nonce[j] = nonce[j]
        in_pos += ln + 16; out_pos += ln + 16; aad_pos += al        # room for the tag on both sides
    return msgs, in_pos, out_pos, aad_pos


@pytest.mark.parametrize("bits,seed,n,max_len,max_aad,align", [(128, 1, 700, 300, 40, 1), (256, 2, 3000, 100, 20, 16),
                                                               (192, 3, 64, 5000, 70000, 4), (128, 4, 1, 1000, 10, 1)])
def test_ccm_batch_matches_oracle(uaes, orc, torch, bits, seed, n, max_len, max_aad, align):
    key = rnd(f"ccmb-k{bits}", bits // 8)
    msgs, in_sz, out_sz, aad_sz = _ccm_batch_case(uaes, seed, n, max_len, max_aad, align)
    pt, aad = bytearray(rnd(f"ccmb-p{seed}", in_sz)), rnd(f"ccmb-a{seed}", max(aad_sz, 1))
    want = bytearray(b"\xee" * out_sz)
    for m in msgs:
        want[m.out_off:m.out_off + m.len + 16] = orc.ccm_encrypt(key, bytes(m.nonce[:11]), aad[m.aad_off:m.aad_off + m.aad_len],
                                                                 bytes(pt[m.in_off:m.in_off + m.len]))
    # host buffers
    out = bytearray(b"\xee" * out_sz)
    assert uaes.ccm_batch(bits, key, msgs, n, aad, pt, out) == 0
    assert out == want                                   # bytes between messages untouched
    assert all(m.result == 0 for m in msgs)
    # device buffers (descriptors stay on the host), then everything on the device
    d_in, d_out, d_aad = dev(torch, bytes(pt)), dev(torch, b"", pad=out_sz), dev(torch, aad)
    assert uaes.ccm_batch(bits, key, msgs, n, d_aad, d_in, d_out) == 0
    got = bytearray(host(d_out, 0, out_sz))
    for m in msgs:
        assert got[m.out_off:m.out_off + m.len + 16] == want[m.out_off:m.out_off + m.len + 16]
    d_msgs = dev(torch, bytes(msgs))
    d_out2 = dev(torch, b"", pad=out_sz)
    assert uaes.ccm_batch(bits, key, d_msgs, n, d_aad, d_in, d_out2) == 0
    torch.cuda.synchronize()
    assert host(d_out2, 0, out_sz) == host(d_out, 0, out_sz)

    # decrypt: swap the roles of the offsets, forge every 7th tag
    dec = (uaes.Msg * n)()
    for i, m in enumerate(msgs):
        dec[i].in_off, dec[i].out_off, dec[i].aad_off = m.out_off, m.in_off, m.aad_off
        dec[i].len, dec[i].aad_len = m.len, m.aad_len
        for j in range(11):
            dec[i].nonce[j] = m.nonce[j]
    ct = bytearray(want)
    forged = set(range(0, n, 7)) if n > 1 else set()
    for i in forged:
        ct[msgs[i].out_off + msgs[i].len + 5] ^= 0x40
    back = bytearray(b"\xdd" * in_sz)
    rc = uaes.ccm_batch(bits, key, dec, n, aad, ct, back, decrypt=True)
    assert rc == (0x1A if forged else 0)
    for i, m in enumerate(msgs):
        assert dec[i].result == (0x1A if i in forged else 0), i
        assert back[m.in_off:m.in_off + m.len] == pt[m.in_off:m.in_off + m.len], i    # produced either way (micro_aes.c:1304-1312)


def test_eax_siv_reference_vectors(uaes, orc):
    m = golden("main_c.json")
    a = uaes.MicroAES(128)
    key, aad, pt = H(m["key_pool"])[:16], H(m["aad"]), H(m["plaintext"])
    assert a.AES_EAX_encrypt(key, H(m["iv16"]), aad, pt) == H(m["eax128"])           # main.c:225-237
    assert a.AES_EAX_decrypt(key, H(m["iv16"]), aad, H(m["eax128"])) == (0, pt)
    for c in golden("eax128.json")["cases"]:                                         # testvectors/EAX_AES128.tv
        assert a.AES_EAX_encrypt(H(c["key"]), H(c["nonce"]), H(c["aad"]), H(c["pt"])) == H(c["ct"]), c
        assert a.AES_EAX_decrypt(H(c["key"]), H(c["nonce"]), H(c["aad"]), H(c["ct"])) == (0, H(c["pt"])), c
    keys = H(m["key_pool"])[:32]
    assert a.AES_SIV_encrypt(keys, aad, pt) == H(m["siv128"])                        # main.c:212-218
    assert a.AES_SIV_decrypt(keys, aad, H(m["siv128"])) == (0, pt)
    for v in m["siv_extra"]:                                                         # main.c:300-321
        assert a.AES_SIV_encrypt(H(v["keys"]), H(v["aad"]), H(v["pt"])) == H(v["out"])
        assert a.AES_SIV_decrypt(H(v["keys"]), H(v["aad"]), H(v["out"])) == (0, H(v["pt"]))
    for bits in (128, 192, 256):
        a = uaes.MicroAES(bits)
        for c in [x for x in golden("oracle_ref_samples_row4.json")["eax"] if x["bits"] == bits]:
            out = a.AES_EAX_encrypt(H(c["key"]), H(c["nonce"]), rnd(c["aad_tag"], c["aadlen"]), rnd(c["pt_tag"], c["n"]))
            assert sha256(out[:-16]) == c["ct_sha256"] and out[-16:].hex() == c["tag"], c
        for c in [x for x in golden("oracle_ref_samples_row4.json")["siv"] if x["bits"] == bits]:
            out = a.AES_SIV_encrypt(H(c["keys"]), rnd(c["aad_tag"], c["aadlen"]), rnd(c["pt_tag"], c["n"]))
            assert out[:16].hex() == c["iv"] and sha256(out[16:]) == c["ct_sha256"], c
    # EAX authenticates before it decrypts: the output stays as it was (micro_aes.c:1637-1645)
    enc = bytearray(orc.eax_encrypt(key, H(m["iv16"]), aad, pt)); enc[0] ^= 1
    assert a.__class__(128).AES_EAX_decrypt(key, H(m["iv16"]), aad, bytes(enc)) == (0x1A, b"\xcc" * len(pt))


@pytest.mark.parametrize("mode", ["eax", "siv", "gcm"])
@pytest.mark.parametrize("bits,seed,n,max_len,max_aad,align", [(128, 11, 600, 300, 40, 1), (256, 12, 2000, 100, 20, 16),
                                                               (192, 13, 48, 5000, 3000, 4)])
def test_eax_siv_batch_matches_oracle(uaes, orc, torch, mode, bits, seed, n, max_len, max_aad, align):
    key = rnd(f"{mode}b-k{bits}", bits // 8 * (2 if mode == "siv" else 1))
    msgs, in_sz, out_sz, aad_sz = _ccm_batch_case(uaes, seed, n, max_len, max_aad, align)
    for i in range(n):                                   # EAX uses all 16 nonce bytes
        nonce = rnd(f"{mode}b-n{seed}-{i}", 16)
        for j in range(16):
            msgs[i].nonce[j] = nonce[j]
    pt, aad = bytearray(rnd(f"{mode}b-p{seed}", in_sz)), rnd(f"{mode}b-a{seed}", max(aad_sz, 1))
    if mode == "eax":
        enc1 = lambda m: orc.eax_encrypt(key, bytes(m.nonce), aad[m.aad_off:m.aad_off + m.aad_len], bytes(pt[m.in_off:m.in_off + m.len]))
    elif mode == "gcm":
        enc1 = lambda m: orc.gcm_encrypt(key, bytes(m.nonce[:12]), aad[m.aad_off:m.aad_off + m.aad_len], bytes(pt[m.in_off:m.in_off + m.len]))
    else:
        enc1 = lambda m: orc.siv_encrypt(key, aad[m.aad_off:m.aad_off + m.aad_len], bytes(pt[m.in_off:m.in_off + m.len]))
    want = bytearray(b"\xee" * out_sz)
    for m in msgs:
        want[m.out_off:m.out_off + m.len + 16] = enc1(m)
    out = bytearray(b"\xee" * out_sz)
    assert uaes.ccm_batch(bits, key, msgs, n, aad, pt, out, mode=mode) == 0
    assert out == want
    d_in, d_aad, d_msgs, d_out = dev(torch, bytes(pt)), dev(torch, aad), dev(torch, bytes(msgs)), dev(torch, b"", pad=out_sz)
    assert uaes.ccm_batch(bits, key, d_msgs, n, d_aad, d_in, d_out, mode=mode) == 0
    torch.cuda.synchronize()
    got = host(d_out, 0, out_sz)
    for m in msgs:
        assert got[m.out_off:m.out_off + m.len + 16] == want[m.out_off:m.out_off + m.len + 16]

    dec = (uaes.Msg * n)()
    for i, m in enumerate(msgs):
        dec[i].in_off, dec[i].out_off, dec[i].aad_off = m.out_off, m.in_off, m.aad_off
        dec[i].len, dec[i].aad_len = m.len, m.aad_len
        for j in range(16):
            dec[i].nonce[j] = m.nonce[j]
    ct = bytearray(want)
    forged = set(range(0, n, 5))
    for i in forged:                                      # EAX, GCM: the tag at the end; SIV: the IV in front
        ct[msgs[i].out_off + (3 if mode == "siv" else msgs[i].len + 7)] ^= 0x04
    back = bytearray(b"\xdd" * in_sz)
    assert uaes.ccm_batch(bits, key, dec, n, aad, ct, back, decrypt=True, mode=mode) == 0x1A
    for i, m in enumerate(msgs):
        assert dec[i].result == (0x1A if i in forged else 0), i
        if i not in forged:
            assert back[m.in_off:m.in_off + m.len] == pt[m.in_off:m.in_off + m.len], i
        elif mode in ("eax", "gcm"):                      # untouched on failure
            assert back[m.in_off:m.in_off + m.len] == b"\xdd" * m.len, i


# ---------------------------------------------------------------- streaming

@pytest.mark.parametrize("bits", [128, 256])
def test_streaming_gcm_and_ctr_match_one_shot(uaes, orc, torch, bits):
    import random
    r = random.Random(bits)
    key, nonce, aad = rnd(f"st-k{bits}", bits // 8), rnd(f"st-n{bits}", 12), rnd(f"st-a{bits}", 45)
    n = 3 * (1 << 20) + 37
    data = rnd(f"st-d{bits}", n)
    want = orc.gcm_encrypt(key, nonce, aad, data)
    # 70 pieces (more than the 30 contributions kept before a fold), all but the last multiples of 16
    cuts = sorted(r.sample(range(1, n // 16), 69))
    edges = [0] + [16 * c for c in cuts] + [n]
    st = uaes.Stream(bits, key, nonce, aad)
    out = bytearray(n)
    src = dev(torch, data)
    dst = dev(torch, b"", pad=n)
    for i, (a, b) in enumerate(zip(edges, edges[1:])):
        if i % 2:                                        # alternate host and device buffers
            piece = ctypes.create_string_buffer(b - a)
            st.update(data[a:b], b - a, piece)
            out[a:b] = piece.raw
        else:
            st.update(src[a:], b - a, dst[a:])
            torch.cuda.synchronize()
            out[a:b] = host(dst, a, b)
    tag = st.final()
    st.close()
    assert bytes(out) == want[:-16] and tag == want[-16:]
    # decrypting stream: right tag, wrong tag
    for forged in (False, True):
        st = uaes.Stream(bits, key, nonce, aad, decrypt=True)
        back = bytearray(n)
        for a, b in zip(edges, edges[1:]):
            piece = ctypes.create_string_buffer(b - a)
            st.update(want[a:b], b - a, piece)
            back[a:b] = piece.raw
        t = bytearray(want[-16:])
        t[0] ^= forged
        assert st.final(bytes(t)) == (0x1A if forged else 0)
        st.close()
        assert bytes(back) == data
    # only the last piece may be ragged
    st = uaes.Stream(bits, key, nonce, aad)
    st.update(data[:20], 20, ctypes.create_string_buffer(20))
    with pytest.raises(uaes.UaesError):
        st.update(data[20:36], 16, ctypes.create_string_buffer(16))
    st.close()
    # CTR
    st = uaes.Stream(bits, key, nonce, gcm=False)
    out = bytearray(n)
    for a, b in zip(edges, edges[1:]):
        piece = ctypes.create_string_buffer(b - a)
        st.update(data[a:b], b - a, piece)
        out[a:b] = piece.raw
    st.close()
    assert bytes(out) == orc.ctr(key, nonce, data)


def test_streaming_edge_cases(uaes, orc):
    key, nonce, aad = rnd("ste-k", 16), rnd("ste-n", 12), rnd("ste-a", 20)
    # no update at all: the tag of the empty message (with and without AAD)
    for a in (aad, b""):
        st = uaes.Stream(128, key, nonce, a)
        assert st.final() == orc.gcm_encrypt(key, nonce, a, b"")
        st.close()
    # one ragged update, zero-length updates in between
    data = rnd("ste-d", 1000 + 7)
    st = uaes.Stream(128, key, nonce, aad)
    out = ctypes.create_string_buffer(len(data))
    st.update(b"", 0, out)
    st.update(data, len(data), out)
    st.update(b"", 0, out)
    assert out.raw + st.final() == orc.gcm_encrypt(key, nonce, aad, data)
    st.close()
    # exactly 31 and 61 pieces: the fold boundary of the pending contributions
    for pieces in (30, 31, 61):
        data = rnd(f"ste-p{pieces}", 16 * pieces * 3)
        st = uaes.Stream(128, key, nonce, aad)
        got = b""
        for i in range(pieces):
            o = ctypes.create_string_buffer(48)
            st.update(data[48 * i:48 * i + 48], 48, o)
            got += o.raw
        assert got + st.final() == orc.gcm_encrypt(key, nonce, aad, data), pieces
        st.close()
