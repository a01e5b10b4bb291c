"""Shared helpers for the tests: deterministic inputs, the checker bindings, golden files."""
import ctypes
import hashlib
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

_u8p = ctypes.c_char_p
_sz = ctypes.c_size_t
_u64 = ctypes.c_uint64
_int = ctypes.c_int


def rnd(tag, n):
    """deterministic bytes: SHA-256 in counter mode over `tag` (same as golden/make_golden.py)"""
    out = bytearray()
    i = 0
    while len(out) < n:
        out += hashlib.sha256(f"{tag}:{i}".encode()).digest()
        i += 1
    return bytes(out[:n])


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def sha256(b):
    return hashlib.sha256(b).hexdigest()


class Oracle:
    """ctypes binding of oracle/liboracle.so (our C restatement of the reference path)"""

    def __init__(self):
        L = self.lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        L.oracle_encrypt_block.argtypes = [_int, _u8p, _u8p, _u8p]
        L.oracle_decrypt_block.argtypes = [_int, _u8p, _u8p, _u8p]
        L.oracle_key_expansion.argtypes = [_int, _u8p, _u8p]
        L.oracle_ecb_encrypt.argtypes = [_int, _u8p, _u8p, _sz, _u8p]
        L.oracle_ecb_decrypt.argtypes = [_int, _u8p, _u8p, _sz, _u8p]
        L.oracle_ctr_crypt.argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p]
        L.oracle_ctr_crypt_at.argtypes = [_int, _u8p, _u8p, _u64, _u8p, _sz, _u8p]
        for f in (L.oracle_xts_encrypt, L.oracle_xts_decrypt):
            f.argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p]
        L.oracle_xts_sectors.argtypes = [_int, _u8p, _u64, _sz, _u8p, _sz, _u8p, _int]
        L.oracle_gcm_encrypt.argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p, _sz, _u8p]
        L.oracle_gcm_decrypt.argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p, _sz, _u8p]
        L.oracle_ghash.argtypes = [_u8p, _u8p, _sz, _u8p, _sz, _u8p]
        L.oracle_gf128_mul.argtypes = [_u8p, _u8p]
        L.oracle_ghash_absorb.argtypes = [_u8p, ctypes.c_void_p, _sz, _u8p]
        L.oracle_xts_double.argtypes = [_u8p]
        L.oracle_gcmsiv_encrypt.argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p, _sz, _u8p]
        L.oracle_gcmsiv_decrypt.argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p, _sz, _u8p]
        L.oracle_polyval.argtypes = [_u8p, _u8p, _sz, _u8p, _sz, _u8p]
        L.oracle_ocb_encrypt.argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p, _sz, _u8p]
        L.oracle_ocb_decrypt.argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p, _sz, _u8p]
        L.oracle_ccm_encrypt.argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p, _sz, _u8p]
        L.oracle_ccm_decrypt.argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p, _sz, _u8p]
        L.oracle_eax_encrypt.argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p, _sz, _u8p]
        L.oracle_eax_decrypt.argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p, _sz, _u8p]
        L.oracle_siv_encrypt.argtypes = [_int, _u8p, _u8p, _sz, _u8p, _sz, _u8p, _u8p]
        L.oracle_siv_decrypt.argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p, _sz, _u8p]
        for f in (L.oracle_cbc_decrypt, L.oracle_cbc_encrypt, L.oracle_cfb_decrypt, L.oracle_cfb_encrypt):
            f.argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p]
        L.oracle_fill_splitmix64.argtypes = [_u64, _u64, ctypes.c_void_p, _sz]
        # compile-time variants of the reference as run-time arguments
        L.oracle_ecb_encrypt_padded.argtypes = [_int, _u8p, _u8p, _sz, _u8p, _int]
        L.oracle_ctr_crypt_block.argtypes = [_int, _u8p, _u8p, _u64, _u8p, _sz, _u8p]
        L.oracle_gcm_encrypt_ex.argtypes = [_int, _u8p, _u8p, _sz, _u8p, _sz, _u8p, _sz, _u8p, _sz]
        L.oracle_gcm_decrypt_ex.argtypes = [_int, _u8p, _u8p, _sz, _u8p, _sz, _u8p, _sz, _u8p, _sz]
        L.oracle_cbc_decrypt_nocts.argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p]
        L.oracle_xts_range.argtypes = [_int, _u8p, _u8p, _u64, _u8p, _sz, _u8p, _int]
        for m in ("ccm", "eax", "ocb"):
            getattr(L, f"oracle_{m}_encrypt_ex").argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p, _sz, _u8p, _sz]
            getattr(L, f"oracle_{m}_decrypt_ex").argtypes = [_int, _u8p, _u8p, _u8p, _sz, _u8p, _sz, _u8p, _sz]

    @staticmethod
    def _buf(n):
        return ctypes.create_string_buffer(max(n, 1))

    def encrypt_block(self, key, blk):
        o = self._buf(16)
        self.lib.oracle_encrypt_block(len(key) * 8, key, blk, o)
        return o.raw[:16]

    def decrypt_block(self, key, blk):
        o = self._buf(16)
        self.lib.oracle_decrypt_block(len(key) * 8, key, blk, o)
        return o.raw[:16]

    def key_expansion(self, key):
        o = self._buf(240)
        r = self.lib.oracle_key_expansion(len(key) * 8, key, o)
        return o.raw[:16 * (r + 1)]

    def ecb_encrypt(self, key, pt):
        m = (len(pt) + 15) // 16 * 16
        o = self._buf(m)
        self.lib.oracle_ecb_encrypt(len(key) * 8, key, pt, len(pt), o)
        return o.raw[:m]

    def ecb_decrypt(self, key, ct):
        o = self._buf(len(ct))
        rc = self.lib.oracle_ecb_decrypt(len(key) * 8, key, ct, len(ct), o)
        return rc, o.raw[:len(ct)]

    def ctr(self, key, iv, data, first_block=0):
        o = self._buf(len(data))
        self.lib.oracle_ctr_crypt_at(len(key) * 8, key, iv, first_block, data, len(data), o)
        return o.raw[:len(data)]

    def xts(self, keys, tweak, data, encrypt=True):
        o = self._buf(len(data))
        f = self.lib.oracle_xts_encrypt if encrypt else self.lib.oracle_xts_decrypt
        rc = f(len(keys) * 4, keys, tweak, data, len(data), o)
        return rc, o.raw[:len(data)]

    def ecb_encrypt_padded(self, key, pt, padding):
        m = (len(pt) // 16 + 1) * 16 if padding else (len(pt) + 15) // 16 * 16
        o = self._buf(m)
        self.lib.oracle_ecb_encrypt_padded(len(key) * 8, key, pt, len(pt), o, padding)
        return o.raw[:m]

    def ctr_block(self, key, ctr16, data, first_block=0):
        o = self._buf(len(data))
        self.lib.oracle_ctr_crypt_block(len(key) * 8, key, ctr16, first_block, data, len(data), o)
        return o.raw[:len(data)]

    def gcm_encrypt_ex(self, key, nonce, aad, pt, taglen=16):
        o = self._buf(len(pt) + 16)
        self.lib.oracle_gcm_encrypt_ex(len(key) * 8, key, nonce, len(nonce), aad, len(aad), pt, len(pt), o, taglen)
        return o.raw[:len(pt) + taglen]

    def gcm_decrypt_ex(self, key, nonce, aad, ct_and_tag, taglen=16):
        n = len(ct_and_tag) - taglen
        o = self._buf(n)
        rc = self.lib.oracle_gcm_decrypt_ex(len(key) * 8, key, nonce, len(nonce), aad, len(aad), ct_and_tag, n, o, taglen)
        return rc, o.raw[:n]

    def aead_ex(self, mode, key, nonce, aad, data, taglen, encrypt=True):
        """CCM / EAX / OCB with a tag of `taglen` bytes: encrypt -> ct || tag, decrypt -> (rc, pt)"""
        if encrypt:
            o = self._buf(len(data) + 16)
            getattr(self.lib, f"oracle_{mode}_encrypt_ex")(len(key) * 8, key, nonce, aad, len(aad), data, len(data), o, taglen)
            return o.raw[:len(data) + taglen]
        n = len(data) - taglen
        o = self._buf(n)
        rc = getattr(self.lib, f"oracle_{mode}_decrypt_ex")(len(key) * 8, key, nonce, aad, len(aad), data, n, o, taglen)
        return rc, o.raw[:n]

    def cbc_nocts(self, key, iv, data):
        o = self._buf(len(data))
        rc = self.lib.oracle_cbc_decrypt_nocts(len(key) * 8, key, iv, data, len(data), o)
        return rc, o.raw[:len(data)]

    def xts_range(self, keys, tweak, first_block, data, encrypt=True):
        """blocks [first_block, ...) of one data unit; walks the tweak chain from T_0 (slow for far offsets)"""
        o = self._buf(len(data))
        rc = self.lib.oracle_xts_range(len(keys) * 4, keys, tweak, first_block, data, len(data), o, 1 if encrypt else 0)
        return rc, o.raw[:len(data)]

    def xts_sectors(self, keys, first_sector, sector_bytes, data, encrypt=True):
        o = self._buf(len(data))
        rc = self.lib.oracle_xts_sectors(len(keys) * 4, keys, first_sector, sector_bytes, data,
                                         len(data), o, 1 if encrypt else 0)
        return rc, o.raw[:len(data)]

    def gcm_encrypt(self, key, nonce, aad, pt):
        o = self._buf(len(pt) + 16)
        self.lib.oracle_gcm_encrypt(len(key) * 8, key, nonce, aad, len(aad), pt, len(pt), o)
        return o.raw[:len(pt) + 16]

    def gcm_decrypt(self, key, nonce, aad, ct_and_tag):
        n = len(ct_and_tag) - 16
        o = self._buf(n)
        rc = self.lib.oracle_gcm_decrypt(len(key) * 8, key, nonce, aad, len(aad), ct_and_tag, n, o)
        return rc, o.raw[:n]

    def gcmsiv_encrypt(self, key, nonce, aad, pt):
        o = self._buf(len(pt) + 16)
        self.lib.oracle_gcmsiv_encrypt(len(key) * 8, key, nonce, aad, len(aad), pt, len(pt), o)
        return o.raw[:len(pt) + 16]

    def gcmsiv_decrypt(self, key, nonce, aad, ct_and_tag):
        n = len(ct_and_tag) - 16
        o = self._buf(n)
        rc = self.lib.oracle_gcmsiv_decrypt(len(key) * 8, key, nonce, aad, len(aad), ct_and_tag, n, o)
        return rc, o.raw[:n]

    def cbc(self, key, iv, data, encrypt=False):
        o = self._buf(len(data))
        f = self.lib.oracle_cbc_encrypt if encrypt else self.lib.oracle_cbc_decrypt
        rc = f(len(key) * 8, key, iv, data, len(data), o)
        return rc, o.raw[:len(data)]

    def cfb(self, key, iv, data, encrypt=False):
        o = self._buf(len(data))
        f = self.lib.oracle_cfb_encrypt if encrypt else self.lib.oracle_cfb_decrypt
        f(len(key) * 8, key, iv, data, len(data), o)
        return o.raw[:len(data)]

    def ocb_encrypt(self, key, nonce, aad, pt):
        o = self._buf(len(pt) + 16)
        self.lib.oracle_ocb_encrypt(len(key) * 8, key, nonce, aad, len(aad), pt, len(pt), o)
        return o.raw[:len(pt) + 16]

    def ocb_decrypt(self, key, nonce, aad, ct_and_tag):
        n = len(ct_and_tag) - 16
        o = self._buf(n)
        rc = self.lib.oracle_ocb_decrypt(len(key) * 8, key, nonce, aad, len(aad), ct_and_tag, n, o)
        return rc, o.raw[:n]

    def ccm_encrypt(self, key, nonce, aad, pt):
        o = self._buf(len(pt) + 16)
        self.lib.oracle_ccm_encrypt(len(key) * 8, key, nonce, aad, len(aad), pt, len(pt), o)
        return o.raw[:len(pt) + 16]

    def ccm_decrypt(self, key, nonce, aad, ct_and_tag):
        """(rc, plaintext): the plaintext is produced even when rc = 0x1A (micro_aes.c:1304-1312)"""
        n = len(ct_and_tag) - 16
        o = self._buf(n)
        rc = self.lib.oracle_ccm_decrypt(len(key) * 8, key, nonce, aad, len(aad), ct_and_tag, n, o)
        return rc, o.raw[:n]

    def eax_encrypt(self, key, nonce, aad, pt):
        o = self._buf(len(pt) + 16)
        self.lib.oracle_eax_encrypt(len(key) * 8, key, nonce, aad, len(aad), pt, len(pt), o)
        return o.raw[:len(pt) + 16]

    def eax_decrypt(self, key, nonce, aad, ct_and_tag):
        """(rc, plaintext); the output buffer (0xCC filled here) is untouched when rc = 0x1A"""
        n = len(ct_and_tag) - 16
        o = ctypes.create_string_buffer(b"\xcc" * max(n, 1), max(n, 1))
        rc = self.lib.oracle_eax_decrypt(len(key) * 8, key, nonce, aad, len(aad), ct_and_tag, n, o)
        return rc, o.raw[:n]

    def siv_encrypt(self, keys, aad, pt):
        """IV || ciphertext, the layout main.c:214 uses"""
        iv, o = self._buf(16), self._buf(len(pt))
        self.lib.oracle_siv_encrypt(len(keys) * 4, keys, aad, len(aad), pt, len(pt), iv, o)
        return iv.raw[:16] + o.raw[:len(pt)]

    def siv_decrypt(self, keys, aad, iv_and_ct):
        n = len(iv_and_ct) - 16
        o = self._buf(n)
        rc = self.lib.oracle_siv_decrypt(len(keys) * 4, keys, iv_and_ct[:16], aad, len(aad), iv_and_ct[16:], n, o)
        return rc, o.raw[:n]

    def polyval(self, H, aad, pt):
        o = self._buf(16)
        self.lib.oracle_polyval(H, aad, len(aad), pt, len(pt), o)
        return o.raw[:16]

    def ghash(self, H, aad, ct):
        o = self._buf(16)
        self.lib.oracle_ghash(H, aad, len(aad), ct, len(ct), o)
        return o.raw[:16]

    def ghash_absorb(self, H, data_ptr, nbytes, state=bytes(16)):
        """data_ptr: address (int) or bytes"""
        o = ctypes.create_string_buffer(state, 16)
        if isinstance(data_ptr, (bytes, bytearray)):
            data_ptr = ctypes.cast(ctypes.c_char_p(bytes(data_ptr)), ctypes.c_void_p).value
        self.lib.oracle_ghash_absorb(H, data_ptr, nbytes, o)
        return o.raw[:16]

    def gf128_pow(self, x, e):
        """x^e by square and multiply with the oracle's mulGF128"""
        r, b = b"\x80" + bytes(15), x
        while e:
            if e & 1:
                r = self.gf128_mul(b, r)
            b = self.gf128_mul(b, b)
            e >>= 1
        return r

    def gf128_mul(self, x, y):
        o = ctypes.create_string_buffer(y, 16)
        self.lib.oracle_gf128_mul(x, o)
        return o.raw[:16]

    def splitmix(self, seed, first_word, nwords):
        o = self._buf(8 * nwords)
        self.lib.oracle_fill_splitmix64(seed, first_word, o, nwords)
        return o.raw[:8 * nwords]


class Reference:
    """ctypes binding of oracle/_ref/libref<bits>.so: the UNMODIFIED micro_aes.c.  The key
    length is baked into each library (AES___, micro_aes.h:17)."""

    def __init__(self, bits, preset_counter=False, variant=""):
        name = f"libref{bits}{'pc' if preset_counter else variant}.so"
        self.path = os.path.join(ROOT, "oracle", "_ref", name)
        self.bits = bits
        self.lib = ctypes.CDLL(self.path)
        for f in ("AES_ECB_decrypt", "AES_XTS_encrypt", "AES_XTS_decrypt", "AES_GCM_decrypt", "GCM_SIV_decrypt",
                  "AES_CBC_encrypt", "AES_CBC_decrypt", "AES_OCB_decrypt", "AES_CCM_decrypt",
                  "AES_EAX_decrypt", "AES_SIV_decrypt"):
            getattr(self.lib, f).restype = ctypes.c_char

    @staticmethod
    def available(bits=128, preset_counter=False, variant=""):
        return os.path.exists(os.path.join(ROOT, "oracle", "_ref",
                                           f"libref{bits}{'pc' if preset_counter else variant}.so"))

    # ---- builds with other compile-time settings (variant = "iv1", "iv128", "tag12", "pad1", "pad2", "cts0")
    def gcm_encrypt_v(self, key, nonce, aad, pt, taglen=16):
        o = ctypes.create_string_buffer(len(pt) + 16)
        self.lib.AES_GCM_encrypt(key, nonce, aad, _sz(len(aad)), pt, _sz(len(pt)), o)
        return o.raw[:len(pt) + taglen]

    def gcm_decrypt_v(self, key, nonce, aad, ct_and_tag, taglen=16):
        n = len(ct_and_tag) - taglen
        o = ctypes.create_string_buffer(b"\xcc" * (n + 16), n + 16)
        rc = self.lib.AES_GCM_decrypt(key, nonce, aad, _sz(len(aad)), ct_and_tag, _sz(n), o)
        return ord(rc), o.raw[:n]

    def ecb_encrypt_padded(self, key, pt):
        m = (len(pt) // 16 + 1) * 16
        o = ctypes.create_string_buffer(m + 16)
        self.lib.AES_ECB_encrypt(key, pt, _sz(len(pt)), o)
        return o.raw[:m]

    def ecb_encrypt(self, key, pt):
        m = (len(pt) + 15) // 16 * 16
        o = ctypes.create_string_buffer(m + 16)
        self.lib.AES_ECB_encrypt(key, pt, _sz(len(pt)), o)
        return o.raw[:m]

    def ecb_decrypt(self, key, ct):
        o = ctypes.create_string_buffer(len(ct) + 16)
        rc = self.lib.AES_ECB_decrypt(key, ct, _sz(len(ct)), o)
        return ord(rc), o.raw[:len(ct)]

    def ctr(self, key, iv, data):
        o = ctypes.create_string_buffer(len(data) + 16)
        self.lib.AES_CTR_encrypt(key, iv, data, _sz(len(data)), o)
        return o.raw[:len(data)]

    def xts(self, keys, tweak, data, encrypt=True):
        o = ctypes.create_string_buffer(len(data) + 16)
        f = self.lib.AES_XTS_encrypt if encrypt else self.lib.AES_XTS_decrypt
        rc = f(keys, tweak, data, _sz(len(data)), o)
        return ord(rc), o.raw[:len(data)]

    def gcm_encrypt(self, key, nonce, aad, pt):
        o = ctypes.create_string_buffer(len(pt) + 16)
        self.lib.AES_GCM_encrypt(key, nonce, aad, _sz(len(aad)), pt, _sz(len(pt)), o)
        return o.raw[:len(pt) + 16]

    def ocb_encrypt(self, key, nonce, aad, pt):
        o = ctypes.create_string_buffer(len(pt) + 16)
        self.lib.AES_OCB_encrypt(key, nonce, aad, _sz(len(aad)), pt, _sz(len(pt)), o)
        return o.raw[:len(pt) + 16]

    def ocb_decrypt(self, key, nonce, aad, ct_and_tag):
        n = len(ct_and_tag) - 16
        o = ctypes.create_string_buffer(n + 16)
        rc = self.lib.AES_OCB_decrypt(key, nonce, aad, _sz(len(aad)), ct_and_tag, _sz(n), o)
        return ord(rc), o.raw[:n]

    def ccm_encrypt(self, key, nonce, aad, pt):
        o = ctypes.create_string_buffer(len(pt) + 16)
        self.lib.AES_CCM_encrypt(key, nonce, aad, _sz(len(aad)), pt, _sz(len(pt)), o)
        return o.raw[:len(pt) + 16]

    def ccm_decrypt(self, key, nonce, aad, ct_and_tag):
        n = len(ct_and_tag) - 16
        o = ctypes.create_string_buffer(n + 16)
        rc = self.lib.AES_CCM_decrypt(key, nonce, aad, _sz(len(aad)), ct_and_tag, _sz(n), o)
        return ord(rc), o.raw[:n]

    def eax_encrypt(self, key, nonce, aad, pt):
        o = ctypes.create_string_buffer(len(pt) + 16)
        self.lib.AES_EAX_encrypt(key, nonce, aad, _sz(len(aad)), pt, _sz(len(pt)), o)
        return o.raw[:len(pt) + 16]

    def eax_decrypt(self, key, nonce, aad, ct_and_tag):
        n = len(ct_and_tag) - 16
        o = ctypes.create_string_buffer(b"\xcc" * (n + 16), n + 16)
        rc = self.lib.AES_EAX_decrypt(key, nonce, aad, _sz(len(aad)), ct_and_tag, _sz(n), o)
        return ord(rc), o.raw[:n]

    def siv_encrypt(self, keys, aad, pt):
        iv, o = ctypes.create_string_buffer(16), ctypes.create_string_buffer(len(pt) + 16)
        self.lib.AES_SIV_encrypt(keys, aad, _sz(len(aad)), pt, _sz(len(pt)), iv, o)
        return iv.raw[:16] + o.raw[:len(pt)]

    def siv_decrypt(self, keys, aad, iv_and_ct):
        n = len(iv_and_ct) - 16
        o = ctypes.create_string_buffer(n + 16)
        rc = self.lib.AES_SIV_decrypt(keys, iv_and_ct[:16], aad, _sz(len(aad)), iv_and_ct[16:], _sz(n), o)
        return ord(rc), o.raw[:n]

    def cbc(self, key, iv, data, encrypt=False):
        o = ctypes.create_string_buffer(b"\xcc" * (len(data) + 16), len(data) + 16)
        f = self.lib.AES_CBC_encrypt if encrypt else self.lib.AES_CBC_decrypt
        rc = f(key, iv, data, _sz(len(data)), o)
        return ord(rc), o.raw[:len(data)]

    def cfb(self, key, iv, data, encrypt=False):
        o = ctypes.create_string_buffer(len(data) + 16)
        f = self.lib.AES_CFB_encrypt if encrypt else self.lib.AES_CFB_decrypt
        f(key, iv, data, _sz(len(data)), o)
        return o.raw[:len(data)]

    def gcmsiv_encrypt(self, key, nonce, aad, pt):
        o = ctypes.create_string_buffer(len(pt) + 16)
        self.lib.GCM_SIV_encrypt(key, nonce, aad, _sz(len(aad)), pt, _sz(len(pt)), o)
        return o.raw[:len(pt) + 16]

    def gcmsiv_decrypt(self, key, nonce, aad, ct_and_tag):
        n = len(ct_and_tag) - 16
        o = ctypes.create_string_buffer(n + 16)
        rc = self.lib.GCM_SIV_decrypt(key, nonce, aad, _sz(len(aad)), ct_and_tag, _sz(n), o)
        return ord(rc), o.raw[:n]

    def gcm_decrypt(self, key, nonce, aad, ct_and_tag):
        n = len(ct_and_tag) - 16
        o = ctypes.create_string_buffer(n + 16)
        rc = self.lib.AES_GCM_decrypt(key, nonce, aad, _sz(len(aad)), ct_and_tag, _sz(n), o)
        return ord(rc), o.raw[:n]
