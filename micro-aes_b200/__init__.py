"""micro-aes_b200: Python mirror of the micro-AES hot-path API over the B200 engine.

The product is the C-ABI shared library `lib/libuaes_b200.so` (CUDA kernels for sm_100a plus an
ANSI-C host side) and the per-key-size shims `lib/libmicro_aes_{128,192,256}.so` that export the
reference's own symbol names (include/micro_aes.h).  This package is only a thin ctypes binding
used by the tests and bench.py; it mirrors the reference interface (same function names,
argument order and return codes, micro_aes.h:173-308) and adds the extensions of
include/uaes_b200.h (device pointers, counter ranges, sector batches).

There is no CPU fallback: importing works anywhere (so the CPU test-suite can check the ABI),
but every compute call needs a CUDA device and raises UaesError otherwise.

The directory name contains a hyphen, so import it with
    importlib.import_module("micro-aes_b200")
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.environ.get("UAES_LIBDIR") or os.path.join(_HERE, "lib")   # override: tuning builds only

M_RESULT_SUCCESS = 0
M_DATALENGTH_ERROR = 0x01
M_AUTHENTICATION_ERROR = 0x1A
M_DECRYPTION_ERROR = 0x1D
M_ENCRYPTION_ERROR = 0x1E

_vp, _sz, _u64, _int = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int
_cp = ctypes.c_char_p

# every symbol include/uaes_b200.h declares: name -> (restype, argtypes)
UAES_ABI = {
    "uaes_last_error": (_int, []),
    "uaes_last_error_string": (_cp, []),
    "uaes_clear_error": (None, []),
    "uaes_device_count": (_int, []),
    "uaes_set_stream": (None, [_vp]),
    "uaes_set_async": (None, [_int]),
    "uaes_host_alloc": (_vp, [_sz]),
    "uaes_host_free": (None, [_vp]),
    "uaes_host_register": (_int, [_vp, _sz]),
    "uaes_host_unregister": (_int, [_vp]),
    "uaes_set_devices": (_int, [_int]),
    "uaes_get_devices": (_int, []),
    "uaes_set_fanout_min": (None, [_sz]),
    "uaes_set_copy_threads": (None, [_int]),
    "uaes_set_staging": (None, [_sz, _int]),
    "uaes_set_burn": (None, [_int]),
    "uaes_trim": (None, []),
    "uaes_shutdown": (None, []),
    "uaes_kernel_launches": (_u64, []),
    "uaes_ctr_tuning": (None, [_int, _int, ctypes.c_longlong]),
    "uaes_ctr_queue_stats": (_int, [ctypes.POINTER(_u64), ctypes.POINTER(_u64), ctypes.POINTER(_u64)]),
    "uaes_ecb_encrypt": (_int, [_int, _cp, _vp, _sz, _vp]),
    "uaes_ecb_decrypt": (_int, [_int, _cp, _vp, _sz, _vp]),
    "uaes_ctr_crypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp]),
    "uaes_ctr_crypt_range": (_int, [_int, _cp, _cp, _u64, _vp, _sz, _vp]),
    "uaes_ctr_crypt_block": (_int, [_int, _cp, _cp, _u64, _vp, _sz, _vp]),
    "uaes_ecb_encrypt_padded": (_int, [_int, _cp, _vp, _sz, _vp, _int]),
    "uaes_xts_crypt_range": (_int, [_int, _cp, _cp, _u64, _vp, _sz, _vp, _int]),
    "uaes_gcm_encrypt_ex": (_int, [_int, _cp, _cp, _sz, _vp, _sz, _vp, _sz, _vp, _sz]),
    "uaes_gcm_decrypt_ex": (_int, [_int, _cp, _cp, _sz, _vp, _sz, _vp, _sz, _vp, _sz]),
    "uaes_cbc_decrypt_ex": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _int]),
    "uaes_ocb_encrypt_ex": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp, _sz]),
    "uaes_ocb_decrypt_ex": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp, _sz]),
    "uaes_ccm_encrypt_ex": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp, _sz]),
    "uaes_ccm_decrypt_ex": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp, _sz]),
    "uaes_eax_encrypt_ex": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp, _sz]),
    "uaes_eax_decrypt_ex": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp, _sz]),
    "uaes_xts_encrypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp]),
    "uaes_xts_decrypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp]),
    "uaes_xts_sectors": (_int, [_int, _cp, _u64, _sz, _vp, _sz, _vp, _int]),
    "uaes_gcm_encrypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "uaes_gcm_decrypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "uaes_ocb_encrypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "uaes_ocb_decrypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "uaes_cbc_decrypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp]),
    "uaes_cfb_decrypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp]),
    "uaes_gcmsiv_encrypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "uaes_gcmsiv_decrypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "uaes_gcm_shard": (_int, [_int, _cp, _cp, _u64, _vp, _sz, _vp, _int, _vp]),
    "uaes_gcm_combine": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _vp, _int, _u64, _vp]),
    "uaes_ccm_encrypt_batch": (_int, [_int, _cp, _vp, _sz, _vp, _vp, _vp]),
    "uaes_ccm_decrypt_batch": (_int, [_int, _cp, _vp, _sz, _vp, _vp, _vp]),
    "uaes_ccm_encrypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "uaes_ccm_decrypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "uaes_eax_encrypt_batch": (_int, [_int, _cp, _vp, _sz, _vp, _vp, _vp]),
    "uaes_eax_decrypt_batch": (_int, [_int, _cp, _vp, _sz, _vp, _vp, _vp]),
    "uaes_siv_encrypt_batch": (_int, [_int, _cp, _vp, _sz, _vp, _vp, _vp]),
    "uaes_siv_decrypt_batch": (_int, [_int, _cp, _vp, _sz, _vp, _vp, _vp]),
    "uaes_gcm_encrypt_batch": (_int, [_int, _cp, _vp, _sz, _vp, _vp, _vp]),
    "uaes_gcm_decrypt_batch": (_int, [_int, _cp, _vp, _sz, _vp, _vp, _vp]),
    "uaes_eax_encrypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "uaes_eax_decrypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "uaes_siv_encrypt": (_int, [_int, _cp, _vp, _sz, _vp, _sz, _vp, _vp]),
    "uaes_siv_decrypt": (_int, [_int, _cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "uaes_stream_ctr": (_vp, [_int, _cp, _cp]),
    "uaes_stream_gcm": (_vp, [_int, _cp, _cp, _vp, _sz, _int]),
    "uaes_stream_update": (_int, [_vp, _vp, _sz, _vp]),
    "uaes_stream_final": (_int, [_vp, _vp]),
    "uaes_stream_free": (None, [_vp]),
    "uaes_fill_splitmix64": (_int, [_u64, _u64, _vp, _sz]),
    "uaes_xor_fold64": (_int, [_vp, _sz, ctypes.POINTER(_u64)]),
}

# the eight reference symbols of include/micro_aes.h: name -> (restype, argtypes)
MICRO_AES_ABI = {
    "AES_ECB_encrypt": (None, [_cp, _vp, _sz, _vp]),
    "AES_ECB_decrypt": (ctypes.c_char, [_cp, _vp, _sz, _vp]),
    "AES_CTR_encrypt": (None, [_cp, _cp, _vp, _sz, _vp]),
    "AES_CTR_decrypt": (None, [_cp, _cp, _vp, _sz, _vp]),
    "AES_XTS_encrypt": (ctypes.c_char, [_cp, _cp, _vp, _sz, _vp]),
    "AES_XTS_decrypt": (ctypes.c_char, [_cp, _cp, _vp, _sz, _vp]),
    "AES_GCM_encrypt": (None, [_cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "AES_GCM_decrypt": (ctypes.c_char, [_cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "AES_OCB_encrypt": (None, [_cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "AES_OCB_decrypt": (ctypes.c_char, [_cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "AES_CCM_encrypt": (None, [_cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "AES_CCM_decrypt": (ctypes.c_char, [_cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "AES_EAX_encrypt": (None, [_cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "AES_EAX_decrypt": (ctypes.c_char, [_cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "AES_SIV_encrypt": (None, [_cp, _vp, _sz, _vp, _sz, _vp, _vp]),
    "AES_SIV_decrypt": (ctypes.c_char, [_cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "AES_CBC_decrypt": (ctypes.c_char, [_cp, _cp, _vp, _sz, _vp]),
    "AES_CFB_decrypt": (None, [_cp, _cp, _vp, _sz, _vp]),
    "GCM_SIV_encrypt": (None, [_cp, _cp, _vp, _sz, _vp, _sz, _vp]),
    "GCM_SIV_decrypt": (ctypes.c_char, [_cp, _cp, _vp, _sz, _vp, _sz, _vp]),
}


class UaesError(RuntimeError):
    pass


def _bind(lib, table):
    for name, (res, args) in table.items():
        f = getattr(lib, name)          # AttributeError if the library does not export it
        f.restype, f.argtypes = res, args
    return lib


_core = None
_shims = {}


def core():
    """libuaes_b200.so, bound.  Raises UaesError if it was not built (run __graft_entry__.build())."""
    global _core
    if _core is None:
        path = os.path.join(LIBDIR, "libuaes_b200.so")
        if not os.path.exists(path):
            raise UaesError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _core = _bind(ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL), UAES_ABI)
    return _core


def shim(bits):
    """libmicro_aes_<bits>.so: the reference's symbol names for one key size; `bits` may also name
    a compile-time variant built by `make variants` (e.g. "128_pc", "128_iv1", "128_pad1")."""
    if bits not in _shims:
        core()
        _shims[bits] = _bind(ctypes.CDLL(os.path.join(LIBDIR, f"libmicro_aes_{bits}.so")), MICRO_AES_ABI)
    return _shims[bits]


def _ptr(x):
    """bytes / bytearray / ctypes buffer / int device pointer / torch tensor -> address"""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, (bytes, bytearray)):
        return ctypes.cast((ctypes.c_char * len(x)).from_buffer(x) if isinstance(x, bytearray)
                           else ctypes.c_char_p(x), ctypes.c_void_p).value
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return ctypes.addressof(x)


def check(rc):
    """raise on engine failures (negative codes); pass reference result codes through"""
    if rc < 0:
        raise UaesError(f"uaes error {rc}: {core().uaes_last_error_string().decode()}")
    return rc


class MicroAES:
    """The reference API for one key size, on host `bytes` (the way main.c uses it).

    Method names, argument order and return values follow micro_aes.h; every call goes through
    the reference-named symbol of libmicro_aes_<bits>.so, i.e. through the same entry points a C
    program linked against the shim would use."""

    def __init__(self, bits=128):
        self.bits = bits
        self.lib = shim(bits)

    def _after(self):
        e = core().uaes_last_error()
        if e < 0:
            msg = core().uaes_last_error_string().decode()
            core().uaes_clear_error()
            raise UaesError(f"uaes error {e}: {msg}")

    def AES_ECB_encrypt(self, key, pntxt):
        out = ctypes.create_string_buffer(max((len(pntxt) + 15) // 16 * 16, 1))
        self.lib.AES_ECB_encrypt(key, pntxt, len(pntxt), out)
        self._after()
        return out.raw[:(len(pntxt) + 15) // 16 * 16]

    def AES_ECB_decrypt(self, key, crtxt):
        out = ctypes.create_string_buffer(max(len(crtxt), 1))
        rc = ord(self.lib.AES_ECB_decrypt(key, crtxt, len(crtxt), out))
        self._after()
        return rc, out.raw[:len(crtxt)]

    def AES_CTR_encrypt(self, key, iv, pntxt):
        out = ctypes.create_string_buffer(max(len(pntxt), 1))
        self.lib.AES_CTR_encrypt(key, iv, pntxt, len(pntxt), out)
        self._after()
        return out.raw[:len(pntxt)]

    def AES_CTR_decrypt(self, key, iv, crtxt):
        out = ctypes.create_string_buffer(max(len(crtxt), 1))
        self.lib.AES_CTR_decrypt(key, iv, crtxt, len(crtxt), out)
        self._after()
        return out.raw[:len(crtxt)]

    def AES_XTS_encrypt(self, keys, tweak, pntxt):
        out = ctypes.create_string_buffer(max(len(pntxt), 1))
        rc = ord(self.lib.AES_XTS_encrypt(keys, tweak, pntxt, len(pntxt), out))
        self._after()
        return rc, out.raw[:len(pntxt)]

    def AES_XTS_decrypt(self, keys, tweak, crtxt):
        out = ctypes.create_string_buffer(max(len(crtxt), 1))
        rc = ord(self.lib.AES_XTS_decrypt(keys, tweak, crtxt, len(crtxt), out))
        self._after()
        return rc, out.raw[:len(crtxt)]

    def AES_GCM_encrypt(self, key, nonce, aData, pntxt):
        out = ctypes.create_string_buffer(len(pntxt) + 16)
        self.lib.AES_GCM_encrypt(key, nonce, aData, len(aData), pntxt, len(pntxt), out)
        self._after()
        return out.raw[:len(pntxt) + 16]

    def AES_GCM_decrypt(self, key, nonce, aData, crtxt_and_tag):
        n = len(crtxt_and_tag) - 16
        out = ctypes.create_string_buffer(b"\xcc" * max(n, 1), max(n, 1))
        rc = ord(self.lib.AES_GCM_decrypt(key, nonce, aData, len(aData), crtxt_and_tag, n, out))
        self._after()
        return rc, out.raw[:n]


    def AES_OCB_encrypt(self, key, nonce, aData, pntxt):
        out = ctypes.create_string_buffer(len(pntxt) + 16)
        self.lib.AES_OCB_encrypt(key, nonce, aData, len(aData), pntxt, len(pntxt), out)
        self._after()
        return out.raw[:len(pntxt) + 16]

    def AES_OCB_decrypt(self, key, nonce, aData, crtxt_and_tag):
        n = len(crtxt_and_tag) - 16
        out = ctypes.create_string_buffer(max(n, 1))
        rc = ord(self.lib.AES_OCB_decrypt(key, nonce, aData, len(aData), crtxt_and_tag, n, out))
        self._after()
        return rc, out.raw[:n]

    def AES_CCM_encrypt(self, key, nonce, aData, pntxt):
        out = ctypes.create_string_buffer(len(pntxt) + 16)
        self.lib.AES_CCM_encrypt(key, nonce, aData, len(aData), pntxt, len(pntxt), out)
        self._after()
        return out.raw[:len(pntxt) + 16]

    def AES_CCM_decrypt(self, key, nonce, aData, crtxt_and_tag):
        n = len(crtxt_and_tag) - 16
        out = ctypes.create_string_buffer(max(n, 1))
        rc = ord(self.lib.AES_CCM_decrypt(key, nonce, aData, len(aData), crtxt_and_tag, n, out))
        self._after()
        return rc, out.raw[:n]

    def AES_EAX_encrypt(self, key, nonce, aData, pntxt):
        out = ctypes.create_string_buffer(len(pntxt) + 16)
        self.lib.AES_EAX_encrypt(key, nonce, aData, len(aData), pntxt, len(pntxt), out)
        self._after()
        return out.raw[:len(pntxt) + 16]

    def AES_EAX_decrypt(self, key, nonce, aData, crtxt_and_tag):
        n = len(crtxt_and_tag) - 16
        out = ctypes.create_string_buffer(b"\xcc" * max(n, 1), max(n, 1))
        rc = ord(self.lib.AES_EAX_decrypt(key, nonce, aData, len(aData), crtxt_and_tag, n, out))
        self._after()
        return rc, out.raw[:n]

    def AES_SIV_encrypt(self, keys, aData, pntxt):
        """returns IV || ciphertext (main.c:214 passes output and output + 16)"""
        iv, out = ctypes.create_string_buffer(16), ctypes.create_string_buffer(max(len(pntxt), 1))
        self.lib.AES_SIV_encrypt(keys, aData, len(aData), pntxt, len(pntxt), iv, out)
        self._after()
        return iv.raw[:16] + out.raw[:len(pntxt)]

    def AES_SIV_decrypt(self, keys, aData, iv_and_crtxt):
        n = len(iv_and_crtxt) - 16
        out = ctypes.create_string_buffer(max(n, 1))
        rc = ord(self.lib.AES_SIV_decrypt(keys, iv_and_crtxt[:16], aData, len(aData), iv_and_crtxt[16:], n, out))
        self._after()
        return rc, out.raw[:n]

    def AES_CBC_decrypt(self, key, iVec, crtxt):
        out = ctypes.create_string_buffer(b"\xcc" * max(len(crtxt), 1), max(len(crtxt), 1))
        rc = ord(self.lib.AES_CBC_decrypt(key, iVec, crtxt, len(crtxt), out))
        self._after()
        return rc, out.raw[:len(crtxt)]

    def AES_CFB_decrypt(self, key, iVec, crtxt):
        out = ctypes.create_string_buffer(max(len(crtxt), 1))
        self.lib.AES_CFB_decrypt(key, iVec, crtxt, len(crtxt), out)
        self._after()
        return out.raw[:len(crtxt)]

    def GCM_SIV_encrypt(self, key, nonce, aData, pntxt):
        out = ctypes.create_string_buffer(len(pntxt) + 16)
        self.lib.GCM_SIV_encrypt(key, nonce, aData, len(aData), pntxt, len(pntxt), out)
        self._after()
        return out.raw[:len(pntxt) + 16]

    def GCM_SIV_decrypt(self, key, nonce, aData, crtxt_and_tag):
        n = len(crtxt_and_tag) - 16
        out = ctypes.create_string_buffer(max(n, 1))
        rc = ord(self.lib.GCM_SIV_decrypt(key, nonce, aData, len(aData), crtxt_and_tag, n, out))
        self._after()
        return rc, out.raw[:n]


# ---- extensions of include/uaes_b200.h on raw pointers (device or host) ----

def ctr_crypt_range(bits, key, iv, first_block, src, nbytes, dst):
    return check(core().uaes_ctr_crypt_range(bits, key, iv, first_block, _ptr(src), nbytes, _ptr(dst)))


def ecb(bits, key, src, nbytes, dst, encrypt=True):
    f = core().uaes_ecb_encrypt if encrypt else core().uaes_ecb_decrypt
    return check(f(bits, key, _ptr(src), nbytes, _ptr(dst)))


def xts_unit(bits, keys, tweak, src, nbytes, dst, encrypt=True):
    f = core().uaes_xts_encrypt if encrypt else core().uaes_xts_decrypt
    return check(f(bits, keys, tweak, _ptr(src), nbytes, _ptr(dst)))


def xts_sectors(bits, keys, first_sector, sector_bytes, src, nbytes, dst, encrypt=True):
    return check(core().uaes_xts_sectors(bits, keys, first_sector, sector_bytes, _ptr(src), nbytes,
                                         _ptr(dst), 1 if encrypt else 0))


def gcm_encrypt(bits, key, nonce, aad, src, nbytes, dst):
    return check(core().uaes_gcm_encrypt(bits, key, nonce, _ptr(aad), len(aad) if aad else 0,
                                         _ptr(src), nbytes, _ptr(dst)))


def gcm_decrypt(bits, key, nonce, aad, src, nbytes, dst):
    return check(core().uaes_gcm_decrypt(bits, key, nonce, _ptr(aad), len(aad) if aad else 0,
                                         _ptr(src), nbytes, _ptr(dst)))


def ocb(bits, key, nonce, aad, src, nbytes, dst, encrypt=True):
    f = core().uaes_ocb_encrypt if encrypt else core().uaes_ocb_decrypt
    return check(f(bits, key, nonce, _ptr(aad), len(aad) if aad else 0, _ptr(src), nbytes, _ptr(dst)))


def chain_decrypt(bits, key, iv, src, nbytes, dst, cbc=True):
    f = core().uaes_cbc_decrypt if cbc else core().uaes_cfb_decrypt
    return check(f(bits, key, iv, _ptr(src), nbytes, _ptr(dst)))


def gcmsiv(bits, key, nonce, aad, src, nbytes, dst, encrypt=True):
    f = core().uaes_gcmsiv_encrypt if encrypt else core().uaes_gcmsiv_decrypt
    return check(f(bits, key, nonce, _ptr(aad), len(aad) if aad else 0, _ptr(src), nbytes, _ptr(dst)))


def gcm_shard(bits, key, nonce, first_block, src, nbytes, dst, decrypt=False, partial_dev=None):
    """fused CTR+GHASH over one shard of a message; returns the shard's 16-byte GHASH contribution,
    or leaves it in `partial_dev` (16 bytes of device memory) when that is given"""
    if partial_dev is not None:
        check(core().uaes_gcm_shard(bits, key, nonce, first_block, _ptr(src), nbytes, _ptr(dst),
                                    1 if decrypt else 0, _ptr(partial_dev)))
        return None
    part = ctypes.create_string_buffer(16)
    check(core().uaes_gcm_shard(bits, key, nonce, first_block, _ptr(src), nbytes, _ptr(dst),
                                1 if decrypt else 0, ctypes.addressof(part)))
    return part.raw


def gcm_combine(bits, key, nonce, aad, partials, blocks_after, total_len, partials_dev=None, tag_dev=None):
    """tag of a sharded message from the gathered contributions (a list of 16-byte strings, or
    `partials_dev` = device memory holding len(blocks_after) x 16 bytes).  With `tag_dev` (16 bytes of device
    memory) the tag stays on the GPU and, in asynchronous mode, the call only enqueues work; returns None then."""
    tag = ctypes.create_string_buffer(16)
    ps = b"".join(partials) if partials_dev is None else None
    n = len(blocks_after)
    after = (ctypes.c_uint64 * max(n, 1))(*blocks_after)
    pp = _ptr(partials_dev) if partials_dev is not None else (_ptr(ps) if ps else None)
    check(core().uaes_gcm_combine(bits, key, nonce, _ptr(aad) if aad else None, len(aad) if aad else 0,
                                  pp, ctypes.addressof(after), n, total_len,
                                  _ptr(tag_dev) if tag_dev is not None else ctypes.addressof(tag)))
    return None if tag_dev is not None else tag.raw


def ctr_crypt_block(bits, key, ctr16, first_block, src, nbytes, dst):
    """PRESET_COUNTER form of CTR: ctr16 is counter block 0 verbatim"""
    return check(core().uaes_ctr_crypt_block(bits, key, ctr16, first_block, _ptr(src), nbytes, _ptr(dst)))


def ecb_encrypt_padded(bits, key, src, nbytes, dst, padding):
    return check(core().uaes_ecb_encrypt_padded(bits, key, _ptr(src), nbytes, _ptr(dst), padding))


def xts_crypt_range(bits, keys, tweak, first_block, src, nbytes, dst, encrypt=True):
    """a block range of one XTS data unit (tweak chain entered by jump-ahead)"""
    return check(core().uaes_xts_crypt_range(bits, keys, tweak, first_block, _ptr(src), nbytes, _ptr(dst),
                                             1 if encrypt else 0))


def gcm_encrypt_ex(bits, key, nonce, aad, src, nbytes, dst, taglen=16):
    return check(core().uaes_gcm_encrypt_ex(bits, key, nonce, len(nonce), _ptr(aad), len(aad) if aad else 0,
                                            _ptr(src), nbytes, _ptr(dst), taglen))


def gcm_decrypt_ex(bits, key, nonce, aad, src, nbytes, dst, taglen=16):
    return check(core().uaes_gcm_decrypt_ex(bits, key, nonce, len(nonce), _ptr(aad), len(aad) if aad else 0,
                                            _ptr(src), nbytes, _ptr(dst), taglen))


def aead_ex(mode, bits, key, nonce, aad, src, nbytes, dst, taglen, encrypt=True):
    """uaes_{ocb,ccm,eax}_{en,de}crypt_ex: one message with a tag of `taglen` bytes; returns the result code"""
    f = getattr(core(), f"uaes_{mode}_{'encrypt' if encrypt else 'decrypt'}_ex")
    return check(f(bits, key, nonce, _ptr(aad) if aad else None, len(aad) if aad else 0, _ptr(src), nbytes, _ptr(dst), taglen))


def cbc_decrypt_ex(bits, key, iv, src, nbytes, dst, cts=True):
    return check(core().uaes_cbc_decrypt_ex(bits, key, iv, _ptr(src), nbytes, _ptr(dst), 1 if cts else 0))


def set_devices(n):
    """GPUs a host-buffer call may be spread over (n <= 0: all); returns the number in effect"""
    return core().uaes_set_devices(n)


def set_fanout_min(nbytes):
    core().uaes_set_fanout_min(nbytes)


def set_staging(chunk_bytes=0, slots=0):
    """staging chunk size / slots per device for host buffers (releases the current ones first)"""
    core().uaes_set_staging(chunk_bytes, slots)


def set_copy_threads(n):
    core().uaes_set_copy_threads(n)


def set_burn(flag):
    core().uaes_set_burn(1 if flag else 0)


def trim():
    core().uaes_trim()


def shutdown():
    core().uaes_shutdown()


def fill_splitmix64(seed, first_word, dst, nwords):
    return check(core().uaes_fill_splitmix64(seed, first_word, _ptr(dst), nwords))


def xor_fold64(src, nwords):
    r = _u64(0)
    check(core().uaes_xor_fold64(_ptr(src), nwords, ctypes.byref(r)))
    return r.value


def set_stream(stream_handle):
    core().uaes_set_stream(stream_handle)


def set_async(flag):
    core().uaes_set_async(1 if flag else 0)


class Msg(ctypes.Structure):
    """uaes_msg of include/uaes_b200.h: one message of a batch"""
    _fields_ = [("in_off", ctypes.c_uint64), ("out_off", ctypes.c_uint64), ("aad_off", ctypes.c_uint64),
                ("len", ctypes.c_uint32), ("aad_len", ctypes.c_uint32), ("nonce", ctypes.c_uint8 * 16),
                ("result", ctypes.c_int32), ("reserved", ctypes.c_uint32)]


def ccm_batch(bits, key, msgs, n, aad, src, dst, decrypt=False, mode="ccm"):
    """uaes_{ccm,eax,siv}_{en,de}crypt_batch; msgs = ctypes array of Msg or a device pointer.  Returns
    the call's result code (0, or 0x1A when a message failed authentication); raises on UAES_E_*."""
    f = getattr(core(), f"uaes_{mode}_{'decrypt' if decrypt else 'encrypt'}_batch")
    rc = f(bits, key, _ptr(msgs), n, _ptr(aad), _ptr(src), _ptr(dst))
    if rc < 0:
        check(rc)
    return rc


class Stream:
    """uaes_stream_* of include/uaes_b200.h: one CTR or GCM message fed in pieces"""

    def __init__(self, bits, key, nonce, aad=None, decrypt=False, gcm=True):
        c = core()
        self.h = c.uaes_stream_gcm(bits, key, nonce, _ptr(aad) if aad else None, len(aad) if aad else 0,
                                   1 if decrypt else 0) if gcm else c.uaes_stream_ctr(bits, key, nonce)
        if not self.h:
            raise UaesError(core().uaes_last_error_string().decode())

    def update(self, src, nbytes, dst):
        return check(core().uaes_stream_update(self.h, _ptr(src), nbytes, _ptr(dst)))

    def final(self, tag=None):
        """encrypt: returns the 16-byte tag; decrypt: pass the received tag, returns 0 or 0x1A"""
        if tag is None:
            t = ctypes.create_string_buffer(16)
            check(core().uaes_stream_final(self.h, ctypes.addressof(t)))
            return t.raw
        return check(core().uaes_stream_final(self.h, _ptr(tag)))

    def close(self):
        if self.h:
            core().uaes_stream_free(self.h)
            self.h = None


def kernel_launches():
    return core().uaes_kernel_launches()


def ctr_queue_stats():
    """(units served by table-driven warps, by bitsliced warps, blocks per unit) of this thread's last
    work-queue CTR launch"""
    a, b, c = _u64(0), _u64(0), _u64(0)
    check(core().uaes_ctr_queue_stats(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
    return a.value, b.value, c.value


def ctr_tuning(tt_threads=-1, bs_permille=-1, bs_min_blocks=-1):
    """CTR kernel geometry (uaes_ctr_tuning in include/uaes_b200.h); negative = leave unchanged"""
    core().uaes_ctr_tuning(tt_threads, bs_permille, bs_min_blocks)
