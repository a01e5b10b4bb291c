#!/usr/bin/env python3
"""Generate tests/golden/*.json from the reference's own known-answer material.

Run in the BUILD container only (needs /root/reference); the outputs are committed
so that the GPU box, which has no /root/reference, can run every parity test.

Sources (all under /root/reference):
  main.c:16-88           the 57-byte demo vectors, for ECB / CTR / XTS / GCM
  main.c:19-21           FIPS-197 C.1 hidden in the key constants (main.c:124-134)
  testvectors/XTSGenAES{128,256}.rsp    filtered as aes_testvectors_XTS.h:84 does
                         (whole-byte data units only) -> 800 / 600 cases
  testvectors/GcmEncryptExtIV{128,192,256}.rsp  filtered as aes_testvectors_GCM.h:86
                         does (IVlen = 96, Taglen = 128) -> 375 cases each

Additionally writes oracle_ref_samples.json: outputs of the UNMODIFIED reference
(oracle/_ref/libref*.so) on seeded inputs that no in-tree vector pins (long CTR,
counter carries, multi-sector XTS, XTS-256 stealing, big GCM) so that the oracle
stays pinned to the reference on the GPU box as well.
"""
import ctypes
import hashlib
import json
import os
import re
import sys

REF = os.environ.get("UAES_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def c_strings(src):
    """name -> concatenated hex of every `*name = "..." "..."` initialiser, in order of
    appearance (a name can appear several times under different #if branches)."""
    out = {}
    for m in re.finditer(r'\*(\w+)\s*=\s*((?:"[^"]*"\s*(?:/\*.*?\*/\s*)*|#.*\n\s*)+)', src):
        name, body = m.group(1), m.group(2)
        # split the initialiser on preprocessor lines so #if/#else alternatives stay apart
        parts, cur = [], []
        for line in body.splitlines():
            if line.strip().startswith("#"):
                if cur:
                    parts.append("".join(cur))
                    cur = []
                continue
            cur += re.findall(r'"([^"]*)"', line)
        if cur:
            parts.append("".join(cur))
        out.setdefault(name, []).extend(p.replace(" ", "").lower() for p in parts)
    return out


def main_c_vectors():
    src = open(os.path.join(REF, "main.c"), encoding="utf-8").read()
    s = c_strings(src)
    key_pool = s["cipherKey"][0] + s["secondKey"][0]          # main.c:119-120
    pt = s["plainText"][0]
    assert len(pt) == 114
    # order of appearance in main.c: AES-256 block first, then AES-128, then AES-192
    gcm256, gcm128 = s["gcmcipher"][0], s["gcmcipher"][1]
    xts256, xts128 = s["xtscipher"][0], s["xtscipher"][1]
    ecb128 = s["ecbcipher"][0]
    ctr128_preset, ctr128 = s["ctrcipher"][0], s["ctrcipher"][1]
    assert len(ecb128) == 128 and len(ctr128) == 114 and len(gcm128) == 114 + 32
    return {
        "source": "main.c:16-88 (57-byte demo vectors) and main.c:19-21 (FIPS-197 C.1)",
        "plaintext": pt,
        "iv16": s["iVec"][0],
        "key_pool": key_pool,
        "aad": s["secretKey"][0][2:],                          # authKey + 1, 31 bytes
        "fips197_c1": {"key": s["secretKey"][0][:32], "pt": s["secondKey"][0][:32],
                       "ct": s["cipherKey"][0][32:]},
        "ecb128": ecb128, "ctr128": ctr128, "ctr128_preset_counter": ctr128_preset,
        "xts128": xts128, "xts256": xts256, "gcm128": gcm128, "gcm256": gcm256,
        # OCB, main.c:69-71 (same inputs as GCM) and the RFC 7253 case main.c:262-274
        "ocb128": s["ocbcipher"][0],
        # CCM, main.c:55-57, 198-204 (nonce = first 11 bytes of iVec, same AAD)
        "ccm128": s["ccmcipher"][0],
        # EAX (not EAX'): main.c:68-70, 225-237 (nonce = iVec, 16 bytes); SIV: main.c:52-54, 212-218 (keys = first
        # 32 bytes of the key pool, output = IV || ciphertext) and the two extra cases main.c:300-321
        "eax128": s["eaxcipher"][0], "siv128": s["sivcipher"][0],
        "siv_extra": [
            {"keys": "fffefdfcfbfaf9f8f7f6f5f4f3f2f1f0f0f1f2f3f4f5f6f7f8f9fafbfcfdfeff",
             "aad": "101112131415161718191a1b1c1d1e1f2021222324252627", "pt": "112233445566778899aabbccddee",
             "out": "85632d07c6e8f37f950acd320a2ecc9340c02b9690c4dc04daef7f6afe5c"},
            {"keys": "fffefdfcfbfaf9f8f7f6f5f4f3f2f1f0f0f1f2f3f4f5f6f7f8f9fafbfcfdfeff", "aad": "",
             "pt": "00112233445566778899aabbccddeeff",
             "out": "f304f912863e303d5b540e5057c7010c942ffaf45b0e5ca5fb9a56a5263bb065"}],
        "ocb_rfc7253": {"key": "000102030405060708090a0b0c0d0e0f", "iv": "bbaa99887766554433221107",
                        "aad": "000102030405060708090a0b0c0d0e0f1011121314151617",
                        "pt": "000102030405060708090a0b0c0d0e0f1011121314151617",
                        "ct": "1ca2207308c87c010756104d8840ce1952f09673a448a122c92c62241051f57356d7f3c90bb0e07f"},
        # CBC with CS3 stealing (CTS = 1, main.c:35-37,146-152) and CFB (main.c:41-42,153-159)
        "cbc128_cts": s["cbccipher"][0] + s["cbccipher"][1], "cfb128": s["cfbcipher"][0],
        # GCM-SIV, main.c:61-63 (same inputs) and the two RFC 8452 cases main.c:275-299
        "gcmsiv128": s["gsvcipher"][0],
        "gcmsiv_rfc8452": [
            {"key": "ee8e1ed9ff2540ae8f2ba9f50bc2f27c", "iv": "752abad3e0afb5f434dc4310", "aad": "6578616d706c65",
             "pt": "48656c6c6f20776f726c64", "ct": "5d349ead175ef6b1def6fd4fbcdeb7e4793f4a1d7e4faa70100af1"},
            {"key": "01000000000000000000000000000000", "iv": "030000000000000000000000", "aad": "01",
             "pt": "0200000000000000000000000000000003000000000000000000000000000000",
             "ct": "620048ef3c1e73e57e02bb8562c416a319e73e4caac8e96a1ecb2933145a1d71e6af6a7f87287da059a71684ed3498e1"}],
    }


def parse_xts(path, keybits):
    cases, cur, direction = [], {}, None
    for line in open(path):
        line = line.strip()
        if line in ("[ENCRYPT]", "[DECRYPT]"):
            direction = line[1:-1].lower()
        elif " = " in line or line.endswith(" ="):
            k, _, v = line.partition(" =")
            cur[k.strip()] = v.strip()
            if "PT" in cur and "CT" in cur:
                if (len(cur["Key"]) == keybits // 4
                        and int(cur["DataUnitLen"]) == 4 * len(cur["PT"])):
                    cases.append({"dir": direction, "key": cur["Key"], "i": cur["i"],
                                  "pt": cur["PT"], "ct": cur["CT"]})
                cur = {}
    return cases


def parse_gcm(path, keybits):
    cases, cur, hdr = [], {}, {}
    for line in open(path):
        line = line.strip()
        m = re.match(r"\[(\w+) = (\d+)\]", line)
        if m:
            hdr[m.group(1)] = int(m.group(2))
        elif " = " in line or line.endswith(" ="):
            k, _, v = line.partition(" =")
            cur[k.strip()] = v.strip()
            if "Tag" in cur:
                if hdr["Keylen"] == keybits and hdr["IVlen"] == 96 and hdr["Taglen"] == 128:
                    cases.append({"key": cur["Key"], "iv": cur["IV"], "pt": cur["PT"],
                                  "aad": cur["AAD"], "ct": cur["CT"], "tag": cur["Tag"]})
                cur = {}
    return cases


def parse_gcm_variants(path, keybits, per_group=2):
    """the groups of GcmEncryptExtIV*.rsp the default build skips: IVlen = 8 / 1024 bits (J0 = GHASH of
    the nonce, micro_aes.c:1145-1149) and Taglen < 128 (truncated tag, micro_aes.c:1178); the first and
    last case of every [Keylen, IVlen, PTlen, AADlen, Taglen] group"""
    groups, cur, hdr = [], {}, {}
    for line in open(path):
        line = line.strip()
        m = re.match(r"\[(\w+) = (\d+)\]", line)
        if m:
            if m.group(1) == "Keylen":
                groups.append([])
            hdr[m.group(1)] = int(m.group(2))
        elif " = " in line or line.endswith(" ="):
            k, _, v = line.partition(" =")
            cur[k.strip()] = v.strip()
            if "Tag" in cur:
                if hdr["Keylen"] == keybits and (hdr["IVlen"] != 96 or hdr["Taglen"] != 128):
                    groups[-1].append({"key": cur["Key"], "iv": cur["IV"], "pt": cur["PT"],
                                       "aad": cur["AAD"], "ct": cur["CT"], "tag": cur["Tag"]})
                cur = {}
    cases = []
    for g in groups:
        pick = g[:1] + (g[-1:] if len(g) > 1 and per_group > 1 else [])
        cases += pick
    return cases


def variant_samples():
    """outputs of the unmodified reference BUILT WITH ITS OTHER COMPILE-TIME SETTINGS
    (oracle/_ref/libref128{pc,iv1,iv128,tag12,pad1,pad2,cts0}.so, see oracle/Makefile) on seeded inputs"""
    L = lambda v: ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", f"libref128{v}.so"))
    sha = lambda b: hashlib.sha256(b).hexdigest()
    out = {"source": "oracle/_ref/libref128<variant>.so = unmodified micro_aes.c with one edited header line each",
           "ctr_preset_counter": [], "gcm_nonce": [], "gcm_tag12": [], "ecb_padding": [], "cbc_nocts": []}
    pc = L("pc")
    for n in (0, 1, 16, 57, 4096 + 5, 1 << 18):
        key, ctr, pt = rnd(f"vpk{n}", 16), rnd(f"vpc{n}", 16), rnd(f"vpp{n}", n)
        ct = ctypes.create_string_buffer(n + 16)
        pc.AES_CTR_encrypt(key, ctr, pt, ctypes.c_size_t(n), ct)
        out["ctr_preset_counter"].append({"n": n, "key": key.hex(), "counter0": ctr.hex(), "pt_tag": f"vpp{n}",
                                          "ct_sha256": sha(ct.raw[:n])})
    civ = L("civ8")                        # CTR_IV_LENGTH = 8, CTR_START_VALUE = 0x01020304 (micro_aes.c:967-971)
    out["ctr_iv8_start"] = []
    for n in (0, 1, 16, 57, 4096 + 5, 1 << 18):
        key, iv, pt = rnd(f"vik{n}", 16), rnd(f"vii{n}", 8), rnd(f"vip{n}", n)
        ct = ctypes.create_string_buffer(n + 16)
        civ.AES_CTR_encrypt(key, iv, pt, ctypes.c_size_t(n), ct)
        out["ctr_iv8_start"].append({"n": n, "key": key.hex(), "iv": iv.hex(), "start": 0x01020304, "pt_tag": f"vip{n}",
                                     "ct_sha256": sha(ct.raw[:n])})
    for v, ivlen in (("iv1", 1), ("iv128", 128)):
        lib = L(v)
        lib.AES_GCM_decrypt.restype = ctypes.c_char
        for n, a in ((0, 0), (1, 0), (16, 16), (57, 31), (4096 + 3, 129), (1 << 18, 7)):
            key, nonce = rnd(f"vgk{v}{n}", 16), rnd(f"vgn{v}{n}", ivlen)
            aad, pt = rnd(f"vga{v}{n}", a), rnd(f"vgp{v}{n}", n)
            ct = ctypes.create_string_buffer(n + 16)
            lib.AES_GCM_encrypt(key, nonce, aad, ctypes.c_size_t(a), pt, ctypes.c_size_t(n), ct)
            back = ctypes.create_string_buffer(n + 16)
            assert ord(lib.AES_GCM_decrypt(key, nonce, aad, ctypes.c_size_t(a), ct, ctypes.c_size_t(n), back)) == 0
            assert back.raw[:n] == pt
            out["gcm_nonce"].append({"noncelen": ivlen, "n": n, "aadlen": a, "key": key.hex(), "nonce": nonce.hex(),
                                     "aad_tag": f"vga{v}{n}", "pt_tag": f"vgp{v}{n}",
                                     "ct_sha256": sha(ct.raw[:n]), "tag": ct.raw[n:n + 16].hex()})
    t12 = L("tag12")
    t12.AES_GCM_decrypt.restype = ctypes.c_char
    for n, a in ((0, 0), (1, 5), (57, 31), (4096 + 3, 129)):
        key, nonce = rnd(f"vtk{n}", 16), rnd(f"vtn{n}", 12)
        aad, pt = rnd(f"vta{n}", a), rnd(f"vtp{n}", n)
        ct = ctypes.create_string_buffer(b"\xee" * (n + 16), n + 16)
        t12.AES_GCM_encrypt(key, nonce, aad, ctypes.c_size_t(a), pt, ctypes.c_size_t(n), ct)
        assert ct.raw[n + 12:n + 16] == b"\xee" * 4          # only 12 tag bytes are written
        bad = bytearray(ct.raw[:n + 12]); bad[-1] ^= 1
        rc_bad = ord(t12.AES_GCM_decrypt(key, nonce, aad, ctypes.c_size_t(a), bytes(bad), ctypes.c_size_t(n),
                                         ctypes.create_string_buffer(n + 16)))
        out["gcm_tag12"].append({"n": n, "aadlen": a, "key": key.hex(), "nonce": nonce.hex(), "aad_tag": f"vta{n}",
                                 "pt_tag": f"vtp{n}", "ct_sha256": sha(ct.raw[:n]), "tag": ct.raw[n:n + 12].hex(),
                                 "rc_forged": rc_bad})
    at = L("atag")                         # CCM_TAG_LEN = 8, EAX_TAG_LEN = 10, OCB_TAG_LEN = 12 in one build
    out["aead_tags"] = []
    for name, fn, dfn, nl, tl in (("ccm", at.AES_CCM_encrypt, at.AES_CCM_decrypt, 11, 8), ("eax", at.AES_EAX_encrypt, at.AES_EAX_decrypt, 16, 10),
                                  ("ocb", at.AES_OCB_encrypt, at.AES_OCB_decrypt, 12, 12)):
        dfn.restype = ctypes.c_char
        for n, a in ((0, 0), (1, 7), (16, 16), (57, 31), (4096 + 3, 129)):
            key, nonce = rnd(f"vak{name}{n}", 16), rnd(f"van{name}{n}", nl)
            aad, pt = rnd(f"vaa{name}{n}", a), rnd(f"vap{name}{n}", n)
            ct = ctypes.create_string_buffer(b"\xee" * (n + 16), n + 16)
            fn(key, nonce, aad, ctypes.c_size_t(a), pt, ctypes.c_size_t(n), ct)
            assert ct.raw[n + tl:n + 16] == b"\xee" * (16 - tl)
            back = ctypes.create_string_buffer(n + 16)
            assert ord(dfn(key, nonce, aad, ctypes.c_size_t(a), ct, ctypes.c_size_t(n), back)) == 0 and back.raw[:n] == pt
            bad = bytearray(ct.raw[:n + tl]); bad[-1] ^= 2
            rc_bad = ord(dfn(key, nonce, aad, ctypes.c_size_t(a), bytes(bad), ctypes.c_size_t(n), ctypes.create_string_buffer(n + 16)))
            out["aead_tags"].append({"mode": name, "taglen": tl, "n": n, "aadlen": a, "key": key.hex(), "nonce": nonce.hex(),
                                     "aad_tag": f"vaa{name}{n}", "pt_tag": f"vap{name}{n}", "ct_sha256": sha(ct.raw[:n]),
                                     "tag": ct.raw[n:n + tl].hex(), "rc_forged": rc_bad})
    for v, mode in (("pad1", 1), ("pad2", 2)):
        lib = L(v)
        for n in (0, 1, 15, 16, 17, 32, 57, 4096, 4096 + 7):
            key, pt = rnd(f"vek{v}{n}", 16), rnd(f"vep{v}{n}", n)
            m = (n // 16 + 1) * 16
            ct = ctypes.create_string_buffer(b"\xee" * (m + 16), m + 16)
            lib.AES_ECB_encrypt(key, pt, ctypes.c_size_t(n), ct)
            assert ct.raw[m:m + 16] == b"\xee" * 16
            out["ecb_padding"].append({"padding": mode, "n": n, "key": key.hex(), "pt_tag": f"vep{v}{n}",
                                       "ct_sha256": sha(ct.raw[:m])})
    c0 = L("cts0")
    c0.AES_CBC_decrypt.restype = ctypes.c_char
    for n in (0, 15, 16, 17, 32, 48, 57, 4096, 1 << 18):
        key, iv, ct = rnd(f"vck{n}", 16), rnd(f"vci{n}", 16), rnd(f"vcc{n}", n)
        o = ctypes.create_string_buffer(n + 16)
        rc = ord(c0.AES_CBC_decrypt(key, iv, ct, ctypes.c_size_t(n), o))
        out["cbc_nocts"].append({"n": n, "key": key.hex(), "iv": iv.hex(), "ct_tag": f"vcc{n}", "rc": rc,
                                 "pt_sha256": sha(o.raw[:n]) if rc == 0 else None})
    return out


def parse_gcmsiv(path, keybits):
    """testvectors/SIV_GCM_ACVP.tv, kept when the key has the build's size
    (aes_testvectors_GCMSIV.h:84)"""
    cases, cur = [], {}
    for line in open(path):
        line = line.strip()
        if " = " in line or line.endswith(" ="):
            k, _, v = line.partition(" =")
            cur[k.strip()] = v.strip().lower()
            if "pt" in cur and "ct" in cur and "key" in cur and "iv" in cur and "aad" in cur:
                if len(cur["key"]) == keybits // 4:
                    cases.append({k2: cur[k2] for k2 in ("key", "iv", "aad", "pt", "ct")})
                cur = {}
    return cases


def parse_ocb(path, keybits):
    """testvectors/OCB_AES128.tv (OpenSSL evp test format): blank-line separated cases; kept when the
    key has the build's size, the IV 12 bytes and the tag 16 bytes and no error is expected
    (aes_testvectors_OCB.h:88-93)"""
    cases, cur = [], {}
    for line in list(open(path)) + [""]:
        line = line.strip()
        if line.startswith("#"):
            continue
        if not line:
            if cur.get("Cipher", "").lower().endswith("-ocb") and "Result" not in cur:
                if len(cur["Key"]) == keybits // 4 and len(cur["IV"]) == 24 and len(cur["Tag"]) == 32:
                    cases.append({"key": cur["Key"].lower(), "iv": cur["IV"].lower(), "aad": cur.get("AAD", "").lower(),
                                  "pt": cur.get("Plaintext", "").lower(),
                                  "ct": (cur.get("Ciphertext", "") + cur["Tag"]).lower()})
            cur = {}
        elif " = " in line or line.endswith(" ="):
            k, _, v = line.partition(" =")
            cur[k.strip()] = v.strip()
    return cases


def parse_ccm(path, keybits):
    """testvectors/VNT{128,192,256}.rsp (CAVS CCM variable-nonce test): kept when the nonce has the
    build's CCM_NONCE_LEN = 11 bytes and the tag CCM_TAG_LEN = 16 (aes_testvectors_CCM.h:97)"""
    cases, key, cur = [], None, {}
    for line in open(path):
        line = line.strip()
        if line.startswith("Key = "):
            key = line.split(" = ")[1].lower()
        for name in ("Nonce", "Adata", "Payload", "CT"):
            if line.startswith(name + " = "):
                cur[name] = line.split(" = ")[1].lower()
        if "Payload" in cur and "CT" in cur:
            if len(key) == keybits // 4 and len(cur["Nonce"]) == 22 and len(cur["CT"]) == len(cur["Payload"]) + 32:
                cases.append({"key": key, "nonce": cur["Nonce"], "aad": cur["Adata"], "pt": cur["Payload"], "ct": cur["CT"]})
            cur = {}
    return cases


def parse_eax(path, keybits):
    """testvectors/EAX_AES128.tv (Bellare-Rogaway-Wagner): MSG/KEY/NONCE/HEADER/CIPHER groups, kept when
    key and nonce have the build's sizes (aes_testvectors_EAX.h:83)"""
    cases, cur = [], {}
    for line in open(path):
        line = line.strip()
        for name in ("MSG", "KEY", "NONCE", "HEADER", "CIPHER"):
            if line.startswith(name + ":"):
                cur[name] = line.split(":", 1)[1].strip().lower()
        if "CIPHER" in cur:
            if len(cur["KEY"]) == keybits // 4 and len(cur["NONCE"]) == 32 and len(cur["CIPHER"]) == len(cur["MSG"]) + 32:
                cases.append({"key": cur["KEY"], "nonce": cur["NONCE"], "aad": cur["HEADER"], "pt": cur["MSG"], "ct": cur["CIPHER"]})
            cur = {}
    return cases


def row4_samples():
    """SURVEY 8f row 4 (CCM now): outputs of the unmodified reference on inputs no in-tree vector
    covers -- empty/ragged payloads, AAD around the 14-byte first block and the 0xFEFF length-encoding
    switch (micro_aes.c:1236-1240), all key sizes"""
    libs = {b: ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", f"libref{b}.so")) for b in (128, 192, 256)}
    sha = lambda b: hashlib.sha256(b).hexdigest()
    out = {"source": "oracle/_ref/libref*.so = unmodified /root/reference/micro_aes.c; inputs = rnd()", "ccm": []}
    for bits, lib in libs.items():
        for n, a in ((0, 0), (0, 9), (1, 0), (15, 13), (16, 14), (17, 15), (57, 31), (1000, 20), (4096 + 3, 129),
                     (100, 65279), (33, 65280), (64, 70000 + 5), (1 << 16, 7)):
            key, nonce = rnd(f"cck{bits}{n}", bits // 8), rnd(f"ccn{bits}{n}", 11)
            aad, pt = rnd(f"cca{bits}{n}{a}", a), rnd(f"ccp{bits}{n}", n)
            ct = ctypes.create_string_buffer(n + 16)
            lib.AES_CCM_encrypt(key, nonce, aad, ctypes.c_size_t(a), pt, ctypes.c_size_t(n), ct)
            out["ccm"].append({"bits": bits, "n": n, "aadlen": a, "key": key.hex(), "nonce": nonce.hex(),
                               "aad_tag": f"cca{bits}{n}{a}", "pt_tag": f"ccp{bits}{n}",
                               "ct_sha256": sha(ct.raw[:n]), "tag": ct.raw[n:n + 16].hex()})
    out["eax"], out["siv"] = [], []
    for bits, lib in libs.items():
        for n, a in ((0, 0), (0, 9), (1, 0), (15, 13), (16, 16), (17, 15), (32, 0), (57, 31), (1000, 20), (4096 + 3, 129), (1 << 16, 7)):
            key, nonce = rnd(f"eak{bits}{n}", bits // 8), rnd(f"ean{bits}{n}", 16)
            aad, pt = rnd(f"eaa{bits}{n}{a}", a), rnd(f"eap{bits}{n}", n)
            ct = ctypes.create_string_buffer(n + 16)
            lib.AES_EAX_encrypt(key, nonce, aad, ctypes.c_size_t(a), pt, ctypes.c_size_t(n), ct)
            out["eax"].append({"bits": bits, "n": n, "aadlen": a, "key": key.hex(), "nonce": nonce.hex(),
                               "aad_tag": f"eaa{bits}{n}{a}", "pt_tag": f"eap{bits}{n}",
                               "ct_sha256": sha(ct.raw[:n]), "tag": ct.raw[n:n + 16].hex()})
            keys = rnd(f"sik{bits}{n}", bits // 4)
            iv, ct = ctypes.create_string_buffer(16), ctypes.create_string_buffer(n + 16)
            lib.AES_SIV_encrypt(keys, aad, ctypes.c_size_t(a), pt, ctypes.c_size_t(n), iv, ct)
            out["siv"].append({"bits": bits, "n": n, "aadlen": a, "keys": keys.hex(), "aad_tag": f"eaa{bits}{n}{a}",
                               "pt_tag": f"eap{bits}{n}", "iv": iv.raw.hex(), "ct_sha256": sha(ct.raw[:n])})
    return out


def rnd(tag, n):
    """deterministic pseudo-random bytes: SHA-256 in counter mode over a tag"""
    out = b""
    i = 0
    while len(out) < n:
        out += hashlib.sha256(f"{tag}:{i}".encode()).digest()
        i += 1
    return out[:n]


def ref_samples():
    """outputs of the unmodified reference on inputs no in-tree vector covers"""
    libs = {b: ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", f"libref{b}.so"))
            for b in (128, 192, 256)}
    pc = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref128pc.so"))
    out = {"source": "oracle/_ref/libref*.so = unmodified /root/reference/micro_aes.c, "
                     "gcc -O2 -fno-strict-aliasing; inputs = SHA-256 counter streams (rnd())",
           "ctr": [], "ctr_preset_counter": [], "ecb": [], "xts": [], "xts_sectors": [],
           "gcm": [], "gcmsiv": []}
    sha = lambda b: hashlib.sha256(b).hexdigest()
    for bits, lib in libs.items():
        ks = bits // 8
        for n in (0, 1, 15, 16, 17, 255, 4096, 65536 + 5, 1 << 20):
            key, iv, pt = rnd(f"ctrk{bits}{n}", ks), rnd(f"ctri{bits}{n}", 12), rnd(f"ctrp{bits}{n}", n)
            ct = ctypes.create_string_buffer(n + 16)
            lib.AES_CTR_encrypt(key, iv, pt, ctypes.c_size_t(n), ct)
            out["ctr"].append({"bits": bits, "n": n, "key": key.hex(), "iv": iv.hex(),
                               "pt_tag": f"ctrp{bits}{n}", "ct_sha256": sha(ct.raw[:n]),
                               "ct_head": ct.raw[:min(n, 64)].hex()})
        for n in (0, 5, 16, 48, 100, 4096 + 7):
            key, pt = rnd(f"ecbk{bits}{n}", ks), rnd(f"ecbp{bits}{n}", n)
            m = (n + 15) // 16 * 16
            ct = ctypes.create_string_buffer(m + 16)
            lib.AES_ECB_encrypt(key, pt, ctypes.c_size_t(n), ct)
            out["ecb"].append({"bits": bits, "n": n, "key": key.hex(), "pt_tag": f"ecbp{bits}{n}",
                               "ct_sha256": sha(ct.raw[:m])})
        if bits != 192:      # XTS-192 is not defined (testvectors/aes_testvectors.h:57-62)
            for n in (16, 17, 31, 32, 33, 57, 512, 4096, 4096 + 9, 65536 + 15):
                keys, tw, pt = rnd(f"xtsk{bits}{n}", 2 * ks), rnd(f"xtst{bits}{n}", 16), rnd(f"xtsp{bits}{n}", n)
                ct = ctypes.create_string_buffer(n + 16)
                rc = lib.AES_XTS_encrypt(keys, tw, pt, ctypes.c_size_t(n), ct)
                assert rc == 0
                out["xts"].append({"bits": bits, "n": n, "keys": keys.hex(), "tweak": tw.hex(),
                                   "pt_tag": f"xtsp{bits}{n}", "ct_sha256": sha(ct.raw[:n])})
            for first, sb, ns in ((0, 512, 40), ((1 << 32) - 3, 512, 8), (7, 4096, 3), (123456789012, 528, 5)):
                keys, pt = rnd(f"xsk{bits}{first}", 2 * ks), rnd(f"xsp{bits}{first}", sb * ns)
                ct = ctypes.create_string_buffer(sb * ns)
                for j in range(ns):
                    tw = (first + j).to_bytes(16, "little")
                    o = ctypes.create_string_buffer(sb)
                    lib.AES_XTS_encrypt(keys, tw, pt[j * sb:(j + 1) * sb], ctypes.c_size_t(sb), o)
                    ct[j * sb:(j + 1) * sb] = o.raw
                out["xts_sectors"].append({"bits": bits, "first_sector": first, "sector_bytes": sb,
                                           "sectors": ns, "keys": keys.hex(),
                                           "pt_tag": f"xsp{bits}{first}", "ct_sha256": sha(ct.raw)})
        for n, a in ((0, 0), (1, 0), (16, 16), (57, 31), (1000, 20), (4096, 0), (65536 + 3, 129), (1 << 18, 7)):
            key, nonce = rnd(f"gcmk{bits}{n}", ks), rnd(f"gcmn{bits}{n}", 12)
            aad, pt = rnd(f"gcma{bits}{n}", a), rnd(f"gcmp{bits}{n}", n)
            ct = ctypes.create_string_buffer(n + 16)
            lib.AES_GCM_encrypt(key, nonce, aad, ctypes.c_size_t(a), pt, ctypes.c_size_t(n), ct)
            out["gcm"].append({"bits": bits, "n": n, "aadlen": a, "key": key.hex(), "nonce": nonce.hex(),
                               "aad_tag": f"gcma{bits}{n}", "pt_tag": f"gcmp{bits}{n}",
                               "ct_sha256": sha(ct.raw[:n]), "tag": ct.raw[n:n + 16].hex()})
    for bits, lib in libs.items():
        ks = bits // 8
        for n, a in ((0, 0), (1, 0), (16, 16), (57, 31), (1000, 20), (4096 + 3, 129), (1 << 18, 7)):
            key, nonce = rnd(f"gsk{bits}{n}", ks), rnd(f"gsn{bits}{n}", 12)
            aad, pt = rnd(f"gsa{bits}{n}", a), rnd(f"gsp{bits}{n}", n)
            ct = ctypes.create_string_buffer(n + 16)
            lib.GCM_SIV_encrypt(key, nonce, aad, ctypes.c_size_t(a), pt, ctypes.c_size_t(n), ct)
            out["gcmsiv"].append({"bits": bits, "n": n, "aadlen": a, "key": key.hex(), "nonce": nonce.hex(),
                                  "aad_tag": f"gsa{bits}{n}", "pt_tag": f"gsp{bits}{n}",
                                  "ct_sha256": sha(ct.raw[:n]), "tag": ct.raw[n:n + 16].hex()})
    out["ocb"] = []
    for bits, lib in libs.items():
        for n, a in ((0, 0), (0, 5), (1, 0), (16, 16), (57, 31), (1000, 20), (4096 + 3, 129), (1 << 18, 7), (100, 70000 + 5)):
            key, nonce = rnd(f"ock{bits}{n}", bits // 8), rnd(f"ocn{bits}{n}", 12)
            aad, pt = rnd(f"oca{bits}{n}{a}", a), rnd(f"ocp{bits}{n}", n)
            ct = ctypes.create_string_buffer(n + 16)
            lib.AES_OCB_encrypt(key, nonce, aad, ctypes.c_size_t(a), pt, ctypes.c_size_t(n), ct)
            out["ocb"].append({"bits": bits, "n": n, "aadlen": a, "key": key.hex(), "nonce": nonce.hex(),
                               "aad_tag": f"oca{bits}{n}{a}", "pt_tag": f"ocp{bits}{n}",
                               "ct_sha256": sha(ct.raw[:n]), "tag": ct.raw[n:n + 16].hex()})
    out["cbc_decrypt"], out["cfb_decrypt"] = [], []
    for bits, lib in libs.items():
        lib.AES_CBC_decrypt.restype = ctypes.c_char
        for n in (15, 16, 17, 31, 32, 33, 48, 57, 4096, 4096 + 5, 65536 + 15, 1 << 18):
            key, iv, ct = rnd(f"cbk{bits}{n}", bits // 8), rnd(f"cbi{bits}{n}", 16), rnd(f"cbc{bits}{n}", n)
            o = ctypes.create_string_buffer(n + 16)
            rc = lib.AES_CBC_decrypt(key, iv, ct, ctypes.c_size_t(n), o)
            out["cbc_decrypt"].append({"bits": bits, "n": n, "key": key.hex(), "iv": iv.hex(), "ct_tag": f"cbc{bits}{n}",
                                       "rc": ord(rc), "pt_sha256": sha(o.raw[:n]) if ord(rc) == 0 else None})
        for n in (0, 1, 16, 17, 57, 4096, 4096 + 5, 1 << 18):
            key, iv, ct = rnd(f"cfk{bits}{n}", bits // 8), rnd(f"cfi{bits}{n}", 16), rnd(f"cfc{bits}{n}", n)
            o = ctypes.create_string_buffer(n + 16)
            lib.AES_CFB_decrypt(key, iv, ct, ctypes.c_size_t(n), o)
            out["cfb_decrypt"].append({"bits": bits, "n": n, "key": key.hex(), "iv": iv.hex(), "ct_tag": f"cfc{bits}{n}",
                                       "pt_sha256": sha(o.raw[:n])})
    # The 32-bit little-endian counter of GCM-SIV wraps (micro_aes.c:935-938).  The counter starts at
    # the tag, and GCM_SIV_decrypt runs CTR with the RECEIVED tag before it authenticates
    # (micro_aes.c:1505-1515), so a forged tag ff ff ff f8 .. pins the wrap: the call fails with
    # M_AUTHENTICATION_ERROR but the output buffer holds input ^ keystream across the wrap.
    out["gcmsiv_forged_tag_decrypt"] = []
    for bits, lib in libs.items():
        lib.GCM_SIV_decrypt.restype = ctypes.c_char
        key, nonce, ct = rnd(f"gsfk{bits}", bits // 8), rnd(f"gsfn{bits}", 12), rnd(f"gsfc{bits}", 16 * 40 + 3)
        tag = bytes.fromhex("f8ffffff") + rnd(f"gsft{bits}", 12)
        o = ctypes.create_string_buffer(len(ct) + 16)
        rc = lib.GCM_SIV_decrypt(key, nonce, b"", ctypes.c_size_t(0), ct + tag, ctypes.c_size_t(len(ct)), o)
        out["gcmsiv_forged_tag_decrypt"].append({"bits": bits, "key": key.hex(), "nonce": nonce.hex(),
                                                 "ct_tag": f"gsfc{bits}", "n": len(ct), "tag": tag.hex(),
                                                 "rc": ord(rc), "out_sha256": sha(o.raw[:len(ct)])})
    # counter carries: the PRESET_COUNTER build takes the caller's 16-byte block as counter 0
    for name, ctr_hex in (("byte15", "00112233445566778899aabbccddeeff"),
                          ("into_nonce_byte11", "000102030405060708090a0bfffffffe"),
                          ("wrap56", "a0a1a2a3a4a5a6a7a8fffffffffffffd")):
        key, pt = rnd("pck" + name, 16), rnd("pcp" + name, 16 * 8)
        ct = ctypes.create_string_buffer(16 * 8)
        pc.AES_CTR_encrypt(key, bytes.fromhex(ctr_hex), pt, ctypes.c_size_t(16 * 8), ct)
        out["ctr_preset_counter"].append({"name": name, "key": key.hex(), "counter0": ctr_hex,
                                          "pt_tag": "pcp" + name, "ct": ct.raw.hex()})
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not found: golden vectors can only be regenerated in the build container")
    w = lambda name, obj: json.dump(obj, open(os.path.join(HERE, name), "w"), indent=0, separators=(",", ":"))
    w("main_c.json", main_c_vectors())
    tv = os.path.join(REF, "testvectors")
    for bits in (128, 256):
        c = parse_xts(os.path.join(tv, f"XTSGenAES{bits}.rsp"), 2 * bits)
        w(f"xts{bits}.json", {"source": f"testvectors/XTSGenAES{bits}.rsp, filter of aes_testvectors_XTS.h:84",
                              "cases": c})
        print(f"xts{bits}: {len(c)} cases")
    for bits in (128, 192, 256):
        c = parse_gcm(os.path.join(tv, f"GcmEncryptExtIV{bits}.rsp"), bits)
        w(f"gcm{bits}.json", {"source": f"testvectors/GcmEncryptExtIV{bits}.rsp, filter of aes_testvectors_GCM.h:86",
                              "cases": c})
        print(f"gcm{bits}: {len(c)} cases")
    c = parse_gcmsiv(os.path.join(tv, "SIV_GCM_ACVP.tv"), 128)
    w("gcmsiv128.json", {"source": "testvectors/SIV_GCM_ACVP.tv (102 AES-128 cases, aes_testvectors_GCMSIV.h)",
                         "cases": c})
    print(f"gcmsiv128: {len(c)} cases")
    c = parse_ocb(os.path.join(tv, "OCB_AES128.tv"), 128)
    ref128 = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref128.so"))
    for case in c:                      # the file is only trusted through the reference itself
        pt, aad = bytes.fromhex(case["pt"]), bytes.fromhex(case["aad"])
        o = ctypes.create_string_buffer(len(pt) + 16)
        ref128.AES_OCB_encrypt(bytes.fromhex(case["key"]), bytes.fromhex(case["iv"]), aad, ctypes.c_size_t(len(aad)),
                               pt, ctypes.c_size_t(len(pt)), o)
        assert o.raw.hex() == case["ct"], case
    w("ocb128.json", {"source": "testvectors/OCB_AES128.tv, cases the reference harness runs "
                                "(aes_testvectors_OCB.h:88-93); each re-checked against oracle/_ref/libref128.so",
                      "cases": c})
    print(f"ocb128: {len(c)} cases")
    w("oracle_ref_samples.json", ref_samples())
    for bits in (128, 192, 256):
        c = parse_ccm(os.path.join(tv, f"VNT{bits}.rsp"), bits)
        w(f"ccm{bits}.json", {"source": f"testvectors/VNT{bits}.rsp, filter of aes_testvectors_CCM.h:97 (11-byte nonce, 16-byte tag)",
                              "cases": c})
        print(f"ccm{bits}: {len(c)} cases")
    c = parse_eax(os.path.join(tv, "EAX_AES128.tv"), 128)
    w("eax128.json", {"source": "testvectors/EAX_AES128.tv, filter of aes_testvectors_EAX.h:83", "cases": c})
    print(f"eax128: {len(c)} cases")
    w("oracle_ref_samples_row4.json", row4_samples())
    import gzip
    allv = {"source": "testvectors/GcmEncryptExtIV{128,192,256}.rsp: the groups with IVlen != 96 or Taglen != 128 "
                      "(GCM_NONCE_LEN / GCM_TAG_LEN variants, micro_aes.c:1145-1149, 1178), first and last case of each group"}
    for bits in (128, 192, 256):
        allv[str(bits)] = parse_gcm_variants(os.path.join(tv, f"GcmEncryptExtIV{bits}.rsp"), bits)
        print(f"gcm variants {bits}: {len(allv[str(bits)])} cases")
    with gzip.GzipFile(os.path.join(HERE, "gcm_variants.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(allv, separators=(",", ":")).encode())
    w("oracle_ref_variant_samples.json", variant_samples())
    print("ok")


if __name__ == "__main__":
    main()
