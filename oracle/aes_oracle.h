/*
 * aes_oracle.h -- CPU oracle for the AES ECB/CTR/XTS/GCM hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the algorithms in
 * polfosol/micro-AES (micro_aes.c) for the one hot path this repository
 * accelerates.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product library (libuaes_b200.so) never
 * links, loads or calls anything in oracle/.
 *
 * Parity is PINNED: tests/test_oracle.py checks every function below against
 *   - the reference's own known-answer vectors (main.c, testvectors/ .rsp files;
 *     committed as tests/golden/ JSON files by tests/golden/make_golden.py), and
 *   - the unmodified reference compiled from /root/reference into oracle/_ref/
 *     (see oracle/Makefile), on random inputs.
 *
 * Unlike the reference, key length is a run-time argument (keybits = 128|192|256;
 * the reference fixes it with the AES___ macro, micro_aes.h:17) and nothing is
 * kept in static storage (the reference keeps RoundKey static, micro_aes.c:72),
 * so tests may call it from several threads/processes.
 */
#ifndef AES_ORACLE_H_
#define AES_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* return codes: same values as micro_aes.h:469-476 */
enum {
    ORACLE_SUCCESS              = 0,
    ORACLE_DATALENGTH_ERROR     = 0x1,
    ORACLE_AUTHENTICATION_ERROR = 0x1A,
    ORACLE_DECRYPTION_ERROR     = 0x1D
};

/* FIPS-197 single block (micro_aes.c:242-259 / 315-332) */
void oracle_encrypt_block(int keybits, const uint8_t *key, const uint8_t in[16], uint8_t out[16]);
void oracle_decrypt_block(int keybits, const uint8_t *key, const uint8_t in[16], uint8_t out[16]);
/* key schedule, FIPS-197 order, 16*(rounds+1) bytes (micro_aes.c:144-178); returns rounds */
int  oracle_key_expansion(int keybits, const uint8_t *key, uint8_t *roundkeys);

/* ECB with the default zero padding, out holds ceil16(len) (micro_aes.c:636-680) */
void oracle_ecb_encrypt(int keybits, const uint8_t *key, const void *in, size_t len, void *out);
int  oracle_ecb_decrypt(int keybits, const uint8_t *key, const void *in, size_t len, void *out);

/* the reference's compile-time variants of this path as run-time arguments (micro_aes.h:56, 78-80,
 * 97-110): AES_PADDING 1 / 2, PRESET_COUNTER, GCM_NONCE_LEN / GCM_TAG_LEN, CTS = 0; each is pinned on
 * a build of the unmodified reference with the same macro (oracle/Makefile, _ref/libref128*.so) */
void oracle_ecb_encrypt_padded(int keybits, const uint8_t *key, const void *in, size_t len, void *out,
                               int padding);
void oracle_ctr_crypt_block(int keybits, const uint8_t *key, const uint8_t ctr0[16],
                            uint64_t first_block, const void *in, size_t len, void *out);
void oracle_gcm_encrypt_ex(int keybits, const uint8_t *key, const uint8_t *nonce, size_t noncelen,
                           const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen);
int  oracle_gcm_decrypt_ex(int keybits, const uint8_t *key, const uint8_t *nonce, size_t noncelen,
                           const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen);
int  oracle_cbc_decrypt_nocts(int keybits, const uint8_t *key, const uint8_t iv[16],
                              const void *in, size_t len, void *out);
/* CCM_TAG_LEN (even, 4..16), EAX_TAG_LEN, OCB_TAG_LEN (1..16) as arguments (micro_aes.c:1229, 1281, 1308, 1594,
 * 1638, 1707, 1783, 1807); pinned on oracle/_ref/libref128atag.so (8 / 10 / 12 bytes) */
void oracle_ccm_encrypt_ex(int keybits, const uint8_t *key, const uint8_t nonce[11],
                           const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen);
int  oracle_ccm_decrypt_ex(int keybits, const uint8_t *key, const uint8_t nonce[11],
                           const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen);
void oracle_eax_encrypt_ex(int keybits, const uint8_t *key, const uint8_t nonce[16],
                           const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen);
int  oracle_eax_decrypt_ex(int keybits, const uint8_t *key, const uint8_t nonce[16],
                           const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen);
void oracle_ocb_encrypt_ex(int keybits, const uint8_t *key, const uint8_t nonce[12],
                           const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen);
int  oracle_ocb_decrypt_ex(int keybits, const uint8_t *key, const uint8_t nonce[12],
                           const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen);
/* blocks [first_block, ...) of one XTS data unit: the chain of micro_aes.c:1030-1036 started late */
int  oracle_xts_range(int keybits, const uint8_t *keys, const uint8_t *tweak, uint64_t first_block,
                      const void *in, size_t len, void *out, int encrypt);

/* CTR, 12-byte IV, counter starts at 1, 56-bit big-endian carry (micro_aes.c:919-976).
 * first_block skips that many keystream blocks (counter-range extension). */
void oracle_ctr_crypt(int keybits, const uint8_t *key, const uint8_t iv[12],
                      const void *in, size_t len, void *out);
void oracle_ctr_crypt_at(int keybits, const uint8_t *key, const uint8_t iv[12],
                         uint64_t first_block, const void *in, size_t len, void *out);

/* XTS, one data unit per call, keys = K1 || K2, tweak NULL -> sector 0
 * (micro_aes.c:1008-1093) */
int  oracle_xts_encrypt(int keybits, const uint8_t *keys, const uint8_t *tweak,
                        const void *in, size_t len, void *out);
int  oracle_xts_decrypt(int keybits, const uint8_t *keys, const uint8_t *tweak,
                        const void *in, size_t len, void *out);
/* caller loop over fixed-size sectors, tweak_j = LE128(first_sector + j)
 * (the copyLint convention of micro_aes.c:1017-1021).  len % sector_bytes == 0. */
int  oracle_xts_sectors(int keybits, const uint8_t *keys, uint64_t first_sector,
                        size_t sector_bytes, const void *in, size_t len, void *out,
                        int encrypt);

/* GCM, 12-byte nonce, 16-byte tag appended at out+len (micro_aes.c:1127-1212) */
void oracle_gcm_encrypt(int keybits, const uint8_t *key, const uint8_t nonce[12],
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out);
int  oracle_gcm_decrypt(int keybits, const uint8_t *key, const uint8_t nonce[12],
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out);
/* GHASH(H; aad, ct) including the length block (micro_aes.c:1127-1137) */
void oracle_ghash(const uint8_t H[16], const void *aad, size_t aadlen,
                  const void *ct, size_t ctlen, uint8_t out[16]);
/* state <- xMac(data) with mix = mulGF128(H, .): the absorb loop alone (micro_aes.c:551-570),
 * so a long message can be hashed in pieces and recombined with powers of H */
void oracle_ghash_absorb(const uint8_t H[16], const void *data, size_t len, uint8_t state[16]);
/* y <- x*y in GCM's GF(2^128) (micro_aes.c:476-493) */
void oracle_gf128_mul(const uint8_t x[16], uint8_t y[16]);
/* T <- alpha*T in XTS's GF(2^128) (micro_aes.c:449-458) */
void oracle_xts_double(uint8_t t[16]);

/* ---- SURVEY.md 8f "next" row 1: AES-GCM-SIV (RFC 8452), micro_aes.c:1418-1516 ---- */
/* POLYVAL(H; aad, pt) including the little-endian length block (micro_aes.c:1423-1434) */
void oracle_polyval(const uint8_t H[16], const void *aad, size_t aadlen,
                    const void *pt, size_t ptlen, uint8_t out[16]);
/* y <- dot(x, y), POLYVAL's field product (micro_aes.c:511-528) */
void oracle_dot128(const uint8_t x[16], uint8_t y[16]);
/* 12-byte nonce, 16-byte tag appended at out+len */
void oracle_gcmsiv_encrypt(int keybits, const uint8_t *key, const uint8_t nonce[12],
                           const void *aad, size_t aadlen, const void *in, size_t len, void *out);
/* in holds len+16; like the reference the plaintext is written BEFORE the tag is checked */
int  oracle_gcmsiv_decrypt(int keybits, const uint8_t *key, const uint8_t nonce[12],
                           const void *aad, size_t aadlen, const void *in, size_t len, void *out);

/* ---- SURVEY.md 8f "next" row 2: the block-parallel DEcrypt directions of CBC (with the
 * reference's default CS3 ciphertext stealing, micro_aes.c:746-782) and CFB (micro_aes.c:799-845).
 * The encrypt directions are serial chains and stay out of scope; oracle_cbc_encrypt /
 * oracle_cfb_encrypt exist only so that tests can build valid ciphertexts (micro_aes.c:697-733). */
int  oracle_cbc_decrypt(int keybits, const uint8_t *key, const uint8_t iv[16],
                        const void *in, size_t len, void *out);
int  oracle_cbc_encrypt(int keybits, const uint8_t *key, const uint8_t iv[16],
                        const void *in, size_t len, void *out);
void oracle_cfb_decrypt(int keybits, const uint8_t *key, const uint8_t iv[16],
                        const void *in, size_t len, void *out);
void oracle_cfb_encrypt(int keybits, const uint8_t *key, const uint8_t iv[16],
                        const void *in, size_t len, void *out);

/* ---- SURVEY.md 8f "next" row 3: AES-OCB (RFC 7253, 12-byte nonce, 16-byte tag),
 * micro_aes.c:1655-1814 ---- */
void oracle_ocb_encrypt(int keybits, const uint8_t *key, const uint8_t nonce[12],
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out);
/* like the reference, the plaintext is written before the tag is checked */
int  oracle_ocb_decrypt(int keybits, const uint8_t *key, const uint8_t nonce[12],
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out);

/* CCM with the reference's default parameters: 11-byte nonce, 16-byte tag
 * (micro_aes.c:1268-1314, micro_aes.h:104-105); out holds len + 16 on encrypt */
void oracle_ccm_encrypt(int keybits, const uint8_t *key, const uint8_t nonce[11],
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out);
int  oracle_ccm_decrypt(int keybits, const uint8_t *key, const uint8_t nonce[11],
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out);

/* EAX (not EAX'): 16-byte nonce, 16-byte tag (micro_aes.c:1564-1648, micro_aes.h:119-121) */
void oracle_eax_encrypt(int keybits, const uint8_t *key, const uint8_t nonce[16],
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out);
int  oracle_eax_decrypt(int keybits, const uint8_t *key, const uint8_t nonce[16],
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out);

/* SIV (RFC 5297) with one AAD unit: keys = K1 || K2 (micro_aes.c:1372-1410) */
void oracle_siv_encrypt(int keybits, const uint8_t *keys, const void *aad, size_t aadlen,
                        const void *in, size_t len, uint8_t iv[16], void *out);
int  oracle_siv_decrypt(int keybits, const uint8_t *keys, const uint8_t iv[16], const void *aad, size_t aadlen,
                        const void *in, size_t len, void *out);

/* splitmix64 synthetic-data generator shared by tests and bench: 64-bit word w of
 * the buffer (byte offset 8w, little-endian) = splitmix64(seed + first_word + w) */
void oracle_fill_splitmix64(uint64_t seed, uint64_t first_word, void *dst, size_t nwords);

#ifdef __cplusplus
}
#endif
#endif
