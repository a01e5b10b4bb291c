/*
 * aes_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see aes_oracle.h).
 *
 * A from-scratch restatement of the hot path of polfosol/micro-AES: the Rijndael
 * block cipher and the ECB / CTR / XTS / GCM chaining loops.  Each function cites
 * the micro_aes.c lines whose behaviour it restates.  It is written for clarity
 * and for being pinned against the reference, not for speed; the only liberties
 * taken are a run-time key length and no static state.
 */
#include "aes_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* GF(2^8) and the S-boxes (values of micro_aes.c:41-65, derived not copied) */
/* ------------------------------------------------------------------------ */

static uint8_t SBOX[256], INV_SBOX[256];

/* multiply by x modulo x^8+x^4+x^3+x+1 (micro_aes.c:115-118 `xtime`) */
static uint8_t gf8_double(uint8_t a)
{
    return (uint8_t)((a << 1) ^ ((a >> 7) * 0x1b));
}

static uint8_t gf8_mul(uint8_t a, uint8_t b)
{
    uint8_t r = 0;
    while (b) {
        if (b & 1) r ^= a;
        a = gf8_double(a);
        b >>= 1;
    }
    return r;
}

/* FIPS-197 5.1.1: S(a) = affine(a^-1) */
__attribute__((constructor)) static void build_sboxes(void)
{
    int a;
    for (a = 0; a < 256; ++a) {
        uint8_t inv = 0, s;
        int b;
        if (a)
            for (b = 1; b < 256; ++b)
                if (gf8_mul((uint8_t)a, (uint8_t)b) == 1) { inv = (uint8_t)b; break; }
        s = inv;
        s ^= (uint8_t)((inv << 1) | (inv >> 7));
        s ^= (uint8_t)((inv << 2) | (inv >> 6));
        s ^= (uint8_t)((inv << 3) | (inv >> 5));
        s ^= (uint8_t)((inv << 4) | (inv >> 4));
        s ^= 0x63;
        SBOX[a] = s;
        INV_SBOX[s] = (uint8_t)a;
    }
}

/* ------------------------------------------------------------------------ */
/* key schedule and the block cipher                                        */
/* ------------------------------------------------------------------------ */

typedef struct {
    int rounds;             /* Nk + 6  (micro_aes.c:16-27) */
    uint8_t rk[240];        /* round r at offset 16*r, FIPS-197 byte order */
} aes_ctx;

/* micro_aes.c:144-178 (KeyExpansion) */
static void key_setup(aes_ctx *c, int keybits, const uint8_t *key)
{
    const int nk = keybits == 256 ? 8 : keybits == 192 ? 6 : 4;
    const int total = 4 * (nk + 7);        /* 4 * (rounds + 1) words */
    uint8_t rcon = 1;
    int i;

    c->rounds = nk + 6;
    memcpy(c->rk, key, (size_t)(4 * nk));
    for (i = nk; i < total; ++i) {
        uint8_t t[4];
        memcpy(t, c->rk + 4 * (i - 1), 4);
        if (i % nk == 0) {
            /* RotWord, SubWord, rcon */
            const uint8_t t0 = t[0];
            t[0] = SBOX[t[1]] ^ rcon;
            t[1] = SBOX[t[2]];
            t[2] = SBOX[t[3]];
            t[3] = SBOX[t0];
            rcon = gf8_double(rcon);
        } else if (nk == 8 && i % nk == 4) {
            /* AES-256 extra SubWord (micro_aes.c:165-172) */
            t[0] = SBOX[t[0]]; t[1] = SBOX[t[1]];
            t[2] = SBOX[t[2]]; t[3] = SBOX[t[3]];
        }
        c->rk[4 * i + 0] = c->rk[4 * (i - nk) + 0] ^ t[0];
        c->rk[4 * i + 1] = c->rk[4 * (i - nk) + 1] ^ t[1];
        c->rk[4 * i + 2] = c->rk[4 * (i - nk) + 2] ^ t[2];
        c->rk[4 * i + 3] = c->rk[4 * (i - nk) + 3] ^ t[3];
    }
}

static void xor16(uint8_t *dst, const uint8_t *src)   /* micro_aes.c:105-112 */
{
    int i;
    for (i = 0; i < 16; ++i) dst[i] ^= src[i];
}

/* state byte i = column i/4, row i%4 (micro_aes.c:74-77) */
static void shift_rows(uint8_t s[16], int inverse)     /* micro_aes.c:198-218, 278-298 */
{
    uint8_t t[16];
    int col, row;
    for (col = 0; col < 4; ++col)
        for (row = 0; row < 4; ++row) {
            const int from = inverse ? (col - row + 4) % 4 : (col + row) % 4;
            t[4 * col + row] = s[4 * from + row];
        }
    memcpy(s, t, 16);
}

static void mix_columns(uint8_t s[16])                 /* micro_aes.c:221-239 */
{
    int c;
    for (c = 0; c < 16; c += 4) {
        const uint8_t a0 = s[c], a1 = s[c + 1], a2 = s[c + 2], a3 = s[c + 3];
        const uint8_t all = a0 ^ a1 ^ a2 ^ a3;
        s[c + 0] = a0 ^ all ^ gf8_double(a0 ^ a1);
        s[c + 1] = a1 ^ all ^ gf8_double(a1 ^ a2);
        s[c + 2] = a2 ^ all ^ gf8_double(a2 ^ a3);
        s[c + 3] = a3 ^ all ^ gf8_double(a3 ^ a0);
    }
}

static void inv_mix_columns(uint8_t s[16])             /* micro_aes.c:301-312 */
{
    int c;
    for (c = 0; c < 16; c += 4) {
        const uint8_t a0 = s[c], a1 = s[c + 1], a2 = s[c + 2], a3 = s[c + 3];
        s[c + 0] = gf8_mul(a0, 14) ^ gf8_mul(a1, 11) ^ gf8_mul(a2, 13) ^ gf8_mul(a3, 9);
        s[c + 1] = gf8_mul(a0, 9) ^ gf8_mul(a1, 14) ^ gf8_mul(a2, 11) ^ gf8_mul(a3, 13);
        s[c + 2] = gf8_mul(a0, 13) ^ gf8_mul(a1, 9) ^ gf8_mul(a2, 14) ^ gf8_mul(a3, 11);
        s[c + 3] = gf8_mul(a0, 11) ^ gf8_mul(a1, 13) ^ gf8_mul(a2, 9) ^ gf8_mul(a3, 14);
    }
}

/* micro_aes.c:242-259 (rijndaelEncrypt); in may alias out */
static void encrypt_block(const aes_ctx *c, const uint8_t *in, uint8_t *out)
{
    uint8_t s[16];
    int r, i;
    memcpy(s, in, 16);
    for (r = 0; r < c->rounds; ++r) {
        xor16(s, c->rk + 16 * r);
        for (i = 0; i < 16; ++i) s[i] = SBOX[s[i]];
        shift_rows(s, 0);
        if (r + 1 < c->rounds) mix_columns(s);
    }
    xor16(s, c->rk + 16 * c->rounds);
    memcpy(out, s, 16);
}

/* micro_aes.c:315-332 (rijndaelDecrypt): the straight inverse cipher */
static void decrypt_block(const aes_ctx *c, const uint8_t *in, uint8_t *out)
{
    uint8_t s[16];
    int r, i;
    memcpy(s, in, 16);
    xor16(s, c->rk + 16 * c->rounds);
    for (r = c->rounds - 1; r >= 0; --r) {
        shift_rows(s, 1);
        for (i = 0; i < 16; ++i) s[i] = INV_SBOX[s[i]];
        xor16(s, c->rk + 16 * r);
        if (r) inv_mix_columns(s);
    }
    memcpy(out, s, 16);
}

int oracle_key_expansion(int keybits, const uint8_t *key, uint8_t *roundkeys)
{
    aes_ctx c;
    key_setup(&c, keybits, key);
    memcpy(roundkeys, c.rk, (size_t)(16 * (c.rounds + 1)));
    return c.rounds;
}

void oracle_encrypt_block(int keybits, const uint8_t *key, const uint8_t in[16], uint8_t out[16])
{
    aes_ctx c;
    key_setup(&c, keybits, key);
    encrypt_block(&c, in, out);
}

void oracle_decrypt_block(int keybits, const uint8_t *key, const uint8_t in[16], uint8_t out[16])
{
    aes_ctx c;
    key_setup(&c, keybits, key);
    decrypt_block(&c, in, out);
}

/* ------------------------------------------------------------------------ */
/* ECB                                                                      */
/* ------------------------------------------------------------------------ */

/* micro_aes.c:636-653 with AES_PADDING == 0: the partial tail block is
 * zero-filled and encrypted (padBlock, micro_aes.c:610-621) */
void oracle_ecb_encrypt(int keybits, const uint8_t *key, const void *in, size_t len, void *out)
{
    aes_ctx c;
    const uint8_t *x = (const uint8_t *)in;
    uint8_t *y = (uint8_t *)out;
    size_t n = len / 16, tail = len % 16;

    key_setup(&c, keybits, key);
    for (; n--; x += 16, y += 16) encrypt_block(&c, x, y);
    if (tail) {
        uint8_t last[16] = {0};
        memcpy(last, x, tail);
        encrypt_block(&c, last, y);
    }
}

/* micro_aes.c:663-680: full blocks are decrypted, the tail bytes are copied
 * through untouched (the initial memcpy), and a ragged length is an error */
/* micro_aes.c:610-621 (padBlock) with AES_PADDING = 1 (PKCS#7) or 2 (ISO/IEC 7816-4): the last block
 * is ALWAYS padded, so out holds (len / 16 + 1) * 16 bytes; padding = 0 is oracle_ecb_encrypt */
void oracle_ecb_encrypt_padded(int keybits, const uint8_t *key, const void *in, size_t len, void *out,
                               int padding)
{
    aes_ctx c;
    const uint8_t *x = (const uint8_t *)in;
    uint8_t *y = (uint8_t *)out, last[16];
    size_t n = len / 16, tail = len % 16, i;

    if (!padding) { oracle_ecb_encrypt(keybits, key, in, len, out); return; }
    key_setup(&c, keybits, key);
    for (; n--; x += 16, y += 16) encrypt_block(&c, x, y);
    for (i = 0; i < 16; ++i)
        last[i] = i < tail ? x[i] : padding == 1 ? (uint8_t)(16 - tail) : i == tail ? 0x80 : 0;
    encrypt_block(&c, last, y);
}

int oracle_ecb_decrypt(int keybits, const uint8_t *key, const void *in, size_t len, void *out)
{
    aes_ctx c;
    const uint8_t *x = (const uint8_t *)in;
    uint8_t *y = (uint8_t *)out;
    size_t n = len / 16, tail = len % 16;

    key_setup(&c, keybits, key);
    for (; n--; x += 16, y += 16) decrypt_block(&c, x, y);
    if (tail) memmove(y, x, tail);
    return tail ? ORACLE_DECRYPTION_ERROR : ORACLE_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* CTR                                                                      */
/* ------------------------------------------------------------------------ */

/* counter += n on the 56-bit big-endian field in bytes 9..15; byte 8 and below
 * never change (incBlock with index = LAST, micro_aes.c:421-427) */
static void ctr_add(uint8_t ctr[16], uint64_t n)
{
    uint64_t v = 0;
    int i;
    for (i = 9; i < 16; ++i) v = v << 8 | ctr[i];
    v += n;                                   /* bits above 55 are discarded */
    for (i = 15; i >= 9; --i, v >>= 8) ctr[i] = (uint8_t)v;
}

/* the shared loop of micro_aes.c:943-949 (CTR_cipher): ctr is consumed */
static void ctr_stream(const aes_ctx *c, uint8_t ctr[16],
                       const uint8_t *x, size_t len, uint8_t *y)
{
    uint8_t ks[16];
    size_t i;
    for (; len >= 16; len -= 16, x += 16, y += 16) {
        encrypt_block(c, ctr, ks);
        for (i = 0; i < 16; ++i) y[i] = x[i] ^ ks[i];
        ctr_add(ctr, 1);
    }
    if (len) {                                /* mixThenXor, micro_aes.c:534-544 */
        encrypt_block(c, ctr, ks);
        for (i = 0; i < len; ++i) y[i] = x[i] ^ ks[i];
    }
}

/* micro_aes.c:962-976 with PRESET_COUNTER == 0, CTR_IV_LENGTH == 12,
 * CTR_START_VALUE == 1 */
void oracle_ctr_crypt_at(int keybits, const uint8_t *key, const uint8_t iv[12],
                         uint64_t first_block, const void *in, size_t len, void *out)
{
    aes_ctx c;
    uint8_t ctr[16] = {0};
    memcpy(ctr, iv, 12);
    ctr[15] ^= 1;                             /* xorBEint(ctr, 1, LAST), :971 */
    ctr_add(ctr, first_block);
    key_setup(&c, keybits, key);
    ctr_stream(&c, ctr, (const uint8_t *)in, len, (uint8_t *)out);
}

void oracle_ctr_crypt(int keybits, const uint8_t *key, const uint8_t iv[12],
                      const void *in, size_t len, void *out)
{
    oracle_ctr_crypt_at(keybits, key, iv, 0, in, len, out);
}

/* micro_aes.c:962-976 with PRESET_COUNTER == 1 (:964-966): the caller's 16 bytes are counter block 0 */
void oracle_ctr_crypt_block(int keybits, const uint8_t *key, const uint8_t ctr0[16],
                            uint64_t first_block, const void *in, size_t len, void *out)
{
    aes_ctx c;
    uint8_t ctr[16];
    memcpy(ctr, ctr0, 16);
    ctr_add(ctr, first_block);
    key_setup(&c, keybits, key);
    ctr_stream(&c, ctr, (const uint8_t *)in, len, (uint8_t *)out);
}

/* ------------------------------------------------------------------------ */
/* XTS                                                                      */
/* ------------------------------------------------------------------------ */

/* micro_aes.c:449-458 (doubleLblock): little-endian shift left, 0x87 fold */
void oracle_xts_double(uint8_t t[16])
{
    unsigned carry = 0;
    int i;
    for (i = 0; i < 16; ++i) {
        const unsigned v = (unsigned)t[i] << 1 | carry;
        carry = v >> 8;
        t[i] = (uint8_t)v;
    }
    if (carry) t[0] ^= 0x87;
}

static void xex_block(const aes_ctx *c, int encrypt, const uint8_t T[16],
                      const uint8_t *x, uint8_t *y)
{
    uint8_t b[16];
    memcpy(b, x, 16);
    xor16(b, T);
    if (encrypt) encrypt_block(c, b, b); else decrypt_block(c, b, b);
    xor16(b, T);
    memcpy(y, b, 16);
}

/* micro_aes.c:1008-1055 (XTS_cipher) */
static int xts_unit(int keybits, const uint8_t *keys, const uint8_t *tweak,
                    const uint8_t *x, size_t len, uint8_t *y, int encrypt)
{
    const int keysize = keybits / 8;
    aes_ctx k1, k2;
    uint8_t T[16] = {0};
    size_t r = len % 16, n;

    if (len < 16) return ORACLE_DATALENGTH_ERROR;       /* :1069, :1088 */
    n = len / 16 - (r > 0);
    if (tweak) memcpy(T, tweak, 16);                    /* NULL -> sector 0, :1017 */
    key_setup(&k2, keybits, keys + keysize);            /* key2 = second half */
    key_setup(&k1, keybits, keys);
    encrypt_block(&k2, T, T);

    for (; n--; x += 16, y += 16) {
        xex_block(&k1, encrypt, T, x, y);
        oracle_xts_double(T);
    }
    if (r) {                                            /* stealing, :1037-1053 */
        uint8_t Tn[16], first[16], cc[16], pp[16];
        size_t i;
        memcpy(Tn, T, 16);
        oracle_xts_double(Tn);
        /* encrypt: block m-1 under T, stolen block under alpha*T; decrypt: swapped */
        memcpy(first, x, 16);
        xex_block(&k1, encrypt, encrypt ? T : Tn, first, cc);
        for (i = 0; i < 16; ++i) pp[i] = i < r ? x[16 + i] : cc[i];
        for (i = 0; i < r; ++i) first[i] = cc[i];       /* keep before y is written */
        xex_block(&k1, encrypt, encrypt ? Tn : T, pp, y);
        memcpy(y + 16, first, r);
    }
    return ORACLE_SUCCESS;
}

/* A block range of one data unit: the blocks [first_block, first_block + len/16) of the chain of
 * micro_aes.c:1030-1036, entered by doubling T_0 first_block times (the reference has no such entry
 * point; this IS its loop started late).  A ragged len steals like the end of a unit. */
int oracle_xts_range(int keybits, const uint8_t *keys, const uint8_t *tweak, uint64_t first_block,
                     const void *in, size_t len, void *out, int encrypt)
{
    const int keysize = keybits / 8;
    const uint8_t *x = (const uint8_t *)in;
    uint8_t *y = (uint8_t *)out;
    aes_ctx k1, k2;
    uint8_t T[16] = {0};
    size_t r = len % 16, n;
    uint64_t j;

    if (len < 16) return ORACLE_DATALENGTH_ERROR;
    n = len / 16 - (r > 0);
    if (tweak) memcpy(T, tweak, 16);
    key_setup(&k2, keybits, keys + keysize);
    key_setup(&k1, keybits, keys);
    encrypt_block(&k2, T, T);
    for (j = 0; j < first_block; ++j) oracle_xts_double(T);
    for (; n--; x += 16, y += 16) {
        xex_block(&k1, encrypt, T, x, y);
        oracle_xts_double(T);
    }
    if (r) {
        uint8_t Tn[16], first[16], cc[16], pp[16];
        size_t i;
        memcpy(Tn, T, 16);
        oracle_xts_double(Tn);
        memcpy(first, x, 16);
        xex_block(&k1, encrypt, encrypt ? T : Tn, first, cc);
        for (i = 0; i < 16; ++i) pp[i] = i < r ? x[16 + i] : cc[i];
        for (i = 0; i < r; ++i) first[i] = cc[i];
        xex_block(&k1, encrypt, encrypt ? Tn : T, pp, y);
        memcpy(y + 16, first, r);
    }
    return ORACLE_SUCCESS;
}

int oracle_xts_encrypt(int keybits, const uint8_t *keys, const uint8_t *tweak,
                       const void *in, size_t len, void *out)
{
    return xts_unit(keybits, keys, tweak, (const uint8_t *)in, len, (uint8_t *)out, 1);
}

int oracle_xts_decrypt(int keybits, const uint8_t *keys, const uint8_t *tweak,
                       const void *in, size_t len, void *out)
{
    return xts_unit(keybits, keys, tweak, (const uint8_t *)in, len, (uint8_t *)out, 0);
}

int oracle_xts_sectors(int keybits, const uint8_t *keys, uint64_t first_sector,
                       size_t sector_bytes, const void *in, size_t len, void *out,
                       int encrypt)
{
    const uint8_t *x = (const uint8_t *)in;
    uint8_t *y = (uint8_t *)out;
    uint64_t s = first_sector;
    if (sector_bytes < 16 || len % sector_bytes) return ORACLE_DATALENGTH_ERROR;
    for (; len; len -= sector_bytes, x += sector_bytes, y += sector_bytes, ++s) {
        uint8_t tweak[16] = {0};
        int i;
        for (i = 0; i < 8; ++i) tweak[i] = (uint8_t)(s >> (8 * i));  /* copyLint, :399 */
        xts_unit(keybits, keys, tweak, x, sector_bytes, y, encrypt);
    }
    return ORACLE_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* GCM                                                                      */
/* ------------------------------------------------------------------------ */

/* micro_aes.c:464-473 (divideBblock): big-endian shift right, 0xE1 fold */
static void gcm_halve(uint8_t y[16])
{
    const unsigned lsb = y[15] & 1;
    int i;
    for (i = 15; i > 0; --i) y[i] = (uint8_t)(y[i] >> 1 | y[i - 1] << 7);
    y[0] >>= 1;
    if (lsb) y[0] ^= 0xe1;
}

/* micro_aes.c:476-493 (mulGF128): y <- x*y, bits of x MSB first */
void oracle_gf128_mul(const uint8_t x[16], uint8_t y[16])
{
    uint8_t acc[16] = {0};
    int i, b;
    for (i = 0; i < 16; ++i)
        for (b = 0x80; b; b >>= 1) {
            if (x[i] & b) xor16(acc, y);
            gcm_halve(y);
        }
    memcpy(y, acc, 16);
}

/* micro_aes.c:551-570 (xMac with mix = mulGF128) */
static void ghash_absorb(const uint8_t H[16], const uint8_t *x, size_t len, uint8_t g[16])
{
    size_t i;
    for (; len >= 16; len -= 16, x += 16) {
        xor16(g, x);
        oracle_gf128_mul(H, g);
    }
    if (len) {
        for (i = 0; i < len; ++i) g[i] ^= x[i];
        oracle_gf128_mul(H, g);
    }
}

void oracle_ghash_absorb(const uint8_t H[16], const void *data, size_t len, uint8_t state[16])
{
    ghash_absorb(H, (const uint8_t *)data, len, state);
}

/* micro_aes.c:1127-1137 (gHash) */
void oracle_ghash(const uint8_t H[16], const void *aad, size_t aadlen,
                  const void *ct, size_t ctlen, uint8_t out[16])
{
    uint8_t lens[16];
    uint64_t abits = (uint64_t)aadlen * 8, cbits = (uint64_t)ctlen * 8;
    int i;
    for (i = 7; i >= 0; --i, abits >>= 8) lens[i] = (uint8_t)abits;
    for (i = 15; i >= 8; --i, cbits >>= 8) lens[i] = (uint8_t)cbits;
    memset(out, 0, 16);
    ghash_absorb(H, (const uint8_t *)aad, aadlen, out);
    ghash_absorb(H, (const uint8_t *)ct, ctlen, out);
    ghash_absorb(H, lens, 16, out);
}

/* micro_aes.c:1140-1152 (GCMsetup) with GCM_NONCE_LEN == 12 */
static void gcm_setup(aes_ctx *c, int keybits, const uint8_t *key, const uint8_t nonce[12],
                      uint8_t H[16], uint8_t j0[16])
{
    key_setup(c, keybits, key);
    memset(H, 0, 16);
    encrypt_block(c, H, H);
    memcpy(j0, nonce, 12);
    j0[12] = j0[13] = j0[14] = 0;
    j0[15] = 1;
}

/* micro_aes.c:1164-1179 */
void oracle_gcm_encrypt(int keybits, const uint8_t *key, const uint8_t nonce[12],
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out)
{
    aes_ctx c;
    uint8_t H[16], j0[16], ctr[16], ekj0[16], g[16];
    gcm_setup(&c, keybits, key, nonce, H, j0);
    memcpy(ctr, j0, 16);
    ctr_add(ctr, 1);                          /* CCM_GCM pre-increment, :939-941 */
    ctr_stream(&c, ctr, (const uint8_t *)in, len, (uint8_t *)out);
    encrypt_block(&c, j0, ekj0);
    oracle_ghash(H, aad, aadlen, out, len, g);
    xor16(g, ekj0);
    memcpy((uint8_t *)out + len, g, 16);
}

/* micro_aes.c:1140-1152 with any GCM_NONCE_LEN: 12 bytes -> nonce || 00000001, otherwise
 * J0 = GHASH_H({}, nonce) (:1145-1149) */
static void gcm_setup_ex(aes_ctx *c, int keybits, const uint8_t *key, const uint8_t *nonce, size_t noncelen,
                         uint8_t H[16], uint8_t j0[16])
{
    if (noncelen == 12) { gcm_setup(c, keybits, key, nonce, H, j0); return; }
    key_setup(c, keybits, key);
    memset(H, 0, 16);
    encrypt_block(c, H, H);
    oracle_ghash(H, NULL, 0, nonce, noncelen, j0);
}

/* micro_aes.c:1164-1179 with GCM_NONCE_LEN = noncelen, GCM_TAG_LEN = taglen: out holds len + taglen */
void oracle_gcm_encrypt_ex(int keybits, const uint8_t *key, const uint8_t *nonce, size_t noncelen,
                           const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen)
{
    aes_ctx c;
    uint8_t H[16], j0[16], ctr[16], ekj0[16], g[16];
    gcm_setup_ex(&c, keybits, key, nonce, noncelen, H, j0);
    memcpy(ctr, j0, 16);
    ctr_add(ctr, 1);
    ctr_stream(&c, ctr, (const uint8_t *)in, len, (uint8_t *)out);
    encrypt_block(&c, j0, ekj0);
    oracle_ghash(H, aad, aadlen, out, len, g);
    xor16(g, ekj0);
    memcpy((uint8_t *)out + len, g, taglen);                  /* :1178 */
}

int oracle_gcm_decrypt_ex(int keybits, const uint8_t *key, const uint8_t *nonce, size_t noncelen,
                          const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen)
{
    aes_ctx c;
    uint8_t H[16], j0[16], ctr[16], ekj0[16], g[16];
    gcm_setup_ex(&c, keybits, key, nonce, noncelen, H, j0);
    oracle_ghash(H, aad, aadlen, in, len, g);
    encrypt_block(&c, j0, ekj0);
    xor16(g, ekj0);
    if (memcmp(g, (const uint8_t *)in + len, taglen)) return ORACLE_AUTHENTICATION_ERROR;   /* :1204 */
    memcpy(ctr, j0, 16);
    ctr_add(ctr, 1);
    ctr_stream(&c, ctr, (const uint8_t *)in, len, (uint8_t *)out);
    return ORACLE_SUCCESS;
}

/* micro_aes.c:1192-1212: verify first, leave `out` untouched on failure */
int oracle_gcm_decrypt(int keybits, const uint8_t *key, const uint8_t nonce[12],
                       const void *aad, size_t aadlen, const void *in, size_t len, void *out)
{
    aes_ctx c;
    uint8_t H[16], j0[16], ctr[16], ekj0[16], g[16];
    gcm_setup(&c, keybits, key, nonce, H, j0);
    oracle_ghash(H, aad, aadlen, in, len, g);
    encrypt_block(&c, j0, ekj0);
    xor16(g, ekj0);
    if (memcmp(g, (const uint8_t *)in + len, 16)) return ORACLE_AUTHENTICATION_ERROR;
    memcpy(ctr, j0, 16);
    ctr_add(ctr, 1);
    ctr_stream(&c, ctr, (const uint8_t *)in, len, (uint8_t *)out);
    return ORACLE_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* GCM-SIV                                                                  */
/* ------------------------------------------------------------------------ */

/* micro_aes.c:498-508 (divideLblock): little-endian shift right, 0xE1 into byte 15 */
static void polyval_halve(uint8_t y[16])
{
    const unsigned lsb = y[0] & 1;
    int i;
    for (i = 0; i < 15; ++i) y[i] = (uint8_t)(y[i] >> 1 | y[i + 1] << 7);
    y[15] >>= 1;
    if (lsb) y[15] ^= 0xe1;
}

/* micro_aes.c:511-528 (dotGF128): bytes of x from last to first, bits MSB first, halve BEFORE use */
void oracle_dot128(const uint8_t x[16], uint8_t y[16])
{
    uint8_t acc[16] = {0};
    int i, b;
    for (i = 15; i >= 0; --i)
        for (b = 0x80; b; b >>= 1) {
            polyval_halve(y);
            if (x[i] & b) xor16(acc, y);
        }
    memcpy(y, acc, 16);
}

static void polyval_absorb(const uint8_t H[16], const uint8_t *x, size_t len, uint8_t g[16])
{
    size_t i;
    for (; len >= 16; len -= 16, x += 16) {
        xor16(g, x);
        oracle_dot128(H, g);
    }
    if (len) {
        for (i = 0; i < len; ++i) g[i] ^= x[i];
        oracle_dot128(H, g);
    }
}

/* micro_aes.c:1423-1434 (polyval) */
void oracle_polyval(const uint8_t H[16], const void *aad, size_t aadlen,
                    const void *pt, size_t ptlen, uint8_t out[16])
{
    uint8_t lens[16];
    uint64_t abits = (uint64_t)aadlen * 8, pbits = (uint64_t)ptlen * 8;
    int i;
    for (i = 0; i < 8; ++i, abits >>= 8) lens[i] = (uint8_t)abits;       /* copyLint(len, .., 0)  */
    for (i = 8; i < 16; ++i, pbits >>= 8) lens[i] = (uint8_t)pbits;      /* copyLint(len, .., HB) */
    memset(out, 0, 16);
    polyval_absorb(H, (const uint8_t *)aad, aadlen, out);
    polyval_absorb(H, (const uint8_t *)pt, ptlen, out);
    polyval_absorb(H, lens, 16, out);
}

/* micro_aes.c:1437-1451 (GCM_SIVsetup): E_K(LE32(i) || nonce), first 8 bytes of each */
static void gcmsiv_setup(int keybits, const uint8_t *key, const uint8_t nonce[12],
                         uint8_t auth[16], aes_ctx *enc)
{
    aes_ctx master;
    uint8_t blk[16], out[16], derived[48];
    const int n = 2 + keybits / 64;
    int i;
    key_setup(&master, keybits, key);
    memset(blk, 0, 4);
    memcpy(blk + 4, nonce, 12);
    for (i = 0; i < n; ++i) {
        blk[0] = (uint8_t)i;
        encrypt_block(&master, blk, out);
        memcpy(derived + 8 * i, out, 8);
    }
    memcpy(auth, derived, 16);
    key_setup(enc, keybits, derived + 16);
}

/* micro_aes.c:1454-1462 (GCM_SIVtag) */
static void gcmsiv_tag(const aes_ctx *enc, const uint8_t nonce[12], uint8_t pv[16], uint8_t tag[16])
{
    int i;
    for (i = 0; i < 12; ++i) pv[i] ^= nonce[i];
    pv[15] &= 0x7f;
    encrypt_block(enc, pv, tag);
}

/* CTR_cipher with mode SIVGCM_CTR (micro_aes.c:935-938, 943-949): bit 7 of byte 15 set, the
 * counter is the little-endian 32-bit word in bytes 0..3 and wraps modulo 2^32 */
static void gcmsiv_ctr(const aes_ctx *enc, const uint8_t tag[16], const uint8_t *x, size_t len, uint8_t *y)
{
    uint8_t c[16], ks[16];
    uint32_t ctr;
    size_t i;
    memcpy(c, tag, 16);
    c[15] |= 0x80;
    ctr = (uint32_t)c[0] | (uint32_t)c[1] << 8 | (uint32_t)c[2] << 16 | (uint32_t)c[3] << 24;
    for (; len; x += 16, y += 16, ++ctr) {
        const size_t n = len < 16 ? len : 16;
        c[0] = (uint8_t)ctr; c[1] = (uint8_t)(ctr >> 8); c[2] = (uint8_t)(ctr >> 16); c[3] = (uint8_t)(ctr >> 24);
        encrypt_block(enc, c, ks);
        for (i = 0; i < n; ++i) y[i] = x[i] ^ ks[i];
        len -= n;
    }
}

/* micro_aes.c:1474-1487 */
void oracle_gcmsiv_encrypt(int keybits, const uint8_t *key, const uint8_t nonce[12],
                           const void *aad, size_t aadlen, const void *in, size_t len, void *out)
{
    aes_ctx enc;
    uint8_t auth[16], pv[16], tag[16];
    gcmsiv_setup(keybits, key, nonce, auth, &enc);
    oracle_polyval(auth, aad, aadlen, in, len, pv);
    gcmsiv_tag(&enc, nonce, pv, tag);
    gcmsiv_ctr(&enc, tag, (const uint8_t *)in, len, (uint8_t *)out);
    memcpy((uint8_t *)out + len, tag, 16);
}

/* micro_aes.c:1499-1516: decrypt first, then authenticate the plaintext */
int oracle_gcmsiv_decrypt(int keybits, const uint8_t *key, const uint8_t nonce[12],
                          const void *aad, size_t aadlen, const void *in, size_t len, void *out)
{
    aes_ctx enc;
    uint8_t auth[16], pv[16], tag[16], got[16];
    memcpy(got, (const uint8_t *)in + len, 16);
    gcmsiv_setup(keybits, key, nonce, auth, &enc);
    gcmsiv_ctr(&enc, got, (const uint8_t *)in, len, (uint8_t *)out);
    oracle_polyval(auth, aad, aadlen, out, len, pv);
    gcmsiv_tag(&enc, nonce, pv, tag);
    return memcmp(tag, got, 16) ? ORACLE_AUTHENTICATION_ERROR : ORACLE_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* CBC with CS3 stealing, CFB                                               */
/* ------------------------------------------------------------------------ */

/* micro_aes.c:697-733 (AES_CBC_encrypt, CTS == 1).  out must not alias in. */
int oracle_cbc_encrypt(int keybits, const uint8_t *key, const uint8_t iv[16],
                       const void *in, size_t len, void *out)
{
    aes_ctx c;
    const uint8_t *x = (const uint8_t *)in, *chain = iv;
    uint8_t *y = (uint8_t *)out;
    size_t n = len / 16, r = len % 16;
    if (n > 1 && !r) { --n; r = 16; }                  /* CS3: the last two blocks always swap */
    if (n == 0) return ORACLE_DATALENGTH_ERROR;
    key_setup(&c, keybits, key);
    for (; n--; x += 16, y += 16) {
        memcpy(y, x, 16);
        xor16(y, chain);
        encrypt_block(&c, y, y);
        chain = y;
    }
    if (r) {
        /* y - 16 holds C_(m-1); the final plaintext chunk P_m (r bytes) is zero padded, chained
         * with C_(m-1) and encrypted into the second-to-last position; C_(m-1)[0..r) moves last */
        uint8_t last[16] = {0}, head[16];
        memcpy(last, x, r);
        memcpy(head, y - 16, 16);
        xor16(last, head);
        encrypt_block(&c, last, y - 16);
        memcpy(y, head, r);
    }
    return ORACLE_SUCCESS;
}

/* micro_aes.c:746-782 built with CTS == 0: whole blocks only (:757-759), plain chain */
int oracle_cbc_decrypt_nocts(int keybits, const uint8_t *key, const uint8_t iv[16],
                             const void *in, size_t len, void *out)
{
    aes_ctx c;
    const uint8_t *x = (const uint8_t *)in, *chain = iv;
    uint8_t *y = (uint8_t *)out;
    size_t n = len / 16;
    if (len % 16) return ORACLE_DATALENGTH_ERROR;
    key_setup(&c, keybits, key);
    for (; n--; x += 16, y += 16) {
        decrypt_block(&c, x, y);
        xor16(y, chain);
        chain = x;
    }
    return ORACLE_SUCCESS;
}

/* micro_aes.c:746-782 (AES_CBC_decrypt, CTS == 1).  out must not alias in (the reference chains
 * through the input buffer, micro_aes.c:766). */
int oracle_cbc_decrypt(int keybits, const uint8_t *key, const uint8_t iv[16],
                       const void *in, size_t len, void *out)
{
    aes_ctx c;
    const uint8_t *x = (const uint8_t *)in, *chain = iv;
    uint8_t *y = (uint8_t *)out;
    size_t n = len / 16, r = len % 16, i;
    if (n > 1 && !r) { --n; r = 16; }
    if (n == 0) return ORACLE_DATALENGTH_ERROR;
    n -= r > 0;                                        /* the last two blocks are the CTS pair */
    key_setup(&c, keybits, key);
    for (; n--; x += 16, y += 16) {
        decrypt_block(&c, x, y);
        xor16(y, chain);
        chain = x;
    }
    if (r) {
        const uint8_t *z = x + 16;                     /* {X, Z}: X full, Z has r bytes */
        uint8_t dx[16], blk[16];
        decrypt_block(&c, x, dx);
        for (i = 0; i < r; ++i) y[16 + i] = dx[i] ^ z[i];      /* P2 = Z ^ Dec(X) */
        memcpy(blk, dx, 16);
        memcpy(blk, z, r);
        decrypt_block(&c, blk, y);                     /* P1 = IV ^ Dec(Z | tail of Dec(X)) */
        xor16(y, chain);
    }
    return ORACLE_SUCCESS;
}

/* micro_aes.c:799-818 (CFB_cipher) */
static void cfb_cipher(int keybits, const uint8_t *key, const uint8_t iv[16], int encrypt,
                       const uint8_t *x, size_t len, uint8_t *y)
{
    aes_ctx c;
    uint8_t fb[16], ks[16];
    size_t i;
    key_setup(&c, keybits, key);
    memcpy(fb, iv, 16);
    for (; len; ) {
        const size_t n = len < 16 ? len : 16;
        encrypt_block(&c, fb, ks);
        for (i = 0; i < n; ++i) {
            const uint8_t ct = encrypt ? (uint8_t)(x[i] ^ ks[i]) : x[i];
            y[i] = x[i] ^ ks[i];
            fb[i] = ct;                                /* IV_next = ciphertext */
        }
        x += n; y += n; len -= n;
    }
}

void oracle_cfb_decrypt(int keybits, const uint8_t *key, const uint8_t iv[16],
                        const void *in, size_t len, void *out)
{
    cfb_cipher(keybits, key, iv, 0, (const uint8_t *)in, len, (uint8_t *)out);
}

void oracle_cfb_encrypt(int keybits, const uint8_t *key, const uint8_t iv[16],
                        const void *in, size_t len, void *out)
{
    cfb_cipher(keybits, key, iv, 1, (const uint8_t *)in, len, (uint8_t *)out);
}

/* ------------------------------------------------------------------------ */
/* OCB                                                                      */
/* ------------------------------------------------------------------------ */

/* micro_aes.c:433-443 (doubleBblock): big-endian shift left, 0x87 into the last byte */
static void ocb_double(uint8_t b[16])
{
    const unsigned msb = b[0] >> 7;
    int i;
    for (i = 0; i < 15; ++i) b[i] = (uint8_t)(b[i] << 1 | b[i + 1] >> 7);
    b[15] = (uint8_t)(b[15] << 1);
    if (msb) b[15] ^= 0x87;
}

/* micro_aes.c:1662-1680 (getDelta): delta = delta0 ^ XOR of the L_k selected by index; the
 * reference's mask arithmetic picks exactly the bits of the Gray code index ^ (index >> 1),
 * i.e. Offset_i = Offset_(i-1) ^ L_ntz(i) of RFC 7253 unrolled */
static void ocb_delta(uint64_t index, const uint8_t Ldollar[16], const uint8_t delta0[16], uint8_t delta[16])
{
    uint8_t L[16];
    uint64_t gray = index ^ (index >> 1);
    memcpy(L, Ldollar, 16);
    memcpy(delta, delta0, 16);
    for (; gray; gray >>= 1) {
        ocb_double(L);                                 /* L_0 = double(L_$), L_k = double(L_(k-1)) */
        if (gray & 1) xor16(delta, L);
    }
}

/* micro_aes.c:1693-1767 (OCB_cipher) */
static void ocb_cipher(int keybits, const uint8_t *key, const uint8_t nonce[12], int encrypt,
                       const uint8_t *aad, size_t aadlen, const uint8_t *x, size_t len,
                       uint8_t *y, size_t taglen, uint8_t tag[16])
{
    aes_ctx c;
    uint8_t Lstar[16] = {0}, Ldollar[16], ktop[16] = {0}, stretch[24], off0[16], delta[16], sum[16] = {0};
    uint8_t blk[16], acc[16];
    const unsigned bottom = nonce[11] % 64;
    size_t n = len / 16, r = len % 16, i, j;

    key_setup(&c, keybits, key);
    encrypt_block(&c, Lstar, Lstar);                   /* L_* = Enc(0), L_$ = double(L_*): getSubkeys */
    memcpy(Ldollar, Lstar, 16);
    ocb_double(Ldollar);
    /* nonce block: tag length (128 mod 128 = 0) in the top 7 bits, then 0..01, then the nonce with
     * its last six bits cleared (micro_aes.c:1709-1712) */
    memcpy(ktop + 4, nonce, 12);
    ktop[0] |= (uint8_t)(taglen << 4);                 /* :1707, OCB_TAG_LEN (16 wraps to 0) */
    ktop[3] |= 1;
    ktop[15] &= 0xC0;
    encrypt_block(&c, ktop, ktop);
    memcpy(stretch, ktop, 16);
    for (i = 0; i < 8; ++i) stretch[16 + i] = ktop[i] ^ ktop[i + 1];
    for (i = 0; i < 16; ++i)                           /* Offset_0 = Stretch[bottom .. bottom+128) */
        off0[i] = (uint8_t)(((unsigned)stretch[i + bottom / 8] << 8 | stretch[i + bottom / 8 + 1]) >> (8 - bottom % 8));

    for (i = 0; i < n; ++i) {
        const uint8_t *p = encrypt ? x + 16 * i : NULL;
        ocb_delta(i + 1, Ldollar, off0, delta);
        memcpy(blk, x + 16 * i, 16);
        xor16(blk, delta);
        if (encrypt) encrypt_block(&c, blk, blk); else decrypt_block(&c, blk, blk);
        xor16(blk, delta);
        if (p) xor16(sum, p);                          /* checksum of the PLAINtext */
        memcpy(y + 16 * i, blk, 16);
        if (!encrypt) xor16(sum, blk);
    }
    if (n) ocb_delta(n, Ldollar, off0, delta); else memcpy(delta, off0, 16);
    if (r) {                                           /* Y_* = Enc(L_* ^ delta_n) ^ X_*, pad the checksum */
        uint8_t pad[16];
        xor16(delta, Lstar);
        encrypt_block(&c, delta, pad);
        for (j = 0; j < r; ++j) {
            const uint8_t in = x[16 * n + j], out = in ^ pad[j];
            y[16 * n + j] = out;
            sum[j] ^= encrypt ? in : out;
        }
        sum[r] ^= 0x80;
    }
    xor16(sum, delta);
    xor16(sum, Ldollar);
    encrypt_block(&c, sum, tag);                       /* tag = Enc(checksum ^ delta ^ L_$) so far */

    /* PMAC of the associated data (micro_aes.c:1750-1765) */
    memset(acc, 0, 16);
    n = aadlen / 16; r = aadlen % 16;
    for (i = 0; i < n; ++i) {
        ocb_delta(i + 1, Ldollar, aad + 16 * i, blk);
        encrypt_block(&c, blk, blk);
        xor16(acc, blk);
    }
    if (r) {
        uint8_t zero[16] = {0};
        ocb_delta(n, Ldollar, zero, blk);
        for (j = 0; j < r; ++j) blk[j] ^= aad[16 * n + j];
        blk[r] ^= 0x80;
        xor16(blk, Lstar);
        encrypt_block(&c, blk, blk);
        xor16(acc, blk);
    }
    xor16(tag, acc);
}

/* micro_aes.c:1779-1789 */
void oracle_ocb_encrypt_ex(int keybits, const uint8_t *key, const uint8_t nonce[12],
                           const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen)
{
    uint8_t tag[16];
    ocb_cipher(keybits, key, nonce, 1, (const uint8_t *)aad, aadlen, (const uint8_t *)in, len, (uint8_t *)out, taglen, tag);
    memcpy((uint8_t *)out + len, tag, taglen);         /* :1783 */
}

int oracle_ocb_decrypt_ex(int keybits, const uint8_t *key, const uint8_t nonce[12],
                          const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen)
{
    uint8_t tag[16], got[16];
    memcpy(got, (const uint8_t *)in + len, taglen);
    ocb_cipher(keybits, key, nonce, 0, (const uint8_t *)aad, aadlen, (const uint8_t *)in, len, (uint8_t *)out, taglen, tag);
    return memcmp(tag, got, taglen) ? ORACLE_AUTHENTICATION_ERROR : ORACLE_SUCCESS;   /* :1807 */
}

void oracle_ocb_encrypt(int keybits, const uint8_t *key, const uint8_t nonce[12],
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out)
{
    oracle_ocb_encrypt_ex(keybits, key, nonce, aad, aadlen, in, len, out, 16);
}

int oracle_ocb_decrypt(int keybits, const uint8_t *key, const uint8_t nonce[12],
                       const void *aad, size_t aadlen, const void *in, size_t len, void *out)
{
    return oracle_ocb_decrypt_ex(keybits, key, nonce, aad, aadlen, in, len, out, 16);
}

/* ------------------------------------------------------------------------ */
/* CCM (micro_aes.c:1219-1315)                              */
/* ------------------------------------------------------------------------ */

/* CBC-MAC absorb of `len` bytes, the last partial block zero-padded: the
 * rijndaelEncrypt instance of xMac (micro_aes.c:551-570) */
static void cbcmac_absorb(const aes_ctx *c, const uint8_t *x, size_t len, uint8_t m[16])
{
    size_t i;
    for (; len >= 16; len -= 16, x += 16) {
        xor16(m, x);
        encrypt_block(c, m, m);
    }
    if (len) {
        for (i = 0; i < len; ++i) m[i] ^= x[i];
        encrypt_block(c, m, m);
    }
}

/* CCMtag (micro_aes.c:1222-1256) with CCM_NONCE_LEN = 11, CCM_TAG_LEN = 16
 * (micro_aes.h:104-105): iv = 03 || nonce || 00000000 */
static void ccm_tag(const aes_ctx *c, const uint8_t iv[16], const uint8_t *aad, size_t aadlen,
                    const uint8_t *pt, size_t len, size_t taglen, uint8_t tag[16])
{
    uint8_t m[16], a[16] = {0}, s0[16];
    size_t head = 0, i;
    memcpy(m, iv, 16);
    m[0] |= (uint8_t)((taglen - 2) << 2);        /* :1229, CCM_TAG_LEN */
    for (i = 0; i < 8 && i < sizeof len; ++i)    /* xorBEint(M, ptextLen, LAST), :1230 */
        m[15 - i] ^= (uint8_t)(len >> (8 * i));
    if (aadlen) {
        size_t p;
        m[0] |= 0x40;
        encrypt_block(c, m, m);                  /* :1235 */
        if (aadlen > 0xFEFF) {                   /* :1236-1240 */
            a[0] = 0xFF; a[1] = 0xFE;
            a[2] = (uint8_t)(aadlen >> 24); a[3] = (uint8_t)(aadlen >> 16);
            a[4] = (uint8_t)(aadlen >> 8);  a[5] = (uint8_t)aadlen;
            p = 6;
        } else {
            a[0] = (uint8_t)(aadlen >> 8); a[1] = (uint8_t)aadlen;
            p = 2;
        }
        head = 16 - p;
        if (head > aadlen) head = aadlen;
        memcpy(a + p, aad, head);                /* :1243 */
    }
    cbcmac_absorb(c, a, 16, m);                  /* :1247 (an all-zero block when there is no AAD) */
    if (aadlen > head) cbcmac_absorb(c, aad + head, aadlen - head, m);
    cbcmac_absorb(c, pt, len, m);                /* :1252 */
    encrypt_block(c, iv, s0);                    /* :1254 */
    for (i = 0; i < 16; ++i) tag[i] = m[i] ^ s0[i];
}

static void ccm_iv(const uint8_t nonce[11], uint8_t iv[16])
{
    memset(iv, 0, 16);
    iv[0] = 14 - 11;                             /* :1273 */
    memcpy(iv + 1, nonce, 11);
}

/* micro_aes.c:1268-1282; out holds len + 16 */
void oracle_ccm_encrypt_ex(int keybits, const uint8_t *key, const uint8_t nonce[11],
                           const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen)
{
    aes_ctx c;
    uint8_t iv[16], ctr[16], tag[16];
    key_setup(&c, keybits, key);
    ccm_iv(nonce, iv);
    ccm_tag(&c, iv, (const uint8_t *)aad, aadlen, (const uint8_t *)in, len, taglen, tag);
    memcpy(ctr, iv, 16);
    ctr_add(ctr, 1);                             /* CCM_GCM pre-increment, :939-941 */
    ctr_stream(&c, ctr, (const uint8_t *)in, len, (uint8_t *)out);
    memcpy((uint8_t *)out + len, tag, taglen);   /* :1281 */
}

int oracle_ccm_decrypt_ex(int keybits, const uint8_t *key, const uint8_t nonce[11],
                          const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen)
{
    aes_ctx c;
    uint8_t iv[16], ctr[16], tag[16];
    key_setup(&c, keybits, key);
    ccm_iv(nonce, iv);
    memcpy(ctr, iv, 16);
    ctr_add(ctr, 1);
    ctr_stream(&c, ctr, (const uint8_t *)in, len, (uint8_t *)out);
    ccm_tag(&c, iv, (const uint8_t *)aad, aadlen, (const uint8_t *)out, len, taglen, tag);
    return memcmp(tag, (const uint8_t *)in + len, taglen) ? ORACLE_AUTHENTICATION_ERROR : ORACLE_SUCCESS;   /* :1308 */
}

void oracle_ccm_encrypt(int keybits, const uint8_t *key, const uint8_t nonce[11],
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out)
{
    oracle_ccm_encrypt_ex(keybits, key, nonce, aad, aadlen, in, len, out, 16);
}

int oracle_ccm_decrypt(int keybits, const uint8_t *key, const uint8_t nonce[11],
                       const void *aad, size_t aadlen, const void *in, size_t len, void *out)
{
    return oracle_ccm_decrypt_ex(keybits, key, nonce, aad, aadlen, in, len, out, 16);
}

/* ------------------------------------------------------------------------ */
/* CMAC helpers, EAX and SIV                              */
/* ------------------------------------------------------------------------ */

/* doubleBblock, micro_aes.c:434-444: big-endian 128-bit value times x, 0x87 fold */
static void dbl_be(uint8_t b[16])
{
    const int carry = b[0] >> 7;
    int i;
    for (i = 0; i < 15; ++i) b[i] = (uint8_t)(b[i] << 1 | b[i + 1] >> 7);
    b[15] = (uint8_t)(b[15] << 1) ^ (uint8_t)(carry * 0x87);
}

/* getSubkeys with quad = 1, micro_aes.c:593-604: K1 = 2 E(0), K2 = 4 E(0) */
static void cmac_subkeys(const aes_ctx *c, uint8_t k1[16], uint8_t k2[16])
{
    memset(k1, 0, 16);
    encrypt_block(c, k1, k1);
    dbl_be(k1);
    memcpy(k2, k1, 16);
    dbl_be(k2);
}

/* cMac, micro_aes.c:576-590: CMAC continued from the running state `mac` */
static void cmac_continue(const aes_ctx *c, const uint8_t k1[16], const uint8_t k2[16],
                          const uint8_t *data, size_t n, uint8_t mac[16])
{
    const size_t s = n ? (n - 1) % 16 + 1 : 0;         /* bytes in the last block */
    uint8_t last[16] = {0};
    cbcmac_absorb(c, data, n - s, mac);
    if (s) memcpy(last, data + n - s, s);
    if (s < 16) { last[s] = 0x80; xor16(last, k2); }
    else xor16(last, k1);
    xor16(mac, last);
    encrypt_block(c, mac, mac);
}

/* oMac without EAXP, micro_aes.c:1531-1550: OMAC^t(data) = CMAC([t]_128 || data) */
static void omac(const aes_ctx *c, int t, const uint8_t k1[16], const uint8_t k2[16],
                 const uint8_t *data, size_t n, uint8_t out[16])
{
    memset(out, 0, 16);
    if (n == 0) memcpy(out, k1, 16);
    out[15] ^= (uint8_t)t;
    encrypt_block(c, out, out);
    if (n) cmac_continue(c, k1, k2, data, n, out);
}

static void eax_tag(const aes_ctx *c, const uint8_t nonce[16], const uint8_t *aad, size_t aadlen,
                    const uint8_t *ct, size_t len, uint8_t N[16], uint8_t tag[16])
{
    uint8_t k1[16], k2[16], h[16], m[16];
    int i;
    cmac_subkeys(c, k1, k2);
    omac(c, 0, k1, k2, nonce, 16, N);                  /* :1578 */
    omac(c, 1, k1, k2, aad, aadlen, h);                /* :1591 */
    omac(c, 2, k1, k2, ct, len, m);                    /* :1593 */
    for (i = 0; i < 16; ++i) tag[i] = N[i] ^ h[i] ^ m[i];
}

/* micro_aes.c:1564-1598 (EAX_NONCE_LEN = 16, EAX_TAG_LEN = 16); out holds len + 16 */
void oracle_eax_encrypt_ex(int keybits, const uint8_t *key, const uint8_t nonce[16],
                           const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen)
{
    aes_ctx c;
    uint8_t k1[16], k2[16], N[16], ctr[16], tag[16];
    key_setup(&c, keybits, key);
    cmac_subkeys(&c, k1, k2);
    omac(&c, 0, k1, k2, nonce, 16, N);
    memcpy(ctr, N, 16);
    ctr_stream(&c, ctr, (const uint8_t *)in, len, (uint8_t *)out);    /* CTR_DEFAULT: starts AT N, :1584 */
    eax_tag(&c, nonce, (const uint8_t *)aad, aadlen, (const uint8_t *)out, len, N, tag);
    memcpy((uint8_t *)out + len, tag, taglen);                        /* :1594 */
}

int oracle_eax_decrypt_ex(int keybits, const uint8_t *key, const uint8_t nonce[16],
                          const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen)
{
    aes_ctx c;
    uint8_t N[16], tag[16];
    key_setup(&c, keybits, key);
    eax_tag(&c, nonce, (const uint8_t *)aad, aadlen, (const uint8_t *)in, len, N, tag);
    if (memcmp(tag, (const uint8_t *)in + len, taglen)) return ORACLE_AUTHENTICATION_ERROR;   /* :1638 */
    ctr_stream(&c, N, (const uint8_t *)in, len, (uint8_t *)out);
    return ORACLE_SUCCESS;
}

void oracle_eax_encrypt(int keybits, const uint8_t *key, const uint8_t nonce[16],
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out)
{
    oracle_eax_encrypt_ex(keybits, key, nonce, aad, aadlen, in, len, out, 16);
}

int oracle_eax_decrypt(int keybits, const uint8_t *key, const uint8_t nonce[16],
                       const void *aad, size_t aadlen, const void *in, size_t len, void *out)
{
    return oracle_eax_decrypt_ex(keybits, key, nonce, aad, aadlen, in, len, out, 16);
}

/* S2V for one AAD unit, micro_aes.c:1325-1359, written the way RFC 5297 states it (xorend for
 * messages of 16 bytes and more, dbl + pad below that); c = first key half */
static void s2v(const aes_ctx *c, const uint8_t *aad, size_t aadlen, const uint8_t *pt, size_t len, uint8_t v[16])
{
    uint8_t k1[16], k2[16], y[16], t[16];
    size_t i;
    cmac_subkeys(c, k1, k2);
    encrypt_block(c, k1, y);                           /* Y_0 = CMAC(0^128) = E(K1), :1332 */
    if (aadlen) {                                      /* :1338-1344 */
        memset(t, 0, 16);
        cmac_continue(c, k1, k2, aad, aadlen, t);
        dbl_be(y);
        xor16(y, t);
    }
    memset(v, 0, 16);
    if (len >= 16) {                                   /* CMAC(pt xorend Y), :1350-1358 */
        uint8_t *tmp = (uint8_t *)malloc(len);
        memcpy(tmp, pt, len);
        for (i = 0; i < 16; ++i) tmp[len - 16 + i] ^= y[i];
        cmac_continue(c, k1, k2, tmp, len, v);
        free(tmp);
    } else {                                           /* CMAC(dbl(Y) ^ pad(pt)), :1345-1349 */
        dbl_be(y);
        for (i = 0; i < len; ++i) y[i] ^= pt[i];
        y[len] ^= 0x80;
        cmac_continue(c, k1, k2, y, 16, v);
    }
}

static void siv_ctr(const aes_ctx *c2, const uint8_t v[16], const uint8_t *x, size_t len, uint8_t *y)
{
    uint8_t ctr[16];
    memcpy(ctr, v, 16);
    ctr[8] &= 0x7F; ctr[12] &= 0x7F;                   /* SIV_CTR, micro_aes.c:931-934 */
    ctr_stream(c2, ctr, x, len, y);
}

/* micro_aes.c:1372-1382: keys = K1 || K2, iv = synthetic IV (16 bytes out), out holds len */
void oracle_siv_encrypt(int keybits, const uint8_t *keys, const void *aad, size_t aadlen,
                        const void *in, size_t len, uint8_t iv[16], void *out)
{
    aes_ctx c1, c2;
    key_setup(&c1, keybits, keys);
    key_setup(&c2, keybits, keys + keybits / 8);
    s2v(&c1, (const uint8_t *)aad, aadlen, (const uint8_t *)in, len, iv);
    siv_ctr(&c2, iv, (const uint8_t *)in, len, (uint8_t *)out);
}

/* micro_aes.c:1394-1410: decrypts first, then compares the recomputed IV (plaintext stays) */
int oracle_siv_decrypt(int keybits, const uint8_t *keys, const uint8_t iv[16], const void *aad, size_t aadlen,
                       const void *in, size_t len, void *out)
{
    aes_ctx c1, c2;
    uint8_t v[16];
    key_setup(&c1, keybits, keys);
    key_setup(&c2, keybits, keys + keybits / 8);
    siv_ctr(&c2, iv, (const uint8_t *)in, len, (uint8_t *)out);
    s2v(&c1, (const uint8_t *)aad, aadlen, (const uint8_t *)out, len, v);
    return memcmp(v, iv, 16) ? ORACLE_AUTHENTICATION_ERROR : ORACLE_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* synthetic data                                                           */
/* ------------------------------------------------------------------------ */

static uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void oracle_fill_splitmix64(uint64_t seed, uint64_t first_word, void *dst, size_t nwords)
{
    uint8_t *p = (uint8_t *)dst;
    size_t w;
    int i;
    for (w = 0; w < nwords; ++w) {
        uint64_t v = splitmix64(seed + first_word + w);
        for (i = 0; i < 8; ++i, v >>= 8) *p++ = (uint8_t)v;
    }
}
