/*
 * micro_aes_shim.c -- the reference's symbol names on top of libuaes_b200.so.  ANSI C.
 *
 * Compiled once per key size (-DAES___=128 / 192 / 256) into libmicro_aes_<bits>.so, mirroring
 * the reference's compile-time choice (micro_aes.h:17).  A program built against the reference
 * keeps its source and its `#include "micro_aes.h"`, and links one of these libraries in place
 * of micro_aes.c.  Each function forwards to the run-time-key-length entry point of
 * uaes_b200.h; return codes are the reference's (micro_aes.h:469-476).  The void functions
 * latch failures for uaes_last_error().
 *
 * The reference's other compile-time variants select the matching run-time entry point here:
 * PRESET_COUNTER (micro_aes.c:964-966), GCM_NONCE_LEN / GCM_TAG_LEN (micro_aes.c:1145-1149, 1178,
 * 1204), AES_PADDING (micro_aes.c:610-621), CTS (micro_aes.c:703-707, 757-759); build a variant with
 * `make shim VARIANT=<name> DEFS="-D..."`.
 */
#include "../../include/micro_aes.h"
#include "../../include/uaes_b200.h"

#define BITS (AES_KEYLENGTH * 8)

/* a failure below zero is a device/runtime problem, not a reference result code: report it as the
 * closest reference code so that callers which test `!= 0` still see an error */
static char code(int rc, char fallback)
{
    if (rc >= 0) return (char)rc;
    return fallback;
}

void AES_ECB_encrypt(const uint8_t *key, const void *pntxt, const size_t ptextLen, void *crtxt)
{
    uaes_ecb_encrypt_padded(BITS, key, pntxt, ptextLen, crtxt, AES_PADDING);     /* micro_aes.c:610-621 */
}

char AES_ECB_decrypt(const uint8_t *key, const void *crtxt, const size_t crtxtLen, void *pntxt)
{
    return code(uaes_ecb_decrypt(BITS, key, crtxt, crtxtLen, pntxt), M_DECRYPTION_ERROR);
}

void AES_CTR_encrypt(const uint8_t *key, const uint8_t *iv,
                     const void *pntxt, const size_t ptextLen, void *crtxt)
{
#if PRESET_COUNTER
    uaes_ctr_crypt_block(BITS, key, iv, 0, pntxt, ptextLen, crtxt);             /* micro_aes.c:964-966 */
#elif UAES_CTR_IV_LENGTH == 12 && UAES_CTR_START_VALUE == 1
    uaes_ctr_crypt(BITS, key, iv, pntxt, ptextLen, crtxt);
#else
    /* any other CTR_IV_LENGTH / CTR_START_VALUE (micro_aes.c:967-971): ctr = iv || 0.., start value XORed
     * big-endian into its last bytes; the block then counts like the preset one */
    uint8_t ctr[16];
    unsigned long start = (unsigned long)UAES_CTR_START_VALUE;
    int i;
    for (i = 0; i < 16; ++i) ctr[i] = i < UAES_CTR_IV_LENGTH ? iv[i] : 0;
    for (i = 15; ; --i) {                                                       /* xorBEint, micro_aes.c:410-415 */
        ctr[i] ^= (uint8_t)start;
        start >>= 8;
        if (!start || i == 0) break;
    }
    uaes_ctr_crypt_block(BITS, key, ctr, 0, pntxt, ptextLen, crtxt);
#endif
}

void AES_CTR_decrypt(const uint8_t *key, const uint8_t *iv,
                     const void *crtxt, const size_t crtxtLen, void *pntxt)
{
    AES_CTR_encrypt(key, iv, crtxt, crtxtLen, pntxt);                           /* micro_aes.c:986-990 */
}

char AES_XTS_encrypt(const uint8_t *keys, const uint8_t *tweak,
                     const void *pntxt, const size_t ptextLen, void *crtxt)
{
    return code(uaes_xts_encrypt(BITS, keys, tweak, pntxt, ptextLen, crtxt), M_ENCRYPTION_ERROR);
}

char AES_XTS_decrypt(const uint8_t *keys, const uint8_t *tweak,
                     const void *crtxt, const size_t crtxtLen, void *pntxt)
{
    return code(uaes_xts_decrypt(BITS, keys, tweak, crtxt, crtxtLen, pntxt), M_DECRYPTION_ERROR);
}

void AES_GCM_encrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *pntxt, const size_t ptextLen, void *crtxt)
{
    uaes_gcm_encrypt_ex(BITS, key, nonce, GCM_NONCE_LEN, aData, aDataLen, pntxt, ptextLen, crtxt, GCM_TAG_LEN);
}

char AES_GCM_decrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *crtxt, const size_t crtxtLen, void *pntxt)
{
    return code(uaes_gcm_decrypt_ex(BITS, key, nonce, GCM_NONCE_LEN, aData, aDataLen, crtxt, crtxtLen, pntxt,
                                    GCM_TAG_LEN), M_DECRYPTION_ERROR);
}

void GCM_SIV_encrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *pntxt, const size_t ptextLen, void *crtxt)
{
    uaes_gcmsiv_encrypt(BITS, key, nonce, aData, aDataLen, pntxt, ptextLen, crtxt);
}

char GCM_SIV_decrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *crtxt, const size_t crtxtLen, void *pntxt)
{
    return code(uaes_gcmsiv_decrypt(BITS, key, nonce, aData, aDataLen, crtxt, crtxtLen, pntxt),
                M_DECRYPTION_ERROR);
}

char AES_CBC_decrypt(const uint8_t *key, const uint8_t iVec[16],
                     const void *crtxt, const size_t crtxtLen, void *pntxt)
{
    return code(uaes_cbc_decrypt_ex(BITS, key, iVec, crtxt, crtxtLen, pntxt, CTS), M_DECRYPTION_ERROR);
}

void AES_CFB_decrypt(const uint8_t *key, const uint8_t iVec[16],
                     const void *crtxt, const size_t crtxtLen, void *pntxt)
{
    uaes_cfb_decrypt(BITS, key, iVec, crtxt, crtxtLen, pntxt);
}

void AES_OCB_encrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *pntxt, const size_t ptextLen, void *crtxt)
{
    uaes_ocb_encrypt_ex(BITS, key, nonce, aData, aDataLen, pntxt, ptextLen, crtxt, OCB_TAG_LEN);
}

char AES_OCB_decrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *crtxt, const size_t crtxtLen, void *pntxt)
{
    return code(uaes_ocb_decrypt_ex(BITS, key, nonce, aData, aDataLen, crtxt, crtxtLen, pntxt, OCB_TAG_LEN),
                M_DECRYPTION_ERROR);
}

void AES_CCM_encrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *pntxt, const size_t ptextLen, void *crtxt)
{
    uaes_ccm_encrypt_ex(BITS, key, nonce, aData, aDataLen, pntxt, ptextLen, crtxt, CCM_TAG_LEN);
}

char AES_CCM_decrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *crtxt, const size_t crtxtLen, void *pntxt)
{
    return code(uaes_ccm_decrypt_ex(BITS, key, nonce, aData, aDataLen, crtxt, crtxtLen, pntxt, CCM_TAG_LEN),
                M_DECRYPTION_ERROR);
}

void AES_EAX_encrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *pntxt, const size_t ptextLen, void *crtxt)
{
    uaes_eax_encrypt_ex(BITS, key, nonce, aData, aDataLen, pntxt, ptextLen, crtxt, EAX_TAG_LEN);
}

char AES_EAX_decrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *crtxt, const size_t crtxtLen, void *pntxt)
{
    return code(uaes_eax_decrypt_ex(BITS, key, nonce, aData, aDataLen, crtxt, crtxtLen, pntxt, EAX_TAG_LEN),
                M_DECRYPTION_ERROR);
}

void AES_SIV_encrypt(const uint8_t *keys,
                     const void *aData, const size_t aDataLen,
                     const void *pntxt, const size_t ptextLen,
                     uint8_t iv[16], void *crtxt)
{
    uaes_siv_encrypt(BITS, keys, aData, aDataLen, pntxt, ptextLen, iv, crtxt);
}

char AES_SIV_decrypt(const uint8_t *keys, const uint8_t iv[16],
                     const void *aData, const size_t aDataLen,
                     const void *crtxt, const size_t crtxtLen, void *pntxt)
{
    return code(uaes_siv_decrypt(BITS, keys, iv, aData, aDataLen, crtxt, crtxtLen, pntxt),
                M_DECRYPTION_ERROR);
}
