/*
 * micro_aes.h -- the hot-path subset of the micro-AES C API, served by the
 * B200 (sm_100a) engine in libuaes_b200.so.
 *
 * This header declares, with the reference's names, argument order and return
 * codes, the entry points the reference exports when only ECB, CTR (CTR_NA),
 * XEX/XTS, GCM, GCM_SIV, OCB, CCM, EAX and SIV are enabled:
 *
 *     function            replaces (polfosol/micro-AES)
 *     ------------------  ---------------------------------------------
 *     AES_ECB_encrypt     micro_aes.h:173-176, micro_aes.c:636-653
 *     AES_ECB_decrypt     micro_aes.h:178-181, micro_aes.c:663-680
 *     AES_CTR_encrypt     micro_aes.h:256-260, micro_aes.c:962-976
 *     AES_CTR_decrypt     micro_aes.h:262-266, micro_aes.c:986-990
 *     AES_XTS_encrypt     micro_aes.h:239-243, micro_aes.c:1066-1074
 *     AES_XTS_decrypt     micro_aes.h:245-249, micro_aes.c:1085-1093
 *     AES_GCM_encrypt     micro_aes.h:294-300, micro_aes.c:1164-1179
 *     AES_GCM_decrypt     micro_aes.h:302-308, micro_aes.c:1192-1212
 *     GCM_SIV_encrypt     micro_aes.h:386-392, micro_aes.c:1474-1487   (SURVEY 8f "next" row 1)
 *     GCM_SIV_decrypt     micro_aes.h:394-400, micro_aes.c:1499-1516
 *     AES_OCB_encrypt     micro_aes.h:336-342, micro_aes.c:1779-1789   (SURVEY 8f "next" row 3)
 *     AES_OCB_decrypt     micro_aes.h:344-350, micro_aes.c:1802-1813
 *     AES_CCM_encrypt     micro_aes.h:315-321, micro_aes.c:1268-1282   (SURVEY 8f "next" row 4; one
 *     AES_CCM_decrypt     micro_aes.h:323-329, micro_aes.c:1295-1314    GPU lane per message: use the
 *                                                                        batch call of uaes_b200.h for speed)
 *     AES_EAX_encrypt     micro_aes.h:357-367, micro_aes.c:1564-1598   (likewise)
 *     AES_EAX_decrypt     micro_aes.h:369-379, micro_aes.c:1613-1648
 *     AES_SIV_encrypt     micro_aes.h:273-279, micro_aes.c:1372-1382   (likewise)
 *     AES_SIV_decrypt     micro_aes.h:281-287, micro_aes.c:1394-1410
 *
 * A program written against the reference keeps its `#include "micro_aes.h"`,
 * drops micro_aes.c from its build and links one of
 *     -lmicro_aes_128   -lmicro_aes_192   -lmicro_aes_256
 * (the reference fixes the key size at compile time with AES___, micro_aes.h:17;
 * each shim library is that choice).  All other modes of the reference are out
 * of scope and their feature macros are 0 here, so code
 * guarded by `#if CBC`, `#if CCM`, ... compiles away exactly as it does with
 * the reference's own switches.
 *
 * Buffers may be ordinary host memory, pinned host memory, or CUDA device /
 * managed memory (detected per call); device buffers are processed in place in
 * HBM with no copies.  The `void` functions cannot report CUDA failures: query
 * uaes_last_error() from uaes_b200.h.
 *
 * Plain ANSI C; no CUDA types appear in any signature.
 */
#ifndef MICRO_AES_H_
#define MICRO_AES_H_

#ifndef AES___
#define AES___          128     /* 128, 192 or 256: must match the linked shim */
#endif

/* feature switches, named as in the reference (micro_aes.h:23-69) */
#define BLOCKCIPHERS    1
#define AEAD_MODES      1
#define ECB             1
#define CTR             1
#define CTR_NA          1
#define XEX             1
#define XTS             1
#define GCM             1
#define CBC             0       /* the serial ENcrypt directions of CBC / CFB are not provided ...  */
#define CFB             0
#define CBC_DECRYPT     1       /* ... their block-parallel DEcrypt directions are */
#define CFB_DECRYPT     1
#define OFB             0
#define KWA             0
#define FPE             0
#define CMAC            0
#define CCM             1
#define EAX             1
#define EAXP            0
#define SIV             1
#define GCM_SIV         1
#define OCB             1
#define POLY1305        0
#define MICRO_RJNDL     0

/* Compile-time variants of the reference (micro_aes.h:56, 78-80, 97-110).  The reference is
 * configured by editing these lines of its header; here the same names may also be set with -D
 * (for the enum constants: -DUAES_GCM_NONCE_LEN=..., -DUAES_GCM_TAG_LEN=...).  The application and
 * the shim library must be built with the same values, exactly as the application and micro_aes.c
 * share one header in the reference; `make shim VARIANT=... DEFS=...` (csrc/Makefile) builds a shim
 * for any combination. */
#ifndef CTS
#define CTS             1       /* CBC: CS3 ciphertext stealing (micro_aes.h:56); 0 = whole blocks only */
#endif
#ifndef AES_PADDING
#define AES_PADDING     0       /* ECB tail: 0 zeros, 1 PKCS#7, 2 ISO/IEC 7816-4 (micro_aes.h:78-80) */
#endif
#define DECRYPTION      1
#ifndef PRESET_COUNTER
#define PRESET_COUNTER  0       /* 0: CTR takes a 12-byte IV, counter field starts at 1;
                                   1: iv[16] is the first counter block verbatim (micro_aes.h:100) */
#endif
#ifndef UAES_CTR_IV_LENGTH
#define UAES_CTR_IV_LENGTH 12   /* bytes of iv copied into the counter block (micro_aes.h:99); <= 16 */
#endif
#ifndef UAES_CTR_START_VALUE
#define UAES_CTR_START_VALUE 1  /* XORed big-endian into the end of the block (micro_aes.h:98, micro_aes.c:971) */
#endif
#ifndef UAES_GCM_NONCE_LEN
#define UAES_GCM_NONCE_LEN 12
#endif
#ifndef UAES_GCM_TAG_LEN
#define UAES_GCM_TAG_LEN   16
#endif
#ifndef UAES_CCM_TAG_LEN
#define UAES_CCM_TAG_LEN   16      /* even, 4..16 (micro_aes.h:105) */
#endif
#ifndef UAES_EAX_TAG_LEN
#define UAES_EAX_TAG_LEN   16
#endif
#ifndef UAES_OCB_TAG_LEN
#define UAES_OCB_TAG_LEN   16
#endif

enum constant_parameters_of_modes
{
    CTR_START_VALUE = UAES_CTR_START_VALUE,     /* micro_aes.h:98  */
#if PRESET_COUNTER
    CTR_IV_LENGTH   = 16,       /* micro_aes.h:99-100: the whole counter block */
#else
    CTR_IV_LENGTH   = UAES_CTR_IV_LENGTH,       /* micro_aes.h:99  */
#endif
    GCM_NONCE_LEN   = UAES_GCM_NONCE_LEN,   /* micro_aes.h:108: 12 is the recommended value, others are supported */
    GCM_TAG_LEN     = UAES_GCM_TAG_LEN,     /* micro_aes.h:109 */
    SIVGCM_NONCE_LEN = 12,      /* micro_aes.h:112 */
    SIVGCM_TAG_LEN  = 16,       /* micro_aes.h:113 */
    CCM_NONCE_LEN   = 11,       /* micro_aes.h:104 */
    CCM_TAG_LEN     = UAES_CCM_TAG_LEN,         /* micro_aes.h:105 */
    EAX_NONCE_LEN   = 16,       /* micro_aes.h:120 */
    EAX_TAG_LEN     = UAES_EAX_TAG_LEN,         /* micro_aes.h:121 */
    OCB_NONCE_LEN   = 12,       /* micro_aes.h:116 */
    OCB_TAG_LEN     = UAES_OCB_TAG_LEN,         /* micro_aes.h:117 */
#if AES___ != 256 && AES___ != 192
    AES_KEYLENGTH   = 16
#else
    AES_KEYLENGTH   = AES___ / 8
#endif
};

#include <stddef.h>
#include <limits.h>
#if defined(__STDC_VERSION__) && __STDC_VERSION__ >= 199901L || defined(__cplusplus)
#include <stdint.h>
#elif CHAR_BIT == 8 && !defined(UINT8_MAX)
typedef unsigned char uint8_t;
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ECB: output holds ceil16(ptextLen) bytes, the tail block is zero padded; with AES_PADDING 1 / 2
 * a padding block is always added: (ptextLen / 16 + 1) * 16 bytes (micro_aes.c:610-621) */
void AES_ECB_encrypt(const uint8_t *key, const void *pntxt, const size_t ptextLen, void *crtxt);
/* returns M_DECRYPTION_ERROR when crtxtLen is not a multiple of 16 (full blocks are
 * still decrypted, tail bytes copied through) */
char AES_ECB_decrypt(const uint8_t *key, const void *crtxt, const size_t crtxtLen, void *pntxt);

/* CTR: iv = CTR_IV_LENGTH bytes; counter block = iv || BE32(1) (or iv[16] itself when
 * PRESET_COUNTER), incremented as a 56-bit big-endian integer in bytes 9..15 (micro_aes.c:421-427) */
void AES_CTR_encrypt(const uint8_t *key, const uint8_t *iv,
                     const void *pntxt, const size_t ptextLen, void *crtxt);
void AES_CTR_decrypt(const uint8_t *key, const uint8_t *iv,
                     const void *crtxt, const size_t crtxtLen, void *pntxt);

/* XTS: keys = K1 || K2 (2 * AES_KEYLENGTH bytes), one data unit per call, tweak = 16
 * bytes or NULL for sector 0; ciphertext stealing when the length is ragged;
 * returns M_DATALENGTH_ERROR below 16 bytes */
char AES_XTS_encrypt(const uint8_t *keys, const uint8_t *tweak,
                     const void *pntxt, const size_t ptextLen, void *crtxt);
char AES_XTS_decrypt(const uint8_t *keys, const uint8_t *tweak,
                     const void *crtxt, const size_t crtxtLen, void *pntxt);

/* GCM: GCM_NONCE_LEN-byte nonce; crtxt holds ptextLen + GCM_TAG_LEN bytes (tag appended) */
void AES_GCM_encrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *pntxt, const size_t ptextLen, void *crtxt);
/* verifies the tag at crtxt + crtxtLen first; on mismatch returns
 * M_AUTHENTICATION_ERROR and leaves pntxt untouched */
char AES_GCM_decrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *crtxt, const size_t crtxtLen, void *pntxt);

/* only the block-parallel DEcrypt directions of CBC and CFB exist on the GPU
 * (CBC_DECRYPT / CFB_DECRYPT above); CBC / CFB stay 0 because the serial encrypt directions
 * (micro_aes.c:697-733, 826-830) are not provided.  CBC follows CTS like the reference: CS3
 * stealing with CTS = 1 (any length >= 16), whole blocks only with CTS = 0 (M_DATALENGTH_ERROR
 * otherwise).
 * Replaces micro_aes.h:194-198 / micro_aes.c:746-782 and micro_aes.h:211-215 / micro_aes.c:840-845. */
char AES_CBC_decrypt(const uint8_t *key, const uint8_t iVec[16],
                     const void *crtxt, const size_t crtxtLen, void *pntxt);
void AES_CFB_decrypt(const uint8_t *key, const uint8_t iVec[16],
                     const void *crtxt, const size_t crtxtLen, void *pntxt);

/* OCB (RFC 7253): 12-byte nonce; crtxt holds ptextLen + OCB_TAG_LEN bytes */
void AES_OCB_encrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *pntxt, const size_t ptextLen, void *crtxt);
char AES_OCB_decrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *crtxt, const size_t crtxtLen, void *pntxt);

/* CCM: 11-byte nonce; crtxt holds ptextLen + CCM_TAG_LEN bytes */
void AES_CCM_encrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *pntxt, const size_t ptextLen, void *crtxt);
/* decrypts, then authenticates (micro_aes.c:1304-1312) */
char AES_CCM_decrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *crtxt, const size_t crtxtLen, void *pntxt);

/* EAX (not EAX'): 16-byte nonce; crtxt holds ptextLen + EAX_TAG_LEN bytes */
void AES_EAX_encrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *pntxt, const size_t ptextLen, void *crtxt);
/* authenticates, then decrypts: pntxt is untouched on M_AUTHENTICATION_ERROR */
char AES_EAX_decrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *crtxt, const size_t crtxtLen, void *pntxt);

/* SIV (RFC 5297): keys = two keys of AES_KEYLENGTH bytes; iv = the 16-byte synthetic IV */
void AES_SIV_encrypt(const uint8_t *keys,
                     const void *aData, const size_t aDataLen,
                     const void *pntxt, const size_t ptextLen,
                     uint8_t iv[16], void *crtxt);
char AES_SIV_decrypt(const uint8_t *keys, const uint8_t iv[16],
                     const void *aData, const size_t aDataLen,
                     const void *crtxt, const size_t crtxtLen, void *pntxt);

/* GCM-SIV (RFC 8452): 12-byte nonce; crtxt holds ptextLen + SIVGCM_TAG_LEN bytes */
void GCM_SIV_encrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *pntxt, const size_t ptextLen, void *crtxt);
/* decrypts, then authenticates: M_AUTHENTICATION_ERROR means pntxt must be discarded */
char GCM_SIV_decrypt(const uint8_t *key, const uint8_t *nonce,
                     const void *aData, const size_t aDataLen,
                     const void *crtxt, const size_t crtxtLen, void *pntxt);

#ifdef __cplusplus
}
#endif

/* result codes, values as in the reference (micro_aes.h:469-476) */
enum function_result_codes
{
    M_ENCRYPTION_ERROR     = 0x1E,
    M_DECRYPTION_ERROR     = 0x1D,
    M_AUTHENTICATION_ERROR = 0x1A,
    M_DATALENGTH_ERROR     = 0x01,
    M_RESULT_SUCCESS       = 0
};

#endif /* MICRO_AES_H_ */
