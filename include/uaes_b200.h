/*
 * uaes_b200.h -- C ABI of libuaes_b200.so, the B200 (sm_100a) AES bulk engine.
 *
 * micro_aes.h is the drop-in contract (the reference's eight hot-path functions,
 * key size fixed per shim library).  This header is what those shims call, plus
 * the extensions a GPU deployment needs and the reference API cannot express
 * (SURVEY.md section 8b):
 *
 *   - run-time key length (the reference bakes it in with AES___, micro_aes.h:17);
 *   - an error latch, because AES_CTR_encrypt / AES_GCM_encrypt / AES_ECB_encrypt
 *     return void (micro_aes.h:173,256,294) and cannot report a CUDA failure;
 *   - counter-range CTR (resume / shard at keystream block k), the multi-GPU
 *     partitioning of micro_aes.c:943-948;
 *   - batched-sector XTS, i.e. the caller loop over AES_XTS_encrypt with
 *     tweak = LE128(sector) (micro_aes.c:1017-1021) as ONE launch;
 *   - stream selection and asynchronous completion for device-resident buffers.
 *
 * Every data pointer may be host memory (pageable or pinned) or CUDA device /
 * managed memory; the library classifies it per call.  Device buffers are processed
 * in HBM with no copies; host buffers are staged in chunks with copies overlapped
 * against the kernels (pageable memory through the library's own pinned bounce
 * chunks), on one GPU or -- uaes_set_devices -- spread over all of them.  key / iv / nonce / tweak / aad are always small
 * HOST arrays, as in the reference.
 *
 * All functions return 0 (UAES_OK) or a micro_aes.h result code (1, 0x1A, 0x1D) or
 * a negative UAES_E_* value; the failure is also latched for uaes_last_error().
 * There is no CPU fallback: without a usable CUDA device every call fails with
 * UAES_E_NO_DEVICE.
 *
 * Plain C89; no CUDA or torch types in any signature (streams travel as void*).
 */
#ifndef UAES_B200_H_
#define UAES_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned char      uaes_u8;
typedef unsigned long long uaes_u64;

enum uaes_status
{
    UAES_OK                 = 0,
    UAES_DATALENGTH_ERROR   = 0x01,   /* = M_DATALENGTH_ERROR     (micro_aes.h:474) */
    UAES_AUTH_ERROR         = 0x1A,   /* = M_AUTHENTICATION_ERROR (micro_aes.h:473) */
    UAES_DECRYPTION_ERROR   = 0x1D,   /* = M_DECRYPTION_ERROR     (micro_aes.h:472) */
    UAES_E_NO_DEVICE        = -1,     /* no CUDA device / driver: nothing was computed */
    UAES_E_CUDA             = -2,     /* a CUDA runtime call or kernel failed */
    UAES_E_BAD_ARGUMENT     = -3,     /* key size not 128/192/256, ragged sectors, tag length */
    UAES_E_NO_MEMORY        = -4
};

/* ---- library state -------------------------------------------------------- */

/* most recent failure on the calling thread's device context (0 = none) */
int         uaes_last_error(void);
const char *uaes_last_error_string(void);
void        uaes_clear_error(void);

/* number of usable CUDA devices (0 when there is no driver) */
int  uaes_device_count(void);
/* stream used by subsequent calls from this thread (a cudaStream_t passed as void*;
 * NULL = the legacy default stream).  The device is the caller's current device. */
void uaes_set_stream(void *stream);
/* 0 (default): every call returns after the result is complete (reference semantics).
 * 1: calls whose data buffers are all device memory only ENQUEUE work on the stream
 *    set above and return; GCM decrypt still waits for the tag check. */
void uaes_set_async(int enable);
/* pinned host memory, for callers that want full-speed host<->device staging */
void *uaes_host_alloc(size_t bytes);
void  uaes_host_free(void *p);
/* page-lock / release a buffer the caller already owns (cudaHostRegister): worth it for buffers
 * that are used more than once; pageable buffers work too, through the library's own bounce chunks */
int   uaes_host_register(void *p, size_t bytes);
int   uaes_host_unregister(void *p);

/* ---- multi-GPU context (extension 5) --------------------------- */
/* Number of GPUs a call on HOST buffers may be spread over (n <= 0: all of them).  The byte range
 * is cut into one contiguous part per device -- blocks, sectors, tweak positions and GCM shards are
 * independent (micro_aes.c:943-948, 1030-1036) -- and each part runs the staging
 * pipeline of its device on its own host thread, so every PCIe link of the box carries data at
 * once.  Part 0 stays on the calling thread's current device.  Returns the number in effect
 * (default 1, or UAES_DEVICES from the environment).  Device-resident buffers are never spread:
 * shard those yourself with the *_range entry points below. */
int  uaes_set_devices(int n);
int  uaes_get_devices(void);
/* a call is spread only as far as every device still gets this many bytes (default 256 MiB) */
void uaes_set_fanout_min(size_t bytes_per_device);
/* staging pipeline geometry for host buffers: bytes per chunk (default 64 MiB) and chunks in flight
 * per device (default 3); 0 keeps a value.  Implies uaes_shutdown(): no call may be in flight. */
void uaes_set_staging(size_t chunk_bytes, int slots);
/* helper threads that move PAGEABLE caller memory into / out of the pinned bounce chunks
 * (default 4; memcpy only -- no cipher work ever runs on the host) */
void uaes_set_copy_threads(int n);

/* ---- lifecycle --------------------------------------------------------------- */
/* 1: wipe every staging chunk, work area and host-side key schedule after each call, the
 * reference's INCREASE_SECURITY / BURN (micro_aes.c:362-364); costs one extra memset per chunk */
void uaes_set_burn(int enable);
/* give back the grow-on-demand device memory of all devices (work areas, full-size staging);
 * streams and staging chunks stay.  Full-size staging above 256 MiB is released after every call
 * anyway. */
void uaes_trim(void);
/* wipe and release everything on all devices; the next call initialises again.  No call may be
 * in flight. */
void uaes_shutdown(void);
/* number of CUDA kernels this library has launched in this process (bench bookkeeping) */
uaes_u64 uaes_kernel_launches(void);
/* CTR kernel geometry (tuning and tests; the defaults are the measured optimum on B200):
 *   tt_threads       geometry code: 388 (default) = ctr_queue8_kernel: table-driven warps + NARROW bitsliced
 *                    co-runner warps (8 blocks per thread) sharing the counter range through a two-ended
 *                    work queue (AES-128: 16 + 8 warps); 386 = ctr_queue_kernel: 384 table-driven threads
 *                    with two blocks in flight each + 128 wide bitsliced threads on the same queue;
 *                    385 = the same warps with a static split (bs_permille);
 *                    384 = static split, one block in flight; 512 / 768 / 1024 = table-driven warps only
 *   bs_permille      share of the blocks, in 1/1024, given to the bitsliced ALU co-runner warps
 *                    (0 = co-runner off)
 *   bs_min_blocks    calls shorter than this many 16-byte blocks never use the co-runner
 *                    (default 2^23 = 128 MiB; the same threshold serves ECB, XTS, OCB and CFB)
 * A negative value leaves that setting unchanged.  Results do not depend on any of them. */
void uaes_ctr_tuning(int tt_threads, int bs_permille, long long bs_min_blocks);
/* how the calling thread's most recent work-queue CTR launch (geometry 388 / 386) was shared out: units
 * served by the table-driven warps, by the bitsliced warps, and 16-byte blocks per unit.  Waits for
 * the device.  (bench.py derives the shared-memory-lookup roofline from it.) */
int uaes_ctr_queue_stats(uaes_u64 *tt_units, uaes_u64 *bs_units, uaes_u64 *unit_blocks);

/* ---- the hot path, run-time key length ------------------------------------- */
/* keybits = 128, 192 or 256 everywhere (XTS: keys = K1 || K2, 2 * keybits / 8 bytes). */

/* micro_aes.c:636-653; out holds ceil16(len) */
int uaes_ecb_encrypt(int keybits, const uaes_u8 *key, const void *in, size_t len, void *out);
/* micro_aes.c:663-680; UAES_DECRYPTION_ERROR when len % 16 (blocks still decrypted) */
int uaes_ecb_decrypt(int keybits, const uaes_u8 *key, const void *in, size_t len, void *out);

/* micro_aes.c:962-990; iv = 12 bytes, counter block k = iv || BE32(1) plus k as a 56-bit
 * big-endian add over bytes 9..15 (micro_aes.c:421-427) */
int uaes_ctr_crypt(int keybits, const uaes_u8 *key, const uaes_u8 *iv,
                   const void *in, size_t len, void *out);
/* the same keystream starting at block `first_block`: bytes [16*first_block, +len) of the
 * stream AES_CTR_encrypt would produce.  This is how a buffer shards over GPUs. */
int uaes_ctr_crypt_range(int keybits, const uaes_u8 *key, const uaes_u8 *iv,
                         uaes_u64 first_block, const void *in, size_t len, void *out);

/* the reference built with PRESET_COUNTER = 1 (micro_aes.c:964-966, micro_aes.h:100): ctr[16] is
 * counter block 0 verbatim; keystream block k uses ctr + first_block + k (56-bit add, bytes 9..15) */
int uaes_ctr_crypt_block(int keybits, const uaes_u8 *key, const uaes_u8 *ctr, uaes_u64 first_block,
                         const void *in, size_t len, void *out);
/* the reference built with AES_PADDING = 1 (PKCS#7) or 2 (ISO/IEC 7816-4), micro_aes.c:610-621:
 * the last block is always padded, out holds (len / 16 + 1) * 16 bytes; padding = 0 is
 * uaes_ecb_encrypt.  (Like the reference, decryption does not strip or check the padding.) */
int uaes_ecb_encrypt_padded(int keybits, const uaes_u8 *key, const void *in, size_t len, void *out, int padding);

/* micro_aes.c:1066-1093: one data unit, tweak = 16 bytes or NULL (sector 0), stealing */
int uaes_xts_encrypt(int keybits, const uaes_u8 *keys, const uaes_u8 *tweak,
                     const void *in, size_t len, void *out);
int uaes_xts_decrypt(int keybits, const uaes_u8 *keys, const uaes_u8 *tweak,
                     const void *in, size_t len, void *out);
/* A block range of ONE data unit ("a single huge data unit also shards"): in[0] is
 * block first_block of the unit, whose tweak chain T_0 * alpha^k (micro_aes.c:1030-1036) is entered
 * by jump-ahead.  len is a multiple of 16 except for the range that ends the unit, which applies the
 * ciphertext stealing of micro_aes.c:1037-1053 to its last two blocks (len >= 16 always).  Ranges
 * may run on different GPUs or at different times; together they equal one AES_XTS_encrypt call. */
int uaes_xts_crypt_range(int keybits, const uaes_u8 *keys, const uaes_u8 *tweak, uaes_u64 first_block,
                         const void *in, size_t len, void *out, int encrypt);
/* len / sector_bytes consecutive data units, unit j tweaked with LE128(first_sector + j);
 * sector_bytes must be a multiple of 16, len a multiple of sector_bytes.  Equals a loop of
 * AES_XTS_encrypt / AES_XTS_decrypt calls over the sectors. */
int uaes_xts_sectors(int keybits, const uaes_u8 *keys, uaes_u64 first_sector,
                     size_t sector_bytes, const void *in, size_t len, void *out, int encrypt);

/* micro_aes.c:1164-1179; nonce = 12 bytes; out holds len + 16 (tag appended) */
int uaes_gcm_encrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce,
                     const void *aad, size_t aadlen, const void *in, size_t len, void *out);
/* micro_aes.c:1192-1212; in holds len + 16; UAES_AUTH_ERROR leaves out untouched */
int uaes_gcm_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce,
                     const void *aad, size_t aadlen, const void *in, size_t len, void *out);

/* the reference built with GCM_NONCE_LEN != 12 and / or GCM_TAG_LEN < 16 (micro_aes.h:107-110):
 * a nonce of any other length becomes J0 = GHASH_H({}, nonce) (micro_aes.c:1145-1149), the tag is
 * cut to its first taglen bytes (micro_aes.c:1178, 1204).  out holds len + taglen (encrypt), in holds
 * len + taglen (decrypt). */
int uaes_gcm_encrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, size_t noncelen,
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen);
int uaes_gcm_decrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, size_t noncelen,
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen);

/* ---- AES-GCM-SIV (RFC 8452), micro_aes.c:1474-1516 ------------- */
/* nonce = 12 bytes; out holds len + 16 (tag appended).  Two passes (POLYVAL, then CTR). */
int uaes_gcmsiv_encrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce,
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out);
/* in holds len + 16.  As in the reference the plaintext is produced before the tag is checked;
 * UAES_AUTH_ERROR means it must be discarded. */
int uaes_gcmsiv_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce,
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out);

/* ---- the block-parallel decrypt directions of CBC and CFB ---------- */
/* micro_aes.c:746-782 with the reference's default CS3 ciphertext stealing (CTS = 1): any
 * len >= 16; UAES_DATALENGTH_ERROR below that.  iv = 16 bytes.  (CBC/CFB ENcryption is a serial
 * chain and is not provided.)  A staged copy is made when in == out. */
int uaes_cbc_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *iv,
                     const void *in, size_t len, void *out);
/* cts = 0: the reference built with CTS = 0 (plain CBC: UAES_DATALENGTH_ERROR unless len % 16 == 0,
 * micro_aes.c:757-759); cts = 1 is uaes_cbc_decrypt */
int uaes_cbc_decrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *iv,
                        const void *in, size_t len, void *out, int cts);
/* micro_aes.c:799-845, any length */
int uaes_cfb_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *iv,
                     const void *in, size_t len, void *out);

/* ---- AES-OCB (RFC 7253), micro_aes.c:1779-1813 ------------------- */
/* nonce = 12 bytes, 16-byte tag appended at out + len.  One pass. */
int uaes_ocb_encrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce,
                     const void *aad, size_t aadlen, const void *in, size_t len, void *out);
/* in holds len + 16; the plaintext is written, then the tag is compared (UAES_AUTH_ERROR) */
int uaes_ocb_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce,
                     const void *aad, size_t aadlen, const void *in, size_t len, void *out);

/* the reference built with OCB_TAG_LEN / CCM_TAG_LEN / EAX_TAG_LEN < 16 (micro_aes.h:105, 117, 121): out holds
 * len + taglen (encrypt), in holds len + taglen (decrypt).  OCB: 1..16, the length also enters the nonce block
 * (micro_aes.c:1707); CCM: even, 4..16, it enters the flags of the first MAC block (micro_aes.c:1229); EAX: 1..16,
 * a plain truncation (micro_aes.c:1594, 1638).  Single messages only; the batch calls use 16-byte tags. */
int uaes_ocb_encrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                        const void *in, size_t len, void *out, size_t taglen);
int uaes_ocb_decrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                        const void *in, size_t len, void *out, size_t taglen);
int uaes_ccm_encrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                        const void *in, size_t len, void *out, size_t taglen);
int uaes_ccm_decrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                        const void *in, size_t len, void *out, size_t taglen);
int uaes_eax_encrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                        const void *in, size_t len, void *out, size_t taglen);
int uaes_eax_decrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                        const void *in, size_t len, void *out, size_t taglen);

/* One GCM message sharded over several GPUs (or calls).  Each shard holds a contiguous byte range
 * starting at block `first_block` of the message; all shards but the last are multiples of 16
 * bytes.  uaes_gcm_shard runs the fused CTR + GHASH pass over the shard (encrypt: GHASH over the
 * output; decrypt: GHASH over the input, plaintext written in the same pass) and returns the
 * shard's 16-byte GHASH contribution in partial[] -- host memory, or DEVICE memory, in which case
 * the contribution never leaves the GPU and (in asynchronous mode) the call only enqueues work, so
 * the gather can run device to device.  Gather the contributions (the one 16-byte-per-rank exchange
 * of the path) and let uaes_gcm_combine fold them with the AAD and the lengths into the tag:
 * blocks_after[r] = number of 16-byte blocks of the message after shard r's end (0 for the last
 * shard); partials, blocks_after and tag may each be host or device memory (a device tag in asynchronous mode:
 * the call only enqueues work), nshards is unbounded.
 * A decrypting caller compares that tag with the received one and discards the shards' output on
 * mismatch (the single-call API does this itself).  A host-buffer uaes_gcm_encrypt / _decrypt
 * larger than one staging chunk runs exactly this scheme internally, one shard per chunk. */
int uaes_gcm_shard(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, uaes_u64 first_block,
                   const void *in, size_t len, void *out, int decrypt, uaes_u8 *partial);
int uaes_gcm_combine(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const uaes_u8 *partials, const uaes_u64 *blocks_after, int nshards,
                     uaes_u64 total_len, uaes_u8 *tag);

/* ---- many independent messages per call ------------------------ */
/* The MACs of CCM, EAX and SIV are serial chains inside one message (CBC-MAC micro_aes.c:1222-1256,
 * OMAC :1531-1550, S2V :1325-1359), so one message cannot fill a GPU but a batch can: one message
 * per GPU lane, all under one key.
 * Message i is described by msgs[i]; offsets are relative to the three base pointers and have no
 * alignment requirement (16-byte aligned offsets are fastest).  msgs, aad, in and out may each
 * be host or device memory (when msgs is device memory the other three must be, too).
 *   encrypt: out + out_off holds len + 16 bytes (ciphertext || tag), result = 0
 *   decrypt: in + in_off holds len + 16 bytes; the plaintext is written, then the tag is checked
 *            (micro_aes.c:1304-1312); result = 0 or UAES_AUTH_ERROR per message
 * Return value: 0, or UAES_AUTH_ERROR when at least one message failed, or a negative UAES_E_*. */
typedef struct uaes_msg {
    uaes_u64 in_off, out_off, aad_off;
    unsigned int len, aad_len;         /* payload bytes (without the tag), associated-data bytes */
    uaes_u8  nonce[16];                /* CCM: the first 11 bytes; GCM: 12; EAX: 16; SIV: unused */
    int      result;
    unsigned int reserved;
} uaes_msg;

int uaes_ccm_encrypt_batch(int keybits, const uaes_u8 *key, uaes_msg *msgs, size_t n,
                           const void *aad, const void *in, void *out);
int uaes_ccm_decrypt_batch(int keybits, const uaes_u8 *key, uaes_msg *msgs, size_t n,
                           const void *aad, const void *in, void *out);
/* EAX (not EAX'; 16-byte nonce, 16-byte tag, micro_aes.c:1564-1648): same layout as CCM; decrypt
 * authenticates first and leaves out untouched on UAES_AUTH_ERROR (:1637-1645) */
int uaes_eax_encrypt_batch(int keybits, const uaes_u8 *key, uaes_msg *msgs, size_t n,
                           const void *aad, const void *in, void *out);
int uaes_eax_decrypt_batch(int keybits, const uaes_u8 *key, uaes_msg *msgs, size_t n,
                           const void *aad, const void *in, void *out);
/* SIV (RFC 5297, one AAD unit, micro_aes.c:1372-1410): key = K1 || K2 (2 x keybits/8 bytes).
 *   encrypt: out + out_off holds 16 + len bytes: synthetic IV || ciphertext
 *   decrypt: in + in_off holds IV || ciphertext; the plaintext is written, then the IV is checked
 * in and out ranges of a message must not overlap when encrypting (the output is 16 bytes longer
 * and shifted). */
int uaes_siv_encrypt_batch(int keybits, const uaes_u8 *key, uaes_msg *msgs, size_t n,
                           const void *aad, const void *in, void *out);
int uaes_siv_decrypt_batch(int keybits, const uaes_u8 *key, uaes_msg *msgs, size_t n,
                           const void *aad, const void *in, void *out);
/* small-packet GCM (12-byte nonces, micro_aes.c:1164-1212): same layout as CCM; decrypt authenticates
 * first and leaves out untouched on UAES_AUTH_ERROR.  For few large messages use uaes_gcm_encrypt. */
int uaes_gcm_encrypt_batch(int keybits, const uaes_u8 *key, uaes_msg *msgs, size_t n,
                           const void *aad, const void *in, void *out);
int uaes_gcm_decrypt_batch(int keybits, const uaes_u8 *key, uaes_msg *msgs, size_t n,
                           const void *aad, const void *in, void *out);
/* one message with the reference's argument lists (micro_aes.c:1268-1314, 1564-1648, 1372-1410):
 * a batch of one, i.e. ONE GPU lane -- correct and table-driven, but no faster than a CPU core;
 * use the batch calls for throughput */
int uaes_eax_encrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce,
                     const void *aad, size_t aadlen, const void *in, size_t len, void *out);
int uaes_eax_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce,
                     const void *aad, size_t aadlen, const void *in, size_t len, void *out);
int uaes_siv_encrypt(int keybits, const uaes_u8 *keys, const void *aad, size_t aadlen,
                     const void *in, size_t len, uaes_u8 *iv, void *out);
int uaes_siv_decrypt(int keybits, const uaes_u8 *keys, const uaes_u8 *iv, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out);
int uaes_ccm_encrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce,
                     const void *aad, size_t aadlen, const void *in, size_t len, void *out);
int uaes_ccm_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce,
                     const void *aad, size_t aadlen, const void *in, size_t len, void *out);

/* ---- streaming: init / update / final (the reference has none) ------ */
/* One message fed in pieces.  Every update but the last must be a multiple of 16 bytes; buffers may
 * be host or device memory as everywhere else.  CTR keeps the keystream position; GCM runs one fused
 * CTR+GHASH pass per update and folds the 16-byte contributions on the device, so the number of
 * updates is unbounded.  A decrypting GCM stream hands out plaintext before it is authenticated:
 * discard it when uaes_stream_final returns UAES_AUTH_ERROR. */
typedef struct uaes_stream uaes_stream;
uaes_stream *uaes_stream_ctr(int keybits, const uaes_u8 *key, const uaes_u8 *iv);
uaes_stream *uaes_stream_gcm(int keybits, const uaes_u8 *key, const uaes_u8 *nonce,
                             const void *aad, size_t aadlen, int decrypt);
int  uaes_stream_update(uaes_stream *s, const void *in, size_t len, void *out);
/* GCM encrypt: writes tag[16]; GCM decrypt: compares with tag[16]; CTR: no-op */
int  uaes_stream_final(uaes_stream *s, uaes_u8 *tag);
void uaes_stream_free(uaes_stream *s);

/* ---- synthetic data (bench / tests) ----------------------------------------- */
/* 64-bit word w of dst (little-endian) = splitmix64(seed + first_word + w); dst is DEVICE memory */
int uaes_fill_splitmix64(uaes_u64 seed, uaes_u64 first_word, void *dst, size_t nwords);
/* XOR-fold of a DEVICE buffer of nwords 64-bit words into one word written to *result (host) */
int uaes_xor_fold64(const void *src, size_t nwords, uaes_u64 *result);

#ifdef __cplusplus
}
#endif
#endif /* UAES_B200_H_ */
