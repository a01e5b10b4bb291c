/*
 * uaes_launch.h -- internal boundary between the ANSI-C host side (uaes_host.c) and the CUDA
 * translation unit (uaes_kernels.cu).  Plain C: structs of words and the launcher prototypes.
 * Nothing here is public; the public ABI is include/uaes_b200.h.
 */
#ifndef UAES_LAUNCH_H_
#define UAES_LAUNCH_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned int       u32;
typedef unsigned long long u64;

/* An expanded key as the kernels consume it: 4*(rounds+1) little-endian column words.
 * For encryption w[] is the FIPS-197 schedule (micro_aes.c:144-178).  For the table-driven
 * decryption rounds w[] is the equivalent-inverse-cipher schedule: round keys in reverse
 * order with InvMixColumns applied to all but the first and last. */
typedef struct {
    u32 w[60];
    int rounds;              /* 10, 12 or 14 */
} uaes_keysched;

/* CTR / GCM counter block (micro_aes.c:962-971, 1150-1151): bytes 0..8 are fixed, bytes 9..15
 * are a 56-bit big-endian counter (micro_aes.c:421-427). */
typedef struct {
    u32 w0, w1;              /* counter-block bytes 0..7 as little-endian words */
    u32 b8;                  /* byte 8 */
    u64 v0;                  /* 56-bit counter value of keystream block 0 */
} uaes_ctrblock;

/* all launchers return a cudaError_t value as int (0 = success); `stream` is a cudaStream_t */

/* out[i] = in[i] ^ E_K(counter(i)); len in bytes, tail block handled in-kernel */
int uaes_launch_ctr(const uaes_keysched *ks, const uaes_ctrblock *cb, const void *in, void *out,
                    u64 len, void *stream);

/* ECB over len/16 full blocks; encrypt additionally pads and encrypts the last block: pad = 0 zero
 * padding of a len%16 tail (nothing added when len%16 == 0), 1 = PKCS#7, 2 = ISO/IEC 7816-4 (both
 * always add a block; micro_aes.c:610-621) */
int uaes_launch_ecb(const uaes_keysched *ks, int encrypt, int pad, const void *in, void *out, u64 len,
                    void *stream);

/* XTS over nsectors data units of sector_blocks*16 bytes; unit j uses tweak LE128(first_sector+j).
 * ks1 = data key schedule (inverse schedule when !encrypt), ks2 = tweak key (always encryption) */
int uaes_launch_xts_sectors(const uaes_keysched *ks1, const uaes_keysched *ks2, int encrypt,
                            u64 first_sector, u64 sector_blocks, u64 nsectors, const void *in,
                            void *out, void *stream);

/* XTS over ONE data unit with an explicit 16-byte tweak, or over a RANGE of it: in[0] is block
 * `first_block` of the unit (tweak T_0 * alpha^(first_block + k) for block k of the range), len >= 16
 * bytes; a ragged len applies ciphertext stealing to the last two blocks of the range, so only the
 * range that ends the unit may be ragged.  ks1e = data key encryption schedule (needed by the
 * stealing path and the small cipher), ks1 = schedule for the bulk direction. */
int uaes_launch_xts_unit(const uaes_keysched *ks1, const uaes_keysched *ks1e,
                         const uaes_keysched *ks2, int encrypt, const unsigned char tweak[16],
                         u64 first_block, const void *in, void *out, u64 len, void *stream);

/* GCM.  `work` is a device scratch area of at least uaes_gcm_work_bytes(len) bytes.
 * mode 0: CTR over in -> out, GHASH over aad and out;  mode 1: GHASH over aad and in only;
 * mode 2: CTR over in -> out, GHASH over in (decrypting shard).
 * first_block: keystream/GHASH position of in[0] inside a larger message (0 for a whole message).
 * partial_only = 0: tag = E_K(J0) ^ GHASH(aad, data, lengths) written to tag_out (16 B, device);
 * partial_only = 1: the shard's GHASH contribution sum X_i * H^(shard end - i) written instead. */
size_t uaes_gcm_work_bytes(u64 len);
/* aad_state_dev: NULL, or 16 bytes of device memory holding the GHASH state after the AAD (from a
 * mode-1 partial_only pass over a large AAD); aad_dev is then not read, aadlen still feeds the
 * length block */
/* j0 = the pre-counter block J0: nonce || 00000001 for a 12-byte nonce, else GHASH_H(nonce)
 * (uaes_launch_gcm_j0).  taglen = bytes of the tag written to tag_out (GCM_TAG_LEN). */
int uaes_launch_gcm(const uaes_keysched *ks, const unsigned char j0[16], const void *aad_dev,
                    u64 aadlen, const void *aad_state_dev, const void *in, void *out, u64 len,
                    int mode, u64 first_block, int partial_only, void *tag_out, unsigned taglen,
                    void *work, void *stream);
/* J0 = GHASH_H({}, iv) for a nonce of ivlen != 12 bytes (micro_aes.c:1145-1149); iv_dev and out_dev
 * (16 bytes) are device memory */
int uaes_launch_gcm_j0(const uaes_keysched *ks, const void *iv_dev, u64 ivlen, void *out_dev, void *stream);
/* tag of a sharded message from the shards' contributions: partials_dev = nshards x 16 B,
 * after_dev = nshards x u64 (GHASH blocks after the end of each shard), both device memory; any
 * number of shards */
int uaes_launch_gcm_combine(const uaes_keysched *ks, const unsigned char j0[16], const void *aad_dev,
                            u64 aadlen, u64 len, const void *partials_dev, const void *after_dev,
                            unsigned nshards, void *tag_out, unsigned taglen, void *stream);

/* GCM-SIV (micro_aes.c:1418-1516): key derivation blocks (8 bytes each, 2 + Nk/2 of them) into
 * device memory; POLYVAL + tag; CTR with the 32-bit little-endian counter seeded by a tag that
 * lives in device memory */
int uaes_launch_gcmsiv_derive(const uaes_keysched *master, const unsigned char nonce[12],
                              void *out_dev, void *stream);
int uaes_launch_gcmsiv_tag(const uaes_keysched *enc, const unsigned char auth[16],
                           const unsigned char nonce[12], const void *aad_dev, u64 aadlen,
                           const void *aad_state_dev, const void *data, u64 len, int partial_only,
                           void *tag_out, void *work, void *stream);
int uaes_launch_ctr32(const uaes_keysched *enc, const void *tag_dev, const void *in, void *out,
                      u64 len, void *stream);

/* CBC / CFB decrypt (block-parallel directions).  cbc: ks = inverse schedule, kse = encryption
 * schedule (one-thread CS3 pair); cfb: ks = kse = encryption schedule.  tail: CBC = bytes of the
 * short member of the CTS pair (1..16, 0 = no pair), CFB = len % 16. */
int uaes_launch_chain_dec(const uaes_keysched *ks, const uaes_keysched *kse, int cbc,
                          const unsigned char iv[16], const void *in, void *out, u64 nblocks,
                          unsigned tail, void *stream);

/* OCB (micro_aes.c:1693-1814): setup + bulk + finish on one stream.  enc = encryption schedule,
 * bulk = enc when encrypting, the inverse schedule when decrypting.  The 16-byte tag is written to
 * tag_out (device); when decrypting it is the tag computed over the produced plaintext. */
size_t uaes_ocb_work_bytes(void);
int uaes_launch_ocb(const uaes_keysched *enc, const uaes_keysched *bulk, int encrypt,
                    const unsigned char nonce[12], const void *aad_dev, u64 aadlen,
                    const void *in, void *out, u64 len, void *tag_out, unsigned taglen, void *work, void *stream);

/* sum_r partial[r] * H^after[r] (16 bytes, device): many contributions folded into the one of their union */
int uaes_launch_gcm_fold(const uaes_keysched *ks, const void *partials_dev, const void *after_dev,
                         unsigned nshards, void *out_dev, void *stream);

/* CCM over a batch of independent messages, one per lane (uaes_batch.cuh); msgs_dev = device array
 * of uaes_msg records, result fields are written by the kernel */
int uaes_launch_ccm_batch(const uaes_keysched *ks, int decrypt, unsigned taglen, void *msgs_dev, u64 n,
                          const void *aad, const void *in, void *out, void *stream);

/* EAX (mode 1) and SIV (mode 2; ks = S2V key, ks2 = CTR key) over a batch, one message per lane */
int uaes_launch_mac_batch(int mode, const uaes_keysched *ks, const uaes_keysched *ks2, int decrypt, unsigned taglen,
                          void *msgs_dev, u64 n, const void *aad, const void *in, void *out, void *stream);

/* synthetic data + checksum helpers */
int uaes_launch_fill(u64 seed, u64 first_word, void *dst, u64 nwords, void *stream);
int uaes_launch_xor_fold(const void *src, u64 nwords, void *result_dev, void *stream);

/* how the calling thread's last work-queue CTR launch was shared: units served by the table-driven
 * warps and by the bitsliced warps, blocks per unit (bench bookkeeping; synchronises) */
int uaes_launch_ctr_queue_stats(u64 *tt_units, u64 *bs_units, u64 *unit_blocks);
/* kernels launched so far by this process */
u64 uaes_launch_count(void);
/* CTR kernel geometry, see uaes_ctr_tuning() in uaes_b200.h */
void uaes_launch_ctr_tuning(int tt_threads, int bs_permille, long long bs_min_blocks);

#ifdef __cplusplus
}
#endif
#endif
