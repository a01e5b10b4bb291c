// uaes_batch.cuh -- many independent messages per launch, ONE MESSAGE PER LANE:
// CCM, EAX, SIV and small-packet GCM.
//
// Their MACs -- CCM's CBC-MAC (micro_aes.c:1222-1256), the OMACs of EAX (:1531-1550) and the CMACs
// inside SIV's S2V (:1325-1359) -- are serial chains M <- E(M ^ X_i) inside one message (xMac with
// rijndaelEncrypt, :551-570), so a single message can never fill a GPU; thousands of messages can.
// Lane l of a warp walks message l's chain with the same lane-private T-tables as every other
// kernel (each lane's lookups stay in its own bank, so 32 unrelated chains cost exactly what 32
// blocks of one ECB row cost), and produces the message's CTR keystream in the same loop, so the
// payload is read once and written once.  Messages are described by uaes_msg records
// (include/uaes_b200.h); offsets and lengths are arbitrary (16-byte aligned rows take 128-bit
// accesses, anything else goes word- or byte-wise).
#pragma once
#include "uaes_core.cuh"
#include "uaes_gf128.cuh"

namespace uaes {

struct BatchMsg {                 // = uaes_msg of include/uaes_b200.h
    uint64_t in_off, out_off, aad_off;
    uint32_t len, aad_len;
    uint8_t nonce[16];
    int32_t result;
    uint32_t reserved;
};

struct BatchArgs {
    uaes_keysched ks;             // CCM / EAX key; SIV: first key half (S2V)
    uaes_keysched ks2;            // SIV: second key half (CTR)
    BatchMsg *msgs;
    uint64_t n;
    const uint8_t *aad;
    const uint8_t *in;
    uint8_t *out;
    int decrypt;
    uint32_t taglen;              // CCM_TAG_LEN / EAX_TAG_LEN (micro_aes.h:105, 121): bytes of tag behind each message
};

struct Blk { uint32_t w[4]; };

// first n bytes of a block differ?  (memcmp_s(tag, computed, TAG_LEN), micro_aes.c:1308, 1638)
__device__ __forceinline__ uint32_t tag_diff(const Blk &got, const Blk &tag, uint32_t n)
{
    uint32_t d = 0;
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        const uint32_t keep = n >= 4 * i + 4 ? 0xffffffffu : n > 4 * i ? (1u << (8 * (n - 4 * i))) - 1 : 0u;
        d |= (got.w[i] ^ tag.w[i]) & keep;
    }
    return d;
}

constexpr int kBatchThreads = 512;        // 128 registers per lane: two cipher states + descriptors, no spills

// up to 16 bytes from p (n <= 16), zero padded; p has no alignment
__device__ __forceinline__ Blk load_bytes(const uint8_t *p, uint32_t n)
{
    Blk b = {{0, 0, 0, 0}};
    if (n >= 16 && ((size_t)p & 15) == 0) {
        const uint4 v = *(const uint4 *)p;
        b.w[0] = v.x; b.w[1] = v.y; b.w[2] = v.z; b.w[3] = v.w;
    } else if (n >= 16 && ((size_t)p & 3) == 0) {
        const uint32_t *q = (const uint32_t *)p;
        b.w[0] = q[0]; b.w[1] = q[1]; b.w[2] = q[2]; b.w[3] = q[3];
    } else {
        for (uint32_t i = 0; i < n && i < 16; ++i) b.w[i >> 2] |= (uint32_t)p[i] << (8 * (i & 3));
    }
    return b;
}

__device__ __forceinline__ void store_bytes(uint8_t *p, const Blk &b, uint32_t n)
{
    if (n >= 16 && ((size_t)p & 15) == 0) {
        *(uint4 *)p = make_uint4(b.w[0], b.w[1], b.w[2], b.w[3]);
    } else if (n >= 16 && ((size_t)p & 3) == 0) {
        uint32_t *q = (uint32_t *)p;
        q[0] = b.w[0]; q[1] = b.w[1]; q[2] = b.w[2]; q[3] = b.w[3];
    } else {
        for (uint32_t i = 0; i < n && i < 16; ++i) p[i] = (uint8_t)(b.w[i >> 2] >> (8 * (i & 3)));
    }
}

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// M <- E(M ^ X)
template <int NR>
__device__ __forceinline__ void mac_step(uint32_t lb, const uint32_t *rk, Blk &m, const Blk &x)
{
    m.w[0] ^= x.w[0]; m.w[1] ^= x.w[1]; m.w[2] ^= x.w[2]; m.w[3] ^= x.w[3];
    enc_block<NR>(lb, m.w[0], m.w[1], m.w[2], m.w[3], rk);
}

// Keystream block E(counter block) ^ x with the CTR hoisting of ctr_kernel, per lane: while bytes
// 0..14 of the counter block stay the same (256 consecutive blocks of ONE message), round 1 is one
// lookup and round 2 four, the rest coming from five cached words; a change of the upper bytes
// recomputes them (19 lookups per 256 blocks).  16 * (NR - 2) + 5 lookups per block instead of 16 * NR.
template <int NR>
struct LaneCtr {
    uint32_t K0, D0, D1, D2, D3, k0, k1, k2, k3;
    bool valid;
    __device__ __forceinline__ LaneCtr() : valid(false) {}
    __device__ __forceinline__ void block(uint32_t lb, const uint32_t *rk, uint32_t w0, uint32_t w1, uint32_t w2,
                                          uint32_t w3, const Blk &x, Blk &out)
    {
        const uint32_t s3 = w3 ^ rk[3];
        if (!valid || w0 != k0 || w1 != k1 || w2 != k2 || (w3 & 0x00ffffffu) != k3) {
            valid = true; k0 = w0; k1 = w1; k2 = w2; k3 = w3 & 0x00ffffffu;
            const uint32_t s0 = w0 ^ rk[0], s1 = w1 ^ rk[1], s2 = w2 ^ rk[2];
            K0 = lut<0, kOffT0>(lb, s0) ^ lut<1, kOffT1>(lb, s1) ^ lut<2, kOffT2>(lb, s2) ^ rk[4];
            const uint32_t C1 = lut<0, kOffT0>(lb, s1) ^ lut<1, kOffT1>(lb, s2) ^ lut<2, kOffT2>(lb, s3) ^ lut<3, kOffT3>(lb, s0) ^ rk[5];
            const uint32_t C2 = lut<0, kOffT0>(lb, s2) ^ lut<1, kOffT1>(lb, s3) ^ lut<2, kOffT2>(lb, s0) ^ lut<3, kOffT3>(lb, s1) ^ rk[6];
            const uint32_t C3 = lut<0, kOffT0>(lb, s3) ^ lut<1, kOffT1>(lb, s0) ^ lut<2, kOffT2>(lb, s1) ^ lut<3, kOffT3>(lb, s2) ^ rk[7];
            D0 = lut<1, kOffT1>(lb, C1) ^ lut<2, kOffT2>(lb, C2) ^ lut<3, kOffT3>(lb, C3) ^ rk[8];
            D1 = lut<0, kOffT0>(lb, C1) ^ lut<1, kOffT1>(lb, C2) ^ lut<2, kOffT2>(lb, C3) ^ rk[9];
            D2 = lut<0, kOffT0>(lb, C2) ^ lut<1, kOffT1>(lb, C3) ^ lut<3, kOffT3>(lb, C1) ^ rk[10];
            D3 = lut<0, kOffT0>(lb, C3) ^ lut<2, kOffT2>(lb, C1) ^ lut<3, kOffT3>(lb, C2) ^ rk[11];
        }
        const uint32_t c0 = K0 ^ lut<3, kOffT3>(lb, s3);
        uint32_t t0 = D0 ^ lut<0, kOffT0>(lb, c0), t1 = D1 ^ lut<3, kOffT3>(lb, c0);
        uint32_t t2 = D2 ^ lut<2, kOffT2>(lb, c0), t3 = D3 ^ lut<1, kOffT1>(lb, c0);
        enc_finish<NR, 3>(lb, t0, t1, t2, t3, rk, x.w[0], x.w[1], x.w[2], x.w[3]);
        out.w[0] = t0; out.w[1] = t1; out.w[2] = t2; out.w[3] = t3;
    }
};

// ---------------------------------------------------------------- CCM (micro_aes.c:1219-1315)
// CCM_NONCE_LEN = 11, CCM_TAG_LEN = 16 (micro_aes.h:104-105): iv = 03 || nonce || 00000000
template <int NR>
__global__ void __launch_bounds__(kBatchThreads, 1) ccm_batch_kernel(const __grid_constant__ BatchArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_tables<true>(dyn);
    const uint32_t *rk = a.ks.w;
    const uint64_t stride = (uint64_t)gridDim.x * kBatchThreads;

    for (uint64_t mi = (uint64_t)blockIdx.x * kBatchThreads + threadIdx.x; mi < a.n; mi += stride) {
        const BatchMsg d = a.msgs[mi];
        const uint8_t *src = a.in + d.in_off;
        uint8_t *dst = a.out + d.out_off;
        const uint8_t *aad = a.aad + d.aad_off;

        Blk iv;                                       // micro_aes.c:1273-1275
        iv.w[0] = 3u | (uint32_t)d.nonce[0] << 8 | (uint32_t)d.nonce[1] << 16 | (uint32_t)d.nonce[2] << 24;
        iv.w[1] = (uint32_t)d.nonce[3] | (uint32_t)d.nonce[4] << 8 | (uint32_t)d.nonce[5] << 16 | (uint32_t)d.nonce[6] << 24;
        iv.w[2] = (uint32_t)d.nonce[7] | (uint32_t)d.nonce[8] << 8 | (uint32_t)d.nonce[9] << 16 | (uint32_t)d.nonce[10] << 24;
        iv.w[3] = 0;

        // ---- CCMtag, header part (micro_aes.c:1228-1251)
        Blk m = iv;
        m.w[0] |= (a.taglen - 2) << 2;                // :1229: M' = (CCM_TAG_LEN - 2) / 2 in bits 3..5
        m.w[3] ^= bswap32(d.len);                     // xorBEint(M, ptextLen, LAST), :1230
        Blk A = {{0, 0, 0, 0}};
        uint32_t head = 0;
        if (d.aad_len) {
            m.w[0] |= 0x40;
            enc_block<NR>(lb, m.w[0], m.w[1], m.w[2], m.w[3], rk);          // :1235
            uint32_t p;
            if (d.aad_len > 0xFEFFu) {                // :1236-1240: FF FE || BE32(aDataLen)
                A.w[0] = 0xFEFFu | (d.aad_len >> 24) << 16 | ((d.aad_len >> 16) & 255u) << 24;
                A.w[1] = ((d.aad_len >> 8) & 255u) | (d.aad_len & 255u) << 8;
                p = 6;
            } else {
                A.w[0] = (d.aad_len >> 8) | (d.aad_len & 255u) << 8;
                p = 2;
            }
            head = 16 - p < d.aad_len ? 16 - p : d.aad_len;
            for (uint32_t i = 0; i < head; ++i) A.w[(p + i) >> 2] |= (uint32_t)aad[i] << (8 * ((p + i) & 3));
        }
        mac_step<NR>(lb, rk, m, A);                   // :1247 (a zero block when there is no AAD)
        for (uint32_t o = head; o < d.aad_len; o += 16)
            mac_step<NR>(lb, rk, m, load_bytes(aad + o, d.aad_len - o));

        // ---- payload: CBC-MAC over the plaintext (:1252) and CTR from iv + 1 (:939-941) in one walk
        uint32_t ctr = 1;
        LaneCtr<NR> lc;
        for (uint32_t o = 0; o < d.len; o += 16, ++ctr) {
            const uint32_t nb = d.len - o < 16 ? d.len - o : 16;
            Blk x = load_bytes(src + o, nb);
            Blk y;
            lc.block(lb, rk, iv.w[0], iv.w[1], iv.w[2], bswap32(ctr), x, y);
            if (a.decrypt) {
                if (nb < 16) {                        // the MAC sees the plaintext zero padded
                    const uint32_t keep = nb & 3 ? (1u << (8 * (nb & 3))) - 1 : 0;
                    for (uint32_t i = 0; i < 4; ++i)
                        if (i > (nb >> 2)) y.w[i] = 0; else if (i == (nb >> 2)) y.w[i] &= keep;
                }
                mac_step<NR>(lb, rk, m, y);
            } else {
                mac_step<NR>(lb, rk, m, x);
            }
            store_bytes(dst + o, y, nb);
        }

        // ---- tag = E(iv) ^ CBC-MAC (:1254-1255)
        Blk s0 = iv;
        enc_block<NR>(lb, s0.w[0], s0.w[1], s0.w[2], s0.w[3], rk);
        Blk tag;
        for (int i = 0; i < 4; ++i) tag.w[i] = m.w[i] ^ s0.w[i];
        if (a.decrypt) {                              // :1308-1313; the plaintext stays (SABOTAGE is off)
            const Blk got = load_bytes(src + d.len, a.taglen);
            a.msgs[mi].result = tag_diff(got, tag, a.taglen) ? 0x1A : 0;
        } else {
            store_bytes(dst + d.len, tag, a.taglen);  // :1281
            a.msgs[mi].result = 0;
        }
    }
}

// ---------------------------------------------------------------- CMAC pieces for EAX and SIV

// big-endian 128-bit value times x with the 0x87 fold: doubleBblock, micro_aes.c:434-444
__device__ __forceinline__ Blk dbl_be(const Blk &b)
{
    const uint32_t a0 = bswap32(b.w[0]), a1 = bswap32(b.w[1]), a2 = bswap32(b.w[2]), a3 = bswap32(b.w[3]);
    Blk r;
    r.w[0] = bswap32(a0 << 1 | a1 >> 31); r.w[1] = bswap32(a1 << 1 | a2 >> 31);
    r.w[2] = bswap32(a2 << 1 | a3 >> 31); r.w[3] = bswap32((a3 << 1) ^ ((a0 >> 31) * 0x87u));
    return r;
}

__device__ __forceinline__ void xor_blk(Blk &a, const Blk &b)
{
    a.w[0] ^= b.w[0]; a.w[1] ^= b.w[1]; a.w[2] ^= b.w[2]; a.w[3] ^= b.w[3];
}

// the block as a little-endian 128-bit integer shifted by whole bytes (0..16); byte j moves to
// byte j + n (left) or j - n (right).  Rare (once per SIV message): plain loops.
__device__ inline Blk shift_bytes(const Blk &a, int n, bool left)
{
    Blk r = {{0, 0, 0, 0}};
    for (int j = 0; j < 16; ++j) {
        const int t = left ? j + n : j - n;
        if (t >= 0 && t < 16) r.w[t >> 2] |= ((a.w[j >> 2] >> (8 * (j & 3))) & 255u) << (8 * (t & 3));
    }
    return r;
}

// getSubkeys with quad = 1 (micro_aes.c:593-604): K1 = 2 E(0), K2 = 4 E(0)
template <int NR>
__device__ __forceinline__ void cmac_subkeys(uint32_t lb, const uint32_t *rk, Blk &k1, Blk &k2)
{
    Blk l = {{0, 0, 0, 0}};
    enc_block<NR>(lb, l.w[0], l.w[1], l.w[2], l.w[3], rk);
    k1 = dbl_be(l);
    k2 = dbl_be(k1);
}

// cMac (micro_aes.c:576-590): CMAC of data[0..n) continued from state m.  `fix` is XORed into the
// message's last 16 bytes on the fly (SIV's xorend; pass zero otherwise; needs n >= 16 if nonzero).
template <int NR>
__device__ __forceinline__ void cmac_continue(uint32_t lb, const uint32_t *rk, const Blk &k1, const Blk &k2,
                                              const uint8_t *data, uint32_t n, Blk &m, const Blk &fix, bool has_fix)
{
    const uint32_t s = n ? (n - 1) % 16 + 1 : 0;               // bytes in the last block
    const uint32_t body = n - s;                                // multiple of 16
    for (uint32_t o = 0; o < body; o += 16) {
        Blk x = load_bytes(data + o, 16);
        if (has_fix && s < 16 && o + 16 == body) xor_blk(x, shift_bytes(fix, (int)s, true));
        mac_step<NR>(lb, rk, m, x);
    }
    Blk last = s ? load_bytes(data + body, s) : Blk{{0, 0, 0, 0}};
    if (has_fix) xor_blk(last, s < 16 ? shift_bytes(fix, 16 - (int)s, false) : fix);
    if (s < 16) { last.w[s >> 2] ^= 0x80u << (8 * (s & 3)); xor_blk(last, k2); }
    else xor_blk(last, k1);
    mac_step<NR>(lb, rk, m, last);
}

// oMac without EAXP (micro_aes.c:1531-1550): CMAC([t]_128 || data)
template <int NR>
__device__ __forceinline__ Blk omac(uint32_t lb, const uint32_t *rk, const Blk &k1, const Blk &k2, uint32_t t,
                                    const uint8_t *data, uint32_t n)
{
    Blk m = {{0, 0, 0, 0}};
    if (n == 0) m = k1;
    m.w[3] ^= t << 24;
    enc_block<NR>(lb, m.w[0], m.w[1], m.w[2], m.w[3], rk);
    if (n) cmac_continue<NR>(lb, rk, k1, k2, data, n, m, Blk{{0, 0, 0, 0}}, false);
    return m;
}

// out = in ^ E(ctr0 + k), the counter a 56-bit big-endian integer in bytes 9..15 (incBlock with
// index LAST, micro_aes.c:421-427), no pre-increment (CTR_DEFAULT / SIV_CTR, :931-942)
template <int NR>
__device__ __forceinline__ void ctr_walk(uint32_t lb, const uint32_t *rk, const Blk &c0, const uint8_t *src,
                                         uint8_t *dst, uint32_t n)
{
    const uint64_t v0 = ((uint64_t)(bswap32(c0.w[2]) & 0x00ffffffu) << 32) | bswap32(c0.w[3]);
    uint64_t k = 0;
    LaneCtr<NR> lc;
    for (uint32_t o = 0; o < n; o += 16, ++k) {
        const uint32_t nb = n - o < 16 ? n - o : 16;
        const Blk x = load_bytes(src + o, nb);
        Blk ks;
        uint32_t w2, w3;
        ctr_words(c0.w[2] & 255u, (v0 + k) & kMask56, w2, w3);
        lc.block(lb, rk, c0.w[0], c0.w[1], w2, w3, x, ks);
        store_bytes(dst + o, ks, nb);
    }
}

// ---------------------------------------------------------------- EAX (micro_aes.c:1518-1649)
// EAX_NONCE_LEN = 16, EAX_TAG_LEN = 16 (micro_aes.h:119-121), not EAX'
template <int NR>
__global__ void __launch_bounds__(kBatchThreads, 1) eax_batch_kernel(const __grid_constant__ BatchArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_tables<true>(dyn);
    const uint32_t *rk = a.ks.w;
    const uint64_t stride = (uint64_t)gridDim.x * kBatchThreads;
    Blk k1, k2;
    cmac_subkeys<NR>(lb, rk, k1, k2);

    for (uint64_t mi = (uint64_t)blockIdx.x * kBatchThreads + threadIdx.x; mi < a.n; mi += stride) {
        const BatchMsg d = a.msgs[mi];
        const uint8_t *src = a.in + d.in_off;
        uint8_t *dst = a.out + d.out_off;
        const Blk N = omac<NR>(lb, rk, k1, k2, 0, d.nonce, 16);                   // :1578
        Blk tag = omac<NR>(lb, rk, k1, k2, 1, a.aad + d.aad_off, d.aad_len);        // :1591
        xor_blk(tag, N);
        if (!a.decrypt) {
            ctr_walk<NR>(lb, rk, N, src, dst, d.len);                             // :1584, counter starts AT N
            xor_blk(tag, omac<NR>(lb, rk, k1, k2, 2, dst, d.len));                // :1593, over the ciphertext
            store_bytes(dst + d.len, tag, a.taglen);                              // :1594
            a.msgs[mi].result = 0;
        } else {                                                                  // :1625-1647: verify, then decrypt
            xor_blk(tag, omac<NR>(lb, rk, k1, k2, 2, src, d.len));
            const Blk got = load_bytes(src + d.len, a.taglen);
            const uint32_t diff = tag_diff(got, tag, a.taglen);                   // :1638
            a.msgs[mi].result = diff ? 0x1A : 0;
            if (!diff) ctr_walk<NR>(lb, rk, N, src, dst, d.len);
        }
    }
}

// ---------------------------------------------------------------- SIV (micro_aes.c:1317-1411)
// RFC 5297 with one AAD unit.  Layout of a message: IV (16 bytes) || ciphertext.

// S2V, micro_aes.c:1325-1359
template <int NR>
__device__ __forceinline__ Blk s2v(uint32_t lb, const uint32_t *rk, const Blk &k1, const Blk &k2,
                                   const uint8_t *aad, uint32_t aad_len, const uint8_t *pt, uint32_t n)
{
    Blk y = k1;                                                   // Y_0 = CMAC(0^128) = E(K1), :1332
    enc_block<NR>(lb, y.w[0], y.w[1], y.w[2], y.w[3], rk);
    if (aad_len) {                                                // :1338-1344
        Blk t = {{0, 0, 0, 0}};
        cmac_continue<NR>(lb, rk, k1, k2, aad, aad_len, t, Blk{{0, 0, 0, 0}}, false);
        y = dbl_be(y);
        xor_blk(y, t);
    }
    Blk v = {{0, 0, 0, 0}};
    if (n >= 16) {                                                // CMAC(pt xorend Y), :1350-1358
        cmac_continue<NR>(lb, rk, k1, k2, pt, n, v, y, true);
    } else {                                                      // CMAC(dbl(Y) ^ pad(pt)), :1345-1349
        y = dbl_be(y);
        xor_blk(y, load_bytes(pt, n));
        y.w[n >> 2] ^= 0x80u << (8 * (n & 3));
        xor_blk(y, k1);
        mac_step<NR>(lb, rk, v, y);
    }
    return v;
}

template <int NR>
__global__ void __launch_bounds__(kBatchThreads, 1) siv_batch_kernel(const __grid_constant__ BatchArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_tables<true>(dyn);
    const uint32_t *rk1 = a.ks.w, *rk2 = a.ks2.w;
    const uint64_t stride = (uint64_t)gridDim.x * kBatchThreads;
    Blk k1, k2;
    cmac_subkeys<NR>(lb, rk1, k1, k2);

    for (uint64_t mi = (uint64_t)blockIdx.x * kBatchThreads + threadIdx.x; mi < a.n; mi += stride) {
        const BatchMsg d = a.msgs[mi];
        const uint8_t *src = a.in + d.in_off;
        uint8_t *dst = a.out + d.out_off;
        const uint8_t *aad = a.aad + d.aad_off;
        if (!a.decrypt) {                                         // :1372-1382
            const Blk v = s2v<NR>(lb, rk1, k1, k2, aad, d.aad_len, src, d.len);
            Blk q = v;
            q.w[2] &= ~0x80u; q.w[3] &= ~0x80u;                   // c[8] &= 0x7F, c[12] &= 0x7F, :931-934
            ctr_walk<NR>(lb, rk2, q, src, dst + 16, d.len);
            store_bytes(dst, v, 16);                              // written last: in == out stays correct
            a.msgs[mi].result = 0;
        } else {                                                  // :1394-1410: decrypt, then compare
            const Blk iv = load_bytes(src, 16);
            Blk q = iv;
            q.w[2] &= ~0x80u; q.w[3] &= ~0x80u;
            ctr_walk<NR>(lb, rk2, q, src + 16, dst, d.len);
            const Blk v = s2v<NR>(lb, rk1, k1, k2, aad, d.aad_len, dst, d.len);
            const uint32_t diff = (iv.w[0] ^ v.w[0]) | (iv.w[1] ^ v.w[1]) | (iv.w[2] ^ v.w[2]) | (iv.w[3] ^ v.w[3]);
            a.msgs[mi].result = diff ? 0x1A : 0;
        }
    }
}

// ---------------------------------------------------------------- GCM (micro_aes.c:1124-1213)
// Small-packet GCM: the lane runs the message's CTR and its GHASH chain G <- H * (G ^ X_i)
// (xMac with mulGF128, :551-570, 476-493).  The per-block product is the carry-less multiply built
// from integer multiplies (gf_mul_fast, uaes_gf128.cuh): its 144 multiplies run on the FMA pipe
// next to the cipher's lookups, and no GHASH table has to be built per message.
__device__ __forceinline__ void ghash_step(Gf &g, const Gf &H, const Blk &x)
{
    const Gf v = gf_from_words(x.w[0], x.w[1], x.w[2], x.w[3]);
    g.hi ^= v.hi; g.lo ^= v.lo;
    g = gf_mul_fast(g, H);
}

__device__ __forceinline__ void ghash_bytes(Gf &g, const Gf &H, const uint8_t *p, uint32_t n)
{
    for (uint32_t o = 0; o < n; o += 16) ghash_step(g, H, load_bytes(p + o, n - o < 16 ? n - o : 16));
}

template <int NR>
__global__ void __launch_bounds__(kBatchThreads, 1) gcm_batch_kernel(const __grid_constant__ BatchArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_tables<true>(dyn);
    const uint32_t *rk = a.ks.w;
    const uint64_t stride = (uint64_t)gridDim.x * kBatchThreads;
    Blk h = {{0, 0, 0, 0}};
    enc_block<NR>(lb, h.w[0], h.w[1], h.w[2], h.w[3], rk);        // H = E(0), :1144
    const Gf H = gf_from_words(h.w[0], h.w[1], h.w[2], h.w[3]);

    for (uint64_t mi = (uint64_t)blockIdx.x * kBatchThreads + threadIdx.x; mi < a.n; mi += stride) {
        const BatchMsg d = a.msgs[mi];
        const uint8_t *src = a.in + d.in_off;
        uint8_t *dst = a.out + d.out_off;
        Blk j0;                                                   // nonce || 00000001, :1150-1151
        j0.w[0] = (uint32_t)d.nonce[0] | (uint32_t)d.nonce[1] << 8 | (uint32_t)d.nonce[2] << 16 | (uint32_t)d.nonce[3] << 24;
        j0.w[1] = (uint32_t)d.nonce[4] | (uint32_t)d.nonce[5] << 8 | (uint32_t)d.nonce[6] << 16 | (uint32_t)d.nonce[7] << 24;
        j0.w[2] = (uint32_t)d.nonce[8] | (uint32_t)d.nonce[9] << 8 | (uint32_t)d.nonce[10] << 16 | (uint32_t)d.nonce[11] << 24;
        j0.w[3] = 0x01000000u;
        Gf g{0, 0};
        ghash_bytes(g, H, a.aad + d.aad_off, d.aad_len);          // gHash, :1134
        if (a.decrypt) ghash_bytes(g, H, src, d.len);             // :1199: over the received ciphertext
        else {
            uint32_t ctr = 2;                                     // CCM_GCM pre-increment: J0 + 1, :939-941
            LaneCtr<NR> lc;
            for (uint32_t o = 0; o < d.len; o += 16, ++ctr) {
                const uint32_t nb = d.len - o < 16 ? d.len - o : 16;
                const Blk x = load_bytes(src + o, nb);
                Blk y;
                lc.block(lb, rk, j0.w[0], j0.w[1], j0.w[2], bswap32(ctr), x, y);
                if (nb < 16) {                                    // GHASH sees the ciphertext zero padded
                    const uint32_t keep = nb & 3 ? (1u << (8 * (nb & 3))) - 1 : 0;
                    for (uint32_t i = 0; i < 4; ++i)
                        if (i > (nb >> 2)) y.w[i] = 0; else if (i == (nb >> 2)) y.w[i] &= keep;
                }
                store_bytes(dst + o, y, nb);
                ghash_step(g, H, y);
            }
        }
        g.hi ^= (uint64_t)d.aad_len * 8; g.lo ^= (uint64_t)d.len * 8;      // length block, :1130-1132
        g = gf_mul_fast(g, H);
        Blk tag = j0;
        enc_block<NR>(lb, tag.w[0], tag.w[1], tag.w[2], tag.w[3], rk);    // E(J0), :1173
        Blk gw;
        gf_to_words(g, gw.w[0], gw.w[1], gw.w[2], gw.w[3]);
        xor_blk(tag, gw);
        if (!a.decrypt) {
            store_bytes(dst + d.len, tag, 16);
            a.msgs[mi].result = 0;
        } else {                                                  // :1204-1209: verify, then decrypt
            const Blk got = load_bytes(src + d.len, 16);
            const uint32_t diff = (got.w[0] ^ tag.w[0]) | (got.w[1] ^ tag.w[1]) | (got.w[2] ^ tag.w[2]) | (got.w[3] ^ tag.w[3]);
            a.msgs[mi].result = diff ? 0x1A : 0;
            if (!diff) {
                Blk c0 = j0;
                c0.w[3] = 0x02000000u;                            // J0 + 1
                ctr_walk<NR>(lb, rk, c0, src, dst, d.len);
            }
        }
    }
}

template <int NR, int MODE>
static cudaError_t launch_batch_nr(const BatchArgs &a, cudaStream_t st)
{
    auto kernel = MODE == 1 ? eax_batch_kernel<NR> : MODE == 2 ? siv_batch_kernel<NR> : gcm_batch_kernel<NR>;
    cudaError_t e = opt_in_smem(kernel);
    if (e != cudaSuccess) return e;
    const uint64_t need = (a.n + kBatchThreads - 1) / kBatchThreads, sms = (uint64_t)sm_count();
    kernel<<<(unsigned)(need < sms ? need : sms), kBatchThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

template <int NR>
static cudaError_t launch_ccm_batch_nr(const BatchArgs &a, cudaStream_t st)
{
    cudaError_t e = opt_in_smem(ccm_batch_kernel<NR>);
    if (e != cudaSuccess) return e;
    const uint64_t need = (a.n + kBatchThreads - 1) / kBatchThreads, sms = (uint64_t)sm_count();
    ccm_batch_kernel<NR><<<(unsigned)(need < sms ? need : sms), kBatchThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace uaes

extern "C" int uaes_launch_ccm_batch(const uaes_keysched *ks, int decrypt, unsigned taglen, void *msgs_dev, u64 n,
                                     const void *aad, const void *in, void *out, void *stream)
{
    if (n == 0) return 0;
    uaes::BatchArgs a;
    a.ks = *ks; a.ks2 = *ks;
    a.msgs = (uaes::BatchMsg *)msgs_dev; a.n = n;
    a.aad = (const uint8_t *)aad; a.in = (const uint8_t *)in; a.out = (uint8_t *)out;
    a.decrypt = decrypt; a.taglen = taglen;
    cudaStream_t st = (cudaStream_t)stream;
    switch (ks->rounds) {
    case 10: return (int)uaes::launch_ccm_batch_nr<10>(a, st);
    case 12: return (int)uaes::launch_ccm_batch_nr<12>(a, st);
    case 14: return (int)uaes::launch_ccm_batch_nr<14>(a, st);
    }
    return (int)cudaErrorInvalidValue;
}

// mode: 1 = EAX (ks), 2 = SIV (ks = S2V key, ks2 = CTR key), 3 = GCM (ks)
extern "C" int uaes_launch_mac_batch(int mode, const uaes_keysched *ks, const uaes_keysched *ks2, int decrypt,
                                     unsigned taglen, void *msgs_dev, u64 n, const void *aad, const void *in, void *out,
                                     void *stream)
{
    if (n == 0) return 0;
    uaes::BatchArgs a;
    a.ks = *ks; a.ks2 = ks2 ? *ks2 : *ks;
    a.msgs = (uaes::BatchMsg *)msgs_dev; a.n = n;
    a.aad = (const uint8_t *)aad; a.in = (const uint8_t *)in; a.out = (uint8_t *)out;
    a.decrypt = decrypt; a.taglen = taglen;
    cudaStream_t st = (cudaStream_t)stream;
    switch (ks->rounds * 4 + mode) {
    case 41: return (int)uaes::launch_batch_nr<10, 1>(a, st);
    case 49: return (int)uaes::launch_batch_nr<12, 1>(a, st);
    case 57: return (int)uaes::launch_batch_nr<14, 1>(a, st);
    case 42: return (int)uaes::launch_batch_nr<10, 2>(a, st);
    case 50: return (int)uaes::launch_batch_nr<12, 2>(a, st);
    case 58: return (int)uaes::launch_batch_nr<14, 2>(a, st);
    case 43: return (int)uaes::launch_batch_nr<10, 3>(a, st);
    case 51: return (int)uaes::launch_batch_nr<12, 3>(a, st);
    case 59: return (int)uaes::launch_batch_nr<14, 3>(a, st);
    }
    return (int)cudaErrorInvalidValue;
}
