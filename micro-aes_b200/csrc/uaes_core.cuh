// uaes_core.cuh -- the AES round loop on sm_100a.
//
// Restates rijndaelEncrypt / rijndaelDecrypt (micro_aes.c:242-259, 315-332) on a state of four
// 32-bit column words (word c = state bytes 4c..4c+3, byte 0 = row 0, matching the reference's
// column-major state_t, micro_aes.c:74-77).  Round keys arrive as a kernel argument, so every
// AddRoundKey operand is read straight from the constant bank by the LOP3 that consumes it.
#pragma once
#include "uaes_tables.cuh"
#include "uaes_launch.h"

namespace uaes {

// ---------------------------------------------------------------- table-driven rounds

// one full encryption round (SubBytes, ShiftRows, MixColumns, AddRoundKey with rk[0..3])
__device__ __forceinline__ void enc_round(uint32_t lb, uint32_t &s0, uint32_t &s1, uint32_t &s2,
                                          uint32_t &s3, const uint32_t *rk)
{
    const uint32_t t0 = lut<0, kOffT0>(lb, s0) ^ lut<1, kOffT1>(lb, s1) ^ lut<2, kOffT2>(lb, s2) ^ lut<3, kOffT3>(lb, s3) ^ rk[0];
    const uint32_t t1 = lut<0, kOffT0>(lb, s1) ^ lut<1, kOffT1>(lb, s2) ^ lut<2, kOffT2>(lb, s3) ^ lut<3, kOffT3>(lb, s0) ^ rk[1];
    const uint32_t t2 = lut<0, kOffT0>(lb, s2) ^ lut<1, kOffT1>(lb, s3) ^ lut<2, kOffT2>(lb, s0) ^ lut<3, kOffT3>(lb, s1) ^ rk[2];
    const uint32_t t3 = lut<0, kOffT0>(lb, s3) ^ lut<1, kOffT1>(lb, s0) ^ lut<2, kOffT2>(lb, s1) ^ lut<3, kOffT3>(lb, s2) ^ rk[3];
    s0 = t0; s1 = t1; s2 = t2; s3 = t3;
}

// bit-select: (a & m) | (b & ~m), one LOP3
__device__ __forceinline__ uint32_t bsel(uint32_t a, uint32_t b, uint32_t m)
{
    return (a & m) | (b & ~m);
}

// last encryption round: SubBytes + ShiftRows only.  S(x) is read out of the T-table that has it
// in the wanted byte lane (Te2 byte 0, Te3 byte 1, Te0 byte 2, Te1 byte 3), then merged.
// Returns the column BEFORE the final AddRoundKey so callers can fold key and data in one XOR.
__device__ __forceinline__ uint32_t enc_last_col(uint32_t lb, uint32_t a, uint32_t b, uint32_t c,
                                                 uint32_t d)
{
    const uint32_t u0 = lut<0, kOffT2>(lb, a), u1 = lut<1, kOffT3>(lb, b);
    const uint32_t u2 = lut<2, kOffT0>(lb, c), u3 = lut<3, kOffT1>(lb, d);
    return bsel(bsel(u0, u1, 0x00ff00ffu), bsel(u2, u3, 0x00ff00ffu), 0x0000ffffu);
}

// rounds FIRST..NR of an encryption whose state already went through AddRoundKey(FIRST-1) and
// rounds < FIRST; `x` is XORed into the result together with the last round key (CTR/XTS fuse
// their data XOR here).  rk points at round key 0.
template <int NR, int FIRST>
__device__ __forceinline__ void enc_finish(uint32_t lb, uint32_t &s0, uint32_t &s1, uint32_t &s2,
                                           uint32_t &s3, const uint32_t *rk, uint32_t x0,
                                           uint32_t x1, uint32_t x2, uint32_t x3)
{
#pragma unroll
    for (int r = FIRST; r < NR; ++r) enc_round(lb, s0, s1, s2, s3, rk + 4 * r);
    const uint32_t o0 = enc_last_col(lb, s0, s1, s2, s3) ^ rk[4 * NR + 0] ^ x0;
    const uint32_t o1 = enc_last_col(lb, s1, s2, s3, s0) ^ rk[4 * NR + 1] ^ x1;
    const uint32_t o2 = enc_last_col(lb, s2, s3, s0, s1) ^ rk[4 * NR + 2] ^ x2;
    const uint32_t o3 = enc_last_col(lb, s3, s0, s1, s2) ^ rk[4 * NR + 3] ^ x3;
    s0 = o0; s1 = o1; s2 = o2; s3 = o3;
}

// the same for N independent blocks, rounds interleaved: N x 16 lookups in flight per warp, which
// is what keeps the lookup pipe fed when few table-driven warps share the SM with the co-runner
template <int NR, int FIRST, int N>
__device__ __forceinline__ void enc_finish_n(uint32_t lb, uint32_t (&s)[N][4], const uint32_t *rk,
                                             const uint4 (&x)[N])
{
#pragma unroll
    for (int r = FIRST; r < NR; ++r)
#pragma unroll
        for (int i = 0; i < N; ++i) enc_round(lb, s[i][0], s[i][1], s[i][2], s[i][3], rk + 4 * r);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint32_t o0 = enc_last_col(lb, s[i][0], s[i][1], s[i][2], s[i][3]) ^ rk[4 * NR + 0] ^ x[i].x;
        const uint32_t o1 = enc_last_col(lb, s[i][1], s[i][2], s[i][3], s[i][0]) ^ rk[4 * NR + 1] ^ x[i].y;
        const uint32_t o2 = enc_last_col(lb, s[i][2], s[i][3], s[i][0], s[i][1]) ^ rk[4 * NR + 2] ^ x[i].z;
        const uint32_t o3 = enc_last_col(lb, s[i][3], s[i][0], s[i][1], s[i][2]) ^ rk[4 * NR + 3] ^ x[i].w;
        s[i][0] = o0; s[i][1] = o1; s[i][2] = o2; s[i][3] = o3;
    }
}

// whole block: out = E_K(s) ^ x
template <int NR>
__device__ __forceinline__ void enc_block(uint32_t lb, uint32_t &s0, uint32_t &s1, uint32_t &s2,
                                          uint32_t &s3, const uint32_t *rk, uint32_t x0 = 0,
                                          uint32_t x1 = 0, uint32_t x2 = 0, uint32_t x3 = 0)
{
    s0 ^= rk[0]; s1 ^= rk[1]; s2 ^= rk[2]; s3 ^= rk[3];
    enc_finish<NR, 1>(lb, s0, s1, s2, s3, rk, x0, x1, x2, x3);
}

// decryption uses the equivalent inverse cipher: dk[] = reversed schedule with InvMixColumns
// applied to the middle round keys (prepared on the host, uaes_host.c), so each round is again
// four lookups per column.  Td2/Td3 = Td0/Td1 rotated by 16 bits.
__device__ __forceinline__ uint32_t rot16(uint32_t x) { return __byte_perm(x, 0, 0x1032); }

__device__ __forceinline__ void dec_round(uint32_t lb, uint32_t &s0, uint32_t &s1, uint32_t &s2,
                                          uint32_t &s3, const uint32_t *dk)
{
    // InvShiftRows: column j takes row r from column j - r
#ifndef UAES_LUT_PRMT
    const uint32_t u0 = lut<0, kOffT0>(lb, s0) ^ lut<1, kOffT1>(lb, s3) ^ lut<2, kOffT2>(lb, s2) ^ lut<3, kOffT3>(lb, s1) ^ dk[0];
    const uint32_t u1 = lut<0, kOffT0>(lb, s1) ^ lut<1, kOffT1>(lb, s0) ^ lut<2, kOffT2>(lb, s3) ^ lut<3, kOffT3>(lb, s2) ^ dk[1];
    const uint32_t u2 = lut<0, kOffT0>(lb, s2) ^ lut<1, kOffT1>(lb, s1) ^ lut<2, kOffT2>(lb, s0) ^ lut<3, kOffT3>(lb, s3) ^ dk[2];
    const uint32_t u3 = lut<0, kOffT0>(lb, s3) ^ lut<1, kOffT1>(lb, s2) ^ lut<2, kOffT2>(lb, s1) ^ lut<3, kOffT3>(lb, s0) ^ dk[3];
    s0 = u0; s1 = u1; s2 = u2; s3 = u3;
    return;
#endif
    const uint32_t t0 = lut<0, kOffT0>(lb, s0) ^ lut<1, kOffT1>(lb, s3) ^ rot16(lut<2, kOffT0>(lb, s2) ^ lut<3, kOffT1>(lb, s1)) ^ dk[0];
    const uint32_t t1 = lut<0, kOffT0>(lb, s1) ^ lut<1, kOffT1>(lb, s0) ^ rot16(lut<2, kOffT0>(lb, s3) ^ lut<3, kOffT1>(lb, s2)) ^ dk[1];
    const uint32_t t2 = lut<0, kOffT0>(lb, s2) ^ lut<1, kOffT1>(lb, s1) ^ rot16(lut<2, kOffT0>(lb, s0) ^ lut<3, kOffT1>(lb, s3)) ^ dk[2];
    const uint32_t t3 = lut<0, kOffT0>(lb, s3) ^ lut<1, kOffT1>(lb, s2) ^ rot16(lut<2, kOffT0>(lb, s1) ^ lut<3, kOffT1>(lb, s0)) ^ dk[3];
    s0 = t0; s1 = t1; s2 = t2; s3 = t3;
}

__device__ __forceinline__ uint32_t dec_last_col(uint32_t lb, uint32_t a, uint32_t b, uint32_t c,
                                                 uint32_t d)
{
    // Td4 holds Si(x) in all four bytes
    const uint32_t u0 = lut<0, kOffTd4>(lb, a), u1 = lut<1, kOffTd4>(lb, b);
    const uint32_t u2 = lut<2, kOffTd4>(lb, c), u3 = lut<3, kOffTd4>(lb, d);
    return bsel(bsel(u0, u1, 0x00ff00ffu), bsel(u2, u3, 0x00ff00ffu), 0x0000ffffu);
}

// whole block: out = D_K(s) ^ x; dk[0..3] is the key added first (= encryption round key NR)
template <int NR>
__device__ __forceinline__ void dec_block(uint32_t lb, uint32_t &s0, uint32_t &s1, uint32_t &s2,
                                          uint32_t &s3, const uint32_t *dk, uint32_t x0 = 0,
                                          uint32_t x1 = 0, uint32_t x2 = 0, uint32_t x3 = 0)
{
    s0 ^= dk[0]; s1 ^= dk[1]; s2 ^= dk[2]; s3 ^= dk[3];
#pragma unroll
    for (int r = 1; r < NR; ++r) dec_round(lb, s0, s1, s2, s3, dk + 4 * r);
    const uint32_t o0 = dec_last_col(lb, s0, s3, s2, s1) ^ dk[4 * NR + 0] ^ x0;
    const uint32_t o1 = dec_last_col(lb, s1, s0, s3, s2) ^ dk[4 * NR + 1] ^ x1;
    const uint32_t o2 = dec_last_col(lb, s2, s1, s0, s3) ^ dk[4 * NR + 2] ^ x2;
    const uint32_t o3 = dec_last_col(lb, s3, s2, s1, s0) ^ dk[4 * NR + 3] ^ x3;
    s0 = o0; s1 = o1; s2 = o2; s3 = o3;
}

// N independent blocks, rounds interleaved (see enc_finish_n): out_i = D_K(s_i) ^ x_i
template <int NR, int N>
__device__ __forceinline__ void dec_block_n(uint32_t lb, uint32_t (&s)[N][4], const uint32_t *dk, const uint4 (&x)[N])
{
#pragma unroll
    for (int i = 0; i < N; ++i) { s[i][0] ^= dk[0]; s[i][1] ^= dk[1]; s[i][2] ^= dk[2]; s[i][3] ^= dk[3]; }
#pragma unroll
    for (int r = 1; r < NR; ++r)
#pragma unroll
        for (int i = 0; i < N; ++i) dec_round(lb, s[i][0], s[i][1], s[i][2], s[i][3], dk + 4 * r);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint32_t o0 = dec_last_col(lb, s[i][0], s[i][3], s[i][2], s[i][1]) ^ dk[4 * NR + 0] ^ x[i].x;
        const uint32_t o1 = dec_last_col(lb, s[i][1], s[i][0], s[i][3], s[i][2]) ^ dk[4 * NR + 1] ^ x[i].y;
        const uint32_t o2 = dec_last_col(lb, s[i][2], s[i][1], s[i][0], s[i][3]) ^ dk[4 * NR + 2] ^ x[i].z;
        const uint32_t o3 = dec_last_col(lb, s[i][3], s[i][2], s[i][1], s[i][0]) ^ dk[4 * NR + 3] ^ x[i].w;
        s[i][0] = o0; s[i][1] = o1; s[i][2] = o2; s[i][3] = o3;
    }
}

// ---------------------------------------------------------------- one-off blocks (no smem)

// Byte-wise AES for the handful of blocks that run on ONE thread (GCM subkey H and E_K(J0), the
// XTS stealing pair, ragged tails): S-box bytes from the constant bank, MixColumns by xtime.
// Follows micro_aes.c:242-259 step by step; `rounds` is a run-time value here.
__device__ inline uint32_t xtime4(uint32_t x)      // four packed bytes times 2 in GF(2^8)
{
    return ((x & 0x7f7f7f7fu) << 1) ^ (((x >> 7) & 0x01010101u) * 0x1bu);
}

__device__ inline uint32_t sub_word(uint32_t x, const ByteTable &sb)
{
    return (uint32_t)sb.v[x & 0xff] | (uint32_t)sb.v[(x >> 8) & 0xff] << 8 |
           (uint32_t)sb.v[(x >> 16) & 0xff] << 16 | (uint32_t)sb.v[x >> 24] << 24;
}

__device__ inline void small_encrypt(const uint32_t *rk, int rounds, uint32_t s[4])
{
    for (int r = 0; r < rounds; ++r) {
        uint32_t a[4], t[4];
        for (int c = 0; c < 4; ++c) a[c] = sub_word(s[c] ^ rk[4 * r + c], c_sbox);
        for (int c = 0; c < 4; ++c)       // ShiftRows: row r of column c comes from column c + r
            t[c] = (a[c] & 0xffu) | (a[(c + 1) & 3] & 0xff00u) | (a[(c + 2) & 3] & 0xff0000u) |
                   (a[(c + 3) & 3] & 0xff000000u);
        if (r + 1 < rounds) {
            for (int c = 0; c < 4; ++c) {  // MixColumns: 2*a ^ 3*rot(a,8) ^ rot(a,16) ^ rot(a,24)
                const uint32_t v = t[c], r8 = __funnelshift_r(v, v, 8);
                t[c] = xtime4(v ^ r8) ^ r8 ^ __funnelshift_r(v, v, 16) ^ __funnelshift_r(v, v, 24);
            }
        }
        for (int c = 0; c < 4; ++c) s[c] = t[c];
    }
    for (int c = 0; c < 4; ++c) s[c] ^= rk[4 * rounds + c];
}

// straight inverse cipher (micro_aes.c:315-332) with the ENCRYPTION schedule rk
__device__ inline void small_decrypt(const uint32_t *rk, int rounds, uint32_t s[4])
{
    for (int c = 0; c < 4; ++c) s[c] ^= rk[4 * rounds + c];
    for (int r = rounds - 1; r >= 0; --r) {
        uint32_t t[4];
        for (int c = 0; c < 4; ++c)       // InvShiftRows: row r of column c comes from column c - r
            t[c] = (s[c] & 0xffu) | (s[(c + 3) & 3] & 0xff00u) | (s[(c + 2) & 3] & 0xff0000u) |
                   (s[(c + 1) & 3] & 0xff000000u);
        for (int c = 0; c < 4; ++c) t[c] = sub_word(t[c], c_inv_sbox) ^ rk[4 * r + c];
        if (r) {
            for (int c = 0; c < 4; ++c) {  // InvMixColumns = MixColumns after the {4,0,5,0} pre-mix
                uint32_t v = t[c];
                const uint32_t u = xtime4(xtime4(v));
                v ^= u ^ __funnelshift_r(u, u, 16);
                const uint32_t r8 = __funnelshift_r(v, v, 8);
                t[c] = xtime4(v ^ r8) ^ r8 ^ __funnelshift_r(v, v, 16) ^ __funnelshift_r(v, v, 24);
            }
        }
        for (int c = 0; c < 4; ++c) s[c] = t[c];
    }
}

}  // namespace uaes
