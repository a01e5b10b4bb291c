// uaes_bitslice.cuh -- bitsliced AES for the ALU co-runner warps: the CTR-specialised form used by
// ctr_kernel (first half of this file), the general form for data-dependent modes used by
// xts_sectors_hybrid_kernel (bs_encrypt_planes) and the inverse cipher (bs_decrypt_planes).
//
// The table-driven warps of ctr_kernel saturate the shared-memory lookup pipe (32 lookups / clk /
// SM) and leave roughly half of the integer ALU pipe idle once their lookup addresses are built on
// the FMA pipe (IDP.4A, uaes_tables.cuh).  The warps in this file spend that idle ALU time on the
// SAME cipher computed WITHOUT any table: rijndaelEncrypt (micro_aes.c:242-259) on 32 counter
// blocks per thread, one state BIT per register ("bitslicing"):
//
//   plane p = 8*i + b  holds bit b (0 = least significant) of state byte i (byte i = column i/4,
//   row i%4, the reference's state_t order, micro_aes.c:74-77); bit t of the register belongs to
//   the thread's t-th block.
//
// SubBytes (micro_aes.c:187) is the generated 74-LOP3 circuit (uaes_sbox_lut3.cuh), ShiftRows
// (:198) is a renaming of registers, MixColumns (:221) + AddRoundKey (:181) are 96 three-input
// XORs per column, the round-key operand being a 0 / ~0 word read straight from the constant bank.
//
// Counter layout of one pass (1024 consecutive counter values c0 .. c0+1023, c0 % 1024 == 0):
// lane l, slot t  <->  counter c0 + 32*t + l, so that after the final 32x32 bit transposes lane l
// holds block 32*t + l of the pass and every 128-bit load / store of a slot is one coalesced
// 512-byte row across the warp.  Counter bits 0..4 (= lane) are constant in a thread, bits 5..9
// (= t) are the five classic bitslice patterns, bits >= 10 are the same for the whole warp.  Only
// counter bytes 14 and 15 differ inside a pass, so (as in the table-driven warps) rounds 1 and 2
// shrink: 2 S-boxes in round 1 and 8 in round 2 instead of 16 + 16; what the other, warp-uniform
// bytes contribute to rounds 1-2 arrives as 192 precomputed 0 / ~0 words (`um`), which only change
// when counter byte 13 does (every 64 passes).
//
// Everything here is plain integer C++ that also compiles for the host: tests/ runs it on the CPU
// against the oracle (tests/test_bitslice_host.py) before any GPU time is spent.
#pragma once
#include <stdint.h>

#ifndef UAES_HD
#if defined(__CUDACC__)
#define UAES_HD __host__ __device__ __forceinline__
#else
#define UAES_HD static inline
#endif
#endif

#include "uaes_sbox_lut3.cuh"

namespace uaes {

constexpr int kBsMaxRounds = 14;
constexpr int kBsUniformMasks = 192;        // K0, K1 (round 1) and D0..D3 (round 2) as bit masks

// launch constants of the bitsliced warps (kernel parameter -> constant bank)
struct BsKeyPlanes {
    uint32_t k0[16];                        // round key 0, bytes 14 and 15, as 0 / ~0 words
    uint32_t k[kBsMaxRounds - 2][128];      // round keys 3..NR
};

UAES_HD uint32_t bs_mask(uint32_t word, int bit) { return 0u - ((word >> bit) & 1u); }

// MixColumns of one column followed by AddRoundKey.  a[r] = the 8 planes of the byte that
// ShiftRows put in row r; k = 32 key masks of the column (8*row + bit); out likewise.
//   out_r = 2 a_r ^ 3 a_(r+1) ^ a_(r+2) ^ a_(r+3) = xtime(a_r ^ a_(r+1)) ^ a_(r+1) ^ (a_(r+2) ^ a_(r+3))
// xtime on planes: bit b takes bit b-1, bits 1, 3, 4 additionally bit 7 (x^8 = x^4+x^3+x+1).
// ZMASK: bit r set = row r is known to be all zero (the uniform rows of rounds 1-2, folded into k).
template <int ZMASK>
UAES_HD void bs_mix_column(const uint32_t *a0, const uint32_t *a1, const uint32_t *a2,
                           const uint32_t *a3, const uint32_t *k, uint32_t *out)
{
    uint32_t a[4][8], t[4][8];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        a[0][b] = (ZMASK & 1) ? 0u : a0[b];
        a[1][b] = (ZMASK & 2) ? 0u : a1[b];
        a[2][b] = (ZMASK & 4) ? 0u : a2[b];
        a[3][b] = (ZMASK & 8) ? 0u : a3[b];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int b = 0; b < 8; ++b) t[r][b] = a[r][b] ^ a[(r + 1) & 3][b];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const uint32_t u = a[(r + 1) & 3][b] ^ t[(r + 2) & 3][b] ^ k[8 * r + b];
            uint32_t x = t[r][(b + 7) & 7];
            if (b == 1 || b == 3 || b == 4) x ^= t[r][7];
            out[8 * r + b] = u ^ x;
        }
}

// one middle round on the full state: SubBytes, ShiftRows, MixColumns, AddRoundKey(kp)
UAES_HD void bs_round(uint32_t s[128], const uint32_t *kp)
{
#pragma unroll
    for (int i = 0; i < 16; ++i) sbox_bitsliced(s + 8 * i);
    uint32_t o[128];
#pragma unroll
    for (int c = 0; c < 4; ++c)       // ShiftRows: row r of column c comes from column c + r
        bs_mix_column<0>(s + 8 * (4 * c), s + 8 * (4 * ((c + 1) & 3) + 1), s + 8 * (4 * ((c + 2) & 3) + 2),
                         s + 8 * (4 * ((c + 3) & 3) + 3), kp + 32 * c, o + 32 * c);
#pragma unroll
    for (int p = 0; p < 128; ++p) s[p] = o[p];
}

// one round, middle or last (no MixColumns), chosen at run time: lets the caller keep ALL rounds in one
// loop body (instruction-cache footprint: one S-box layer instead of two)
UAES_HD void bs_round_or_last(uint32_t s[128], const uint32_t *kp, bool last)
{
#pragma unroll
    for (int i = 0; i < 16; ++i) sbox_bitsliced(s + 8 * i);
    uint32_t o[128];
    if (!last) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
            bs_mix_column<0>(s + 8 * (4 * c), s + 8 * (4 * ((c + 1) & 3) + 1), s + 8 * (4 * ((c + 2) & 3) + 2),
                             s + 8 * (4 * ((c + 3) & 3) + 3), kp + 32 * c, o + 32 * c);
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int b = 0; b < 8; ++b)
                    o[8 * (4 * c + r) + b] = s[8 * (4 * ((c + r) & 3) + r) + b] ^ kp[8 * (4 * c + r) + b];
    }
#pragma unroll
    for (int p = 0; p < 128; ++p) s[p] = o[p];
}

// last round: SubBytes, ShiftRows, AddRoundKey(kp)
UAES_HD void bs_last_round(uint32_t s[128], const uint32_t *kp)
{
#pragma unroll
    for (int i = 0; i < 16; ++i) sbox_bitsliced(s + 8 * i);
    uint32_t o[128];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int b = 0; b < 8; ++b)
                o[8 * (4 * c + r) + b] = s[8 * (4 * ((c + r) & 3) + r) + b] ^ kp[8 * (4 * c + r) + b];
#pragma unroll
    for (int p = 0; p < 128; ++p) s[p] = o[p];
}

// Rounds 0-2 of a pass.  `lane` = counter bits 0..4, `c14` = counter bits 8..15 of the pass base
// (its bits 0..1 are zero: they are slot bits), k0 = round-key-0 planes of bytes 14 and 15,
// um = the 192 uniform masks: K0 (32), K1 (32), D0..D3 (4 x 32), see bs_uniform_words().
UAES_HD void bs_first_rounds(uint32_t s[128], uint32_t lane, uint32_t c14, const uint32_t *k0,
                             const uint32_t *um)
{
    uint32_t x14[8], x15[8];
    // counter byte 15 = counter bits 0..7: bits 0..4 = lane, bits 5..7 = slot bits 0..2
#pragma unroll
    for (int b = 0; b < 5; ++b) x15[b] = bs_mask(lane, b) ^ k0[8 + b];
    x15[5] = 0xAAAAAAAAu ^ k0[13];
    x15[6] = 0xCCCCCCCCu ^ k0[14];
    x15[7] = 0xF0F0F0F0u ^ k0[15];
    // counter byte 14 = counter bits 8..15: bits 0..1 = slot bits 3..4, the rest from the pass base
    x14[0] = 0xFF00FF00u ^ k0[0];
    x14[1] = 0xFFFF0000u ^ k0[1];
#pragma unroll
    for (int b = 2; b < 8; ++b) x14[b] = bs_mask(c14, b) ^ k0[b];
    sbox_bitsliced(x15);
    sbox_bitsliced(x14);
    // round 1: byte 15 lands in column 0 row 3, byte 14 in column 1 row 2; K0 / K1 carry the rest
    bs_mix_column<0x7>(x15, x15, x15, x15, um + 0, s + 0);
    bs_mix_column<0xB>(x14, x14, x14, x14, um + 32, s + 32);
    // round 2: bytes 0..7 vary, columns 2 and 3 of the state are uniform (inside D0..D3)
#pragma unroll
    for (int i = 0; i < 8; ++i) sbox_bitsliced(s + 8 * i);
    uint32_t o[128];
    bs_mix_column<0xC>(s + 8 * 0, s + 8 * 5, s, s, um + 64, o + 0);         // rows 0,1 <- bytes 0, 5
    bs_mix_column<0x6>(s + 8 * 4, s, s, s + 8 * 3, um + 96, o + 32);        // rows 0,3 <- bytes 4, 3
    bs_mix_column<0x3>(s, s, s + 8 * 2, s + 8 * 7, um + 128, o + 64);       // rows 2,3 <- bytes 2, 7
    bs_mix_column<0x9>(s, s + 8 * 1, s + 8 * 6, s, um + 160, o + 96);       // rows 1,2 <- bytes 1, 6
#pragma unroll
    for (int p = 0; p < 128; ++p) s[p] = o[p];
}

#ifndef UAES_TRANSPOSE_SELECT
#define UAES_TRANSPOSE_SELECT 1
#endif
// 32 x 32 bit transpose: on return bit q of m[t] = bit t of the old m[q]
UAES_HD void bs_transpose32(uint32_t m[32])
{
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const uint32_t lo = __byte_perm(m[k], m[k + 16], 0x5410), hi = __byte_perm(m[k], m[k + 16], 0x7632);
        m[k] = lo; m[k + 16] = hi;
    }
#pragma unroll
    for (int g = 0; g < 32; g += 16)
#pragma unroll
        for (int k = g; k < g + 8; ++k) {
            const uint32_t lo = __byte_perm(m[k], m[k + 8], 0x6240), hi = __byte_perm(m[k], m[k + 8], 0x7351);
            m[k] = lo; m[k + 8] = hi;
        }
#else
    for (int j = 16; j >= 8; j >>= 1) {
        const uint32_t mask = j == 16 ? 0x0000FFFFu : 0x00FF00FFu;
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
            const uint32_t t = ((m[k] >> j) ^ m[k + j]) & mask;
            m[k + j] ^= t; m[k] ^= t << j;
        }
    }
#endif
#pragma unroll
    for (int j = 4; j >= 1; j >>= 1) {
        const uint32_t mask = j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
#if defined(__CUDA_ARCH__) && UAES_TRANSPOSE_SELECT
            // the swap as two bit selects (one LOP3 each) on two shifted copies: 4 instructions per pair instead of 5
            const uint32_t lo = m[k], hi = m[k + j];
            m[k] = lut3<0xE4>(lo, hi << j, mask);            // (a & c) | (b & ~c)
            m[k + j] = lut3<0xE4>(lo >> j, hi, mask);
#else
            const uint32_t t = ((m[k] >> j) ^ m[k + j]) & mask;
            m[k + j] ^= t; m[k] ^= t << j;
#endif
        }
    }
}

// The warp-uniform words behind the 192 masks.  s0..s3 = counter block of the pass base XOR round
// key 0 (bytes 14, 15 are ignored), te(tbl, x) = T-table lookup Te_tbl[x], rk = round keys.
//   K0 = round-1 column 0 without the byte-15 term, K1 = column 1 without the byte-14 term,
//   D_j = round-2 column j without the terms of the varying bytes 0..7.
template <class TE>
UAES_HD void bs_uniform_words(TE te, uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3,
                              const uint32_t *rk, uint32_t w[6])
{
    const uint32_t C2 = te(0, s2 & 255) ^ te(1, (s3 >> 8) & 255) ^ te(2, (s0 >> 16) & 255) ^ te(3, s1 >> 24) ^ rk[6];
    const uint32_t C3 = te(0, s3 & 255) ^ te(1, (s0 >> 8) & 255) ^ te(2, (s1 >> 16) & 255) ^ te(3, s2 >> 24) ^ rk[7];
    w[0] = te(0, s0 & 255) ^ te(1, (s1 >> 8) & 255) ^ te(2, (s2 >> 16) & 255) ^ rk[4];
    w[1] = te(0, s1 & 255) ^ te(1, (s2 >> 8) & 255) ^ te(3, s0 >> 24) ^ rk[5];
    w[2] = te(2, (C2 >> 16) & 255) ^ te(3, C3 >> 24) ^ rk[8];
    w[3] = te(1, (C2 >> 8) & 255) ^ te(2, (C3 >> 16) & 255) ^ rk[9];
    w[4] = te(0, C2 & 255) ^ te(1, (C3 >> 8) & 255) ^ rk[10];
    w[5] = te(0, C3 & 255) ^ te(3, C2 >> 24) ^ rk[11];
}

// ---- the general form: any 32 blocks (ECB-like modes with data-dependent input, e.g. XTS) ------
// All NR + 1 round keys as planes; the state comes from four 32x32 transposes of the blocks' words.
struct BsKeyPlanesFull {
    uint32_t k[kBsMaxRounds + 1][128];      // round keys 0..NR
};

// rijndaelEncrypt (micro_aes.c:242-259) on 32 blocks held as planes.  MERGED: all rounds in one loop body
// (smaller instruction footprint; measured better for CTR / ECB / XTS / CFB, worse for OCB, whose
// co-runner keeps more values live across the rounds).
template <int NR, bool MERGED = true>
UAES_HD void bs_encrypt_planes(uint32_t s[128], const BsKeyPlanesFull &kp)
{
#pragma unroll
    for (int p = 0; p < 128; ++p) s[p] ^= kp.k[0][p];
    if (MERGED) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int r = 1; r <= NR; ++r) bs_round_or_last(s, kp.k[r], r == NR);    // one loop body, see ctr_kernel
    } else {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int r = 1; r < NR; ++r) bs_round(s, kp.k[r]);
        bs_last_round(s, kp.k[NR]);
    }
}

// ---- decryption: the equivalent inverse cipher on planes ---------------------------------------
// dk = the schedule the table-driven decrypt kernels use (uaes_host.c invert_schedule): dk[0] =
// round key NR, dk[r] = InvMixColumns(round key NR - r), dk[NR] = round key 0; with it every round is
// InvShiftRows, InvSubBytes, InvMixColumns, AddRoundKey -- the same result as the reference's
// straight inverse cipher (micro_aes.c:315-332; FIPS-197 5.3.5).
// InvMixColumns = MixColumns after the pre-mix  a0 ^= u, a2 ^= u, a1 ^= v, a3 ^= v  with
// u = 4 (a0 ^ a2), v = 4 (a1 ^ a3)  (the {0e,0b,0d,09} circulant = {02,03,01,01} x {05,00,04,00});
// times-4 on planes: [x6, x6^x7, x0^x7, x1^x6, x2^x6^x7, x3^x7, x4, x5].
UAES_HD void bs_times4(const uint32_t x[8], uint32_t y[8])
{
    const uint32_t w = x[6] ^ x[7];
    y[0] = x[6]; y[1] = w; y[2] = x[0] ^ x[7]; y[3] = x[1] ^ x[6];
    y[4] = x[2] ^ w; y[5] = x[3] ^ x[7]; y[6] = x[4]; y[7] = x[5];
}

UAES_HD void bs_inv_mix_column(const uint32_t *a0, const uint32_t *a1, const uint32_t *a2,
                               const uint32_t *a3, const uint32_t *k, uint32_t *out)
{
    uint32_t e[8], o[8], u[8], v[8], b[4][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { e[i] = a0[i] ^ a2[i]; o[i] = a1[i] ^ a3[i]; }
    bs_times4(e, u);
    bs_times4(o, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        b[0][i] = a0[i] ^ u[i]; b[1][i] = a1[i] ^ v[i]; b[2][i] = a2[i] ^ u[i]; b[3][i] = a3[i] ^ v[i];
    }
    bs_mix_column<0>(b[0], b[1], b[2], b[3], k, out);
}

UAES_HD void bs_inv_round(uint32_t s[128], const uint32_t *kp)
{
#pragma unroll
    for (int i = 0; i < 16; ++i) sbox_inv_bitsliced(s + 8 * i);
    uint32_t o[128];
#pragma unroll
    for (int c = 0; c < 4; ++c)       // InvShiftRows: row r of column c comes from column c - r
        bs_inv_mix_column(s + 8 * (4 * c), s + 8 * (4 * ((c + 3) & 3) + 1), s + 8 * (4 * ((c + 2) & 3) + 2),
                          s + 8 * (4 * ((c + 1) & 3) + 3), kp + 32 * c, o + 32 * c);
#pragma unroll
    for (int p = 0; p < 128; ++p) s[p] = o[p];
}

UAES_HD void bs_inv_last_round(uint32_t s[128], const uint32_t *kp)
{
#pragma unroll
    for (int i = 0; i < 16; ++i) sbox_inv_bitsliced(s + 8 * i);
    uint32_t o[128];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int b = 0; b < 8; ++b)
                o[8 * (4 * c + r) + b] = s[8 * (4 * ((c + 4 - r) & 3) + r) + b] ^ kp[8 * (4 * c + r) + b];
#pragma unroll
    for (int p = 0; p < 128; ++p) s[p] = o[p];
}

// rijndaelDecrypt (micro_aes.c:315-332) on 32 blocks held as planes; kp = planes of dk[0..NR]
template <int NR>
UAES_HD void bs_decrypt_planes(uint32_t s[128], const BsKeyPlanesFull &kp)
{
#pragma unroll
    for (int p = 0; p < 128; ++p) s[p] ^= kp.k[0][p];
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int r = 1; r < NR; ++r) bs_inv_round(s, kp.k[r]);
    bs_inv_last_round(s, kp.k[NR]);
}

UAES_HD void bs_make_key_planes_full(const uint32_t *rk, int rounds, BsKeyPlanesFull *kp)
{
    for (int r = 0; r <= rounds; ++r)
        for (int p = 0; p < 128; ++p) kp->k[r][p] = bs_mask(rk[4 * r + p / 32], p % 32);
}

// host-side (uaes_host.c does the same in C): key planes from the expanded key
UAES_HD void bs_make_key_planes(const uint32_t *rk, int rounds, BsKeyPlanes *kp)
{
    for (int b = 0; b < 8; ++b) {
        kp->k0[b] = bs_mask(rk[3] >> 16, b);
        kp->k0[8 + b] = bs_mask(rk[3] >> 24, b);
    }
    for (int r = 3; r <= rounds; ++r)
        for (int p = 0; p < 128; ++p) kp->k[r - 3][p] = bs_mask(rk[4 * r + p / 32], p % 32);
}

}  // namespace uaes
