// uaes_bitslice8.cuh -- the NARROW bitsliced form: 8 blocks per thread, 32 state registers.
//
// uaes_bitslice.cuh keeps one state bit of 32 blocks per register: 128 registers of state, 224 per
// thread with the temporaries -- four co-runner warps own 44 % of the SM's register file, there can
// be only one of them per scheduler, and one round is 1 600 instructions of code.  Here a register
// holds one bit position of one state ROW for all four columns of 8 blocks:
//
//   plane j = 8*r + b  holds bit b (0 = least significant) of the state bytes in row r;
//   bit 8*c + t of the register belongs to column c of the thread's t-th block (t = 0..7).
//
// (state byte i = column i/4, row i%4, micro_aes.c:74-77; word c of a block = column c, so plane j
// is simply bit j of the four words, and one 32x32 bit transpose of M[8*c + t] = word c of block t
// converts between blocks and planes in either direction.)
//
// SubBytes (micro_aes.c:187) = the same generated 74-LOP3 circuit, once per ROW (it covers 4 columns
// x 8 blocks = 32 S-boxes, like one call in the wide form); ShiftRows (:198) = a byte rotation of the
// row's registers by r (one PRMT each, 24 per round); MixColumns (:221) + AddRoundKey (:181) = the
// wide form's column routine applied ONCE, because the four columns sit side by side in each
// register.  Per block the instruction count is that of the wide form plus the 24 PRMTs (+6 %), but a
// thread needs 32 + 32 + temporaries instead of 128 + 32 + temporaries registers and one round is
// 440 instructions: two such warps fit per scheduler next to the table-driven ones, and the whole
// round loop stays in the instruction cache.  (Measured, DESIGN.md 5.1: this pays for CTR, whose
// hoisted rounds 1-2 make a bitsliced block cheap; as GCM or XTS co-runner the wide form stays ahead.)
//
// Counter layout of one pass = one GROUP (the 256 counters that share bytes 0..14): lane l, slot t
// <-> counter byte 15 = 32*t + l, so slot t of all lanes is one coalesced 512-byte row.  Rounds 1-2
// factor exactly as for the table-driven warps (uaes_kernels.cu): the state entering round 3 is
// D_j ^ U_j(byte 15), D warp-uniform per group, U a function of byte 15 alone -- per thread a constant
// for 2^40 blocks.  U is kept as 32 planes (computed once), D arrives as 32 mask words per group.
//
// Plain integer C++ that also compiles for the host (tests/test_bitslice_host.py).
#pragma once
#include "uaes_bitslice.cuh"

namespace uaes {

struct alignas(16) BsKeyPlanes8 {           // 16-byte aligned inside the kernel parameters: the key words load as LDC.128
    uint32_t k[kBsMaxRounds - 2][32];       // round keys 3..NR: word j, byte c = 0xFF iff bit j of key word c is set
};

// byte c of the result = 0xFF iff bit j of w[c] is set: four column words -> the plane-j word of a value
// that is the same for all 8 blocks of the thread
UAES_HD uint32_t bs8_spread(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int j)
{
    const uint32_t z = ((w0 >> j) & 1u) | (((w1 >> j) & 1u) << 8) | (((w2 >> j) & 1u) << 16) | (((w3 >> j) & 1u) << 24);
    return z * 0xFFu;
}

// ShiftRows on a row register: the new column c takes the old column c + R
template <int R> UAES_HD uint32_t bs8_shift_row(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return R == 0 ? x : __byte_perm(x, 0, R == 1 ? 0x0321 : R == 2 ? 0x1032 : 0x2103);
#else
    return R == 0 ? x : (x >> (8 * R)) | (x << (32 - 8 * R));
#endif
}

// one round, middle or last (no MixColumns), chosen at run time so that all rounds share one loop body
UAES_HD void bs8_round_or_last(uint32_t s[32], const uint32_t *kp, bool last)
{
#pragma unroll
    for (int r = 0; r < 4; ++r) sbox_bitsliced(s + 8 * r);
    uint32_t a[4][8], o[32];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        a[0][b] = bs8_shift_row<0>(s[b]);      a[1][b] = bs8_shift_row<1>(s[8 + b]);
        a[2][b] = bs8_shift_row<2>(s[16 + b]); a[3][b] = bs8_shift_row<3>(s[24 + b]);
    }
    // the round's 32 key words as eight 128-bit loads (the planes are 16-byte aligned inside the kernel parameters)
    uint32_t kw[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const uint4 v = reinterpret_cast<const uint4 *>(kp)[q];
        kw[4 * q] = v.x; kw[4 * q + 1] = v.y; kw[4 * q + 2] = v.z; kw[4 * q + 3] = v.w;
    }
    if (!last) {
        bs_mix_column<0>(a[0], a[1], a[2], a[3], kw, o);
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int b = 0; b < 8; ++b) o[8 * r + b] = a[r][b] ^ kw[8 * r + b];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) s[j] = o[j];
}

// rounds 3..NR on a state that already went through rounds 0..2
template <int NR>
UAES_HD void bs8_finish(uint32_t s[32], const BsKeyPlanes8 &kp)
{
#if defined(__CUDA_ARCH__) && defined(UAES_BS8_UNROLL)
    // every round its own code: the key words become constant-bank operands of the LOP3s (no loads), at 8 x the code
#pragma unroll
    for (int r = 3; r <= NR; ++r) bs8_round_or_last(s, kp.k[r - 3], r == NR);
#else
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int r = 3; r <= NR; ++r) bs8_round_or_last(s, kp.k[r - 3], r == NR);
#endif
}

// ---- the general form: any 8 blocks (data-dependent modes: XTS, ECB, ...) -------------------------------
// All NR + 1 round keys as packed plane words; the state comes from ONE 32x32 transpose of
// M[8 c + t] = word c of block t and goes back the same way.
struct alignas(16) BsKeyPlanes8Full {
    uint32_t k[kBsMaxRounds + 1][32];       // round keys 0..NR (for decryption: the equivalent-inverse schedule dk[0..NR])
};

// rijndaelEncrypt (micro_aes.c:242-259) on 8 blocks held as packed planes
template <int NR>
UAES_HD void bs8_encrypt_planes(uint32_t s[32], const BsKeyPlanes8Full &kp)
{
#pragma unroll
    for (int j = 0; j < 32; ++j) s[j] ^= kp.k[0][j];
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int r = 1; r <= NR; ++r) bs8_round_or_last(s, kp.k[r], r == NR);
}

// InvShiftRows on a row register: the new column c takes the old column c - R
template <int R> UAES_HD uint32_t bs8_inv_shift_row(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return R == 0 ? x : __byte_perm(x, 0, R == 1 ? 0x2103 : R == 2 ? 0x1032 : 0x0321);
#else
    return R == 0 ? x : (x << (8 * R)) | (x >> (32 - 8 * R));
#endif
}

// one round of the equivalent inverse cipher (see bs_inv_round in uaes_bitslice.cuh), middle or last
UAES_HD void bs8_inv_round_or_last(uint32_t s[32], const uint32_t *kp, bool last)
{
#pragma unroll
    for (int r = 0; r < 4; ++r) sbox_inv_bitsliced(s + 8 * r);
    uint32_t a[4][8], o[32];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        a[0][b] = bs8_inv_shift_row<0>(s[b]);      a[1][b] = bs8_inv_shift_row<1>(s[8 + b]);
        a[2][b] = bs8_inv_shift_row<2>(s[16 + b]); a[3][b] = bs8_inv_shift_row<3>(s[24 + b]);
    }
    if (!last) {
        bs_inv_mix_column(a[0], a[1], a[2], a[3], kp, o);
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int b = 0; b < 8; ++b) o[8 * r + b] = a[r][b] ^ kp[8 * r + b];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) s[j] = o[j];
}

// rijndaelDecrypt (micro_aes.c:315-332) on 8 blocks held as packed planes; kp = planes of dk[0..NR]
template <int NR>
UAES_HD void bs8_decrypt_planes(uint32_t s[32], const BsKeyPlanes8Full &kp)
{
#pragma unroll
    for (int j = 0; j < 32; ++j) s[j] ^= kp.k[0][j];
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int r = 1; r <= NR; ++r) bs8_inv_round_or_last(s, kp.k[r], r == NR);
}

UAES_HD void bs8_make_key_planes_full(const uint32_t *rk, int rounds, BsKeyPlanes8Full *kp)
{
    for (int r = 0; r <= rounds; ++r)
        for (int j = 0; j < 32; ++j)
            kp->k[r][j] = bs8_spread(rk[4 * r], rk[4 * r + 1], rk[4 * r + 2], rk[4 * r + 3], j);
}

// host-side: key planes from the expanded key (the launcher does this per call)
UAES_HD void bs8_make_key_planes(const uint32_t *rk, int rounds, BsKeyPlanes8 *kp)
{
    for (int r = 3; r <= rounds; ++r)
        for (int j = 0; j < 32; ++j)
            kp->k[r - 3][j] = bs8_spread(rk[4 * r], rk[4 * r + 1], rk[4 * r + 2], rk[4 * r + 3], j);
}

}  // namespace uaes
