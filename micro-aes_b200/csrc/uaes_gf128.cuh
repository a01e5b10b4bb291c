// uaes_gf128.cuh -- GF(2^128) arithmetic for XTS tweaks and GHASH on sm_100a.
//
// XTS (micro_aes.c:449-458 doubleLblock): an element is the 16 bytes read as a LITTLE-endian
// 128-bit integer; times-alpha is a left shift with 0x87 folded into bit 0 on carry.  The
// reference walks T_{j+1} = alpha*T_j serially; here every thread jumps straight to
// T_0 * alpha^j from its block index j.
//
// GHASH (micro_aes.c:464-493 divideBblock / mulGF128): an element is the 16 bytes read as a
// BIG-endian bit string, bit 0 (MSB of byte 0) = coefficient of x^0; times-x is a right shift
// with 0xE1 folded into byte 0.  The reference multiplies bit-serially (128 shifts per block);
// here the multiplier of the hot loop is a per-call constant, so the product is a table walk.
#pragma once
#include <stdint.h>

namespace uaes {

// ================================================================= XTS field (little-endian)

struct Tweak { uint64_t lo, hi; };

// T * x^k, 0 <= k <= 56: shift left, fold the k overflow bits with x^128 = x^7 + x^2 + x + 1
__host__ __device__ constexpr Tweak xts_shl(Tweak t, unsigned k)
{
    if (k == 0) return t;
    const uint64_t c = t.hi >> (64 - k);
    Tweak r{};
    r.hi = t.hi << k | t.lo >> (64 - k);
    r.lo = (t.lo << k) ^ c ^ (c << 1) ^ (c << 2) ^ (c << 7);
    return r;
}

// T * x^e for 0 <= e < 128
__host__ __device__ constexpr Tweak xts_mul_xe(Tweak t, unsigned e)
{
    while (e) {
        const unsigned s = e > 56 ? 56 : e;
        t = xts_shl(t, s);
        e -= s;
    }
    return t;
}

// generic product (Horner over the bits of b, MSB first)
__host__ __device__ constexpr Tweak xts_mul(Tweak a, Tweak b)
{
    Tweak r{0, 0};
    for (int i = 127; i >= 0; --i) {
        r = xts_shl(r, 1);
        const uint64_t bit = i >= 64 ? (b.hi >> (i - 64)) & 1 : (b.lo >> i) & 1;
        if (bit) { r.lo ^= a.lo; r.hi ^= a.hi; }
    }
    return r;
}

// Q[i] = x^(128 * 2^i) mod p: the jump-ahead ladder.  x^128 = 0x87, each next entry is the
// square of the previous one.  Evaluated by the compiler.
struct TweakLadder { Tweak q[57]; };
constexpr TweakLadder make_ladder()
{
    TweakLadder l{};
    l.q[0] = Tweak{0x87, 0};
    for (int i = 1; i < 57; ++i) l.q[i] = xts_mul(l.q[i - 1], l.q[i - 1]);
    return l;
}
__constant__ TweakLadder c_ladder = make_ladder();

// T * x^e for any e < 2^63: low 7 bits by shifting, the rest by the ladder (one generic
// product per set bit; runs once per warp at kernel start, never in the block loop)
__device__ inline Tweak xts_jump(Tweak t, uint64_t e)
{
    t = xts_mul_xe(t, (unsigned)(e & 127));
    e >>= 7;
    for (int i = 0; e; ++i, e >>= 1)
        if (e & 1) t = xts_mul(t, c_ladder.q[i]);
    return t;
}

// 32-bit word view used in the block loop (w[0] = bytes 0..3)
__device__ __forceinline__ void tweak_words(Tweak t, uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3)
{
    w0 = (uint32_t)t.lo; w1 = (uint32_t)(t.lo >> 32); w2 = (uint32_t)t.hi; w3 = (uint32_t)(t.hi >> 32);
}

// ================================================================= GHASH field (big-endian bits)

// hi = bytes 0..7 as a big-endian integer (bit 63 = x^0), lo = bytes 8..15 (bit 0 = x^127)
struct Gf { uint64_t hi, lo; };

__host__ __device__ constexpr Gf gf_mulx(Gf v)          // micro_aes.c:464-473
{
    const uint64_t carry = v.lo & 1;
    Gf r{};
    r.lo = v.lo >> 1 | v.hi << 63;
    r.hi = (v.hi >> 1) ^ (carry ? 0xE100000000000000ull : 0);
    return r;
}

__host__ __device__ constexpr Gf gf_mul(Gf x, Gf y)     // micro_aes.c:476-493
{
    Gf z{0, 0};
    for (int i = 0; i < 128; ++i) {
        const uint64_t bit = i < 64 ? (x.hi >> (63 - i)) & 1 : (x.lo >> (127 - i)) & 1;
        if (bit) { z.hi ^= y.hi; z.lo ^= y.lo; }
        y = gf_mulx(y);
    }
    return z;
}

__device__ __forceinline__ uint64_t bswap64(uint64_t v)
{
    const uint32_t a = __byte_perm((uint32_t)v, 0, 0x0123), b = __byte_perm((uint32_t)(v >> 32), 0, 0x0123);
    return (uint64_t)a << 32 | b;
}

// conversions between the Gf view and the four little-endian memory words of a block
__device__ __forceinline__ Gf gf_from_words(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
    Gf g;
    g.hi = bswap64((uint64_t)w1 << 32 | w0);
    g.lo = bswap64((uint64_t)w3 << 32 | w2);
    return g;
}

__device__ __forceinline__ void gf_to_words(Gf g, uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3)
{
    const uint64_t a = bswap64(g.hi), b = bswap64(g.lo);
    w0 = (uint32_t)a; w1 = (uint32_t)(a >> 32); w2 = (uint32_t)b; w3 = (uint32_t)(b >> 32);
}

// R[d] = d(x) * x^128 mod p as the first two bytes of a block (little-endian 16-bit value):
// what falls off the end when a block is multiplied by x^8.  Key independent.
struct GhashReduce { uint16_t v[256]; };
constexpr GhashReduce make_ghash_reduce()
{
    GhashReduce t{};
    for (int d = 0; d < 256; ++d) {
        // d's MSB is the coefficient of x^120 -> x^128 after the byte shift
        Gf acc{0, 0};
        Gf term{0xE100000000000000ull, 0};              // x^128 mod p
        for (int j = 0; j < 8; ++j) {
            if (d & (0x80 >> j)) { acc.hi ^= term.hi; acc.lo ^= term.lo; }
            term = gf_mulx(term);
        }
        // bytes 0 and 1 of the block are the top 16 bits of acc.hi
        const uint32_t b0 = (uint32_t)(acc.hi >> 56) & 0xff, b1 = (uint32_t)(acc.hi >> 48) & 0xff;
        t.v[d] = (uint16_t)(b0 | b1 << 8);
    }
    return t;
}
__constant__ GhashReduce c_ghash_reduce = make_ghash_reduce();

}  // namespace uaes
