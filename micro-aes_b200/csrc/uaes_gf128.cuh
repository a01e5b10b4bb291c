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

// ---- fast generic product for the one-off multiplications (setup, chunk scaling, final fold).
// gf_mul above is a 128-step dependent chain (~8 us on one thread); this one is a carry-less
// 128x128 multiply built from integer multiplies on operands whose bits are spread 4 apart
// ("holes": column sums stay below 16, so no carry crosses into a neighbouring coefficient),
// Karatsuba 128 -> 64 -> 32, then a fold by x^128 = x^7 + x^2 + x + 1.  It works on the natural
// bit order (bit i = coefficient of x^i), i.e. the bit reversal of GHASH's serialisation.
__device__ __forceinline__ uint64_t clmul32(uint32_t x, uint32_t y)
{
    const uint32_t x0 = x & 0x11111111u, x1 = x & 0x22222222u, x2 = x & 0x44444444u, x3 = x & 0x88888888u;
    const uint32_t y0 = y & 0x11111111u, y1 = y & 0x22222222u, y2 = y & 0x44444444u, y3 = y & 0x88888888u;
    auto m = [](uint32_t a, uint32_t b) { return (uint64_t)a * b; };
    const uint64_t z0 = m(x0, y0) ^ m(x1, y3) ^ m(x2, y2) ^ m(x3, y1);
    const uint64_t z1 = m(x0, y1) ^ m(x1, y0) ^ m(x2, y3) ^ m(x3, y2);
    const uint64_t z2 = m(x0, y2) ^ m(x1, y1) ^ m(x2, y0) ^ m(x3, y3);
    const uint64_t z3 = m(x0, y3) ^ m(x1, y2) ^ m(x2, y1) ^ m(x3, y0);
    return (z0 & 0x1111111111111111ull) | (z1 & 0x2222222222222222ull) |
           (z2 & 0x4444444444444444ull) | (z3 & 0x8888888888888888ull);
}

__device__ __forceinline__ void clmul64(uint64_t a, uint64_t b, uint64_t &lo, uint64_t &hi)
{
    const uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    const uint64_t z0 = clmul32(a0, b0), z2 = clmul32(a1, b1);
    const uint64_t z1 = clmul32(a0 ^ a1, b0 ^ b1) ^ z0 ^ z2;
    lo = z0 ^ (z1 << 32);
    hi = z2 ^ (z1 >> 32);
}

__device__ inline Gf gf_mul_fast(Gf x, Gf y)
{
    // natural order: n.lo = coefficients x^0..x^63
    const uint64_t al = __brevll(x.hi), ah = __brevll(x.lo), bl = __brevll(y.hi), bh = __brevll(y.lo);
    uint64_t z0l, z0h, z2l, z2h, z1l, z1h;
    clmul64(al, bl, z0l, z0h);
    clmul64(ah, bh, z2l, z2h);
    clmul64(al ^ ah, bl ^ bh, z1l, z1h);
    z1l ^= z0l ^ z2l; z1h ^= z0h ^ z2h;
    const uint64_t r0 = z0l, r1 = z0h ^ z1l, h0 = z2l ^ z1h, h1 = z2h;
    // fold the upper 128 bits: H * (1 + x + x^2 + x^7), then the <= 7 bits that overflow again
    const uint64_t t0 = h0 ^ (h0 << 1) ^ (h0 << 2) ^ (h0 << 7);
    const uint64_t t1 = h1 ^ (h1 << 1 | h0 >> 63) ^ (h1 << 2 | h0 >> 62) ^ (h1 << 7 | h0 >> 57);
    const uint64_t t2 = (h1 >> 63) ^ (h1 >> 62) ^ (h1 >> 57);
    const uint64_t lo = r0 ^ t0 ^ t2 ^ (t2 << 1) ^ (t2 << 2) ^ (t2 << 7);
    const uint64_t hi = r1 ^ t1;
    Gf r;
    r.hi = __brevll(lo);
    r.lo = __brevll(hi);
    return r;
}

// x^e by square and multiply
__device__ inline Gf gf_pow_fast(Gf x, uint64_t e)
{
    Gf r{0x8000000000000000ull, 0};
    for (; e; e >>= 1) {
        if (e & 1) r = gf_mul_fast(r, x);
        if (e > 1) x = gf_mul_fast(x, x);
    }
    return r;
}

// ---------------------------------------------------------------- GHASH: y <- y * C by table
//
// The multiplier of the GCM hot loop is a per-call constant C, so y * C = sum_i M[y_i] * x^(8i) over
// the 16 bytes of y with M[b] = b(x) * C (bit 7 of b = coefficient of x^0, micro_aes.c:476-493).
// y and the table entries are held as BIG-endian words (W0 = bytes 0..3 of the block, x^0 = bit 31
// of W0), so that "times x^k" is a right shift of the string W0:W1:W2:W3:...; i = 4q + r is byte r
// of word q.  The 16 lookups depend on y alone (no serial chain).  Entries with the same r are
// XORed word-aligned at word offset q into an UNREDUCED 8-word string; the three byte shifts are
// Horner steps over r; the 120 bits beyond x^127 are folded once with x^128 = 1 + x + x^2 + x^7
// (O * x^7 still fits in 128 bits, so one fold is exact).
// fetch(word, k) returns M[byte k of word] (k = register byte, 3 - r).  Host + device: the host
// instance is checked against the oracle's mulGF128 in tests/test_ghash_host.py.
__host__ __device__ __forceinline__ uint32_t shr_pair(uint32_t lo, uint32_t hi, int n)   // low word of (hi:lo) >> n
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, n);
#else
    return lo >> n | hi << (32 - n);
#endif
}

template <class Fetch>
__host__ __device__ __forceinline__ void ghash_mul_table(Fetch fetch, uint32_t &y0, uint32_t &y1, uint32_t &y2, uint32_t &y3)
{
    const uint32_t yw[4] = {y0, y1, y2, y3};
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int r = 3; r >= 0; --r) {
        if (r != 3) {                                        // acc <- acc * x^8
#pragma unroll
            for (int j = 7; j >= 1; --j) acc[j] = shr_pair(acc[j], acc[j - 1], 8);
            acc[0] >>= 8;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 m = fetch(yw[q], 3 - r);
            acc[q] ^= m.x; acc[q + 1] ^= m.y; acc[q + 2] ^= m.z; acc[q + 3] ^= m.w;
        }
    }
    // fold O = acc[4..7] (x^128 .. x^247): y = lo ^ O ^ O*x ^ O*x^2 ^ O*x^7
    const uint32_t o0 = acc[4], o1 = acc[5], o2 = acc[6], o3 = acc[7];
    y0 = acc[0] ^ o0 ^ (o0 >> 1) ^ (o0 >> 2) ^ (o0 >> 7);
    y1 = acc[1] ^ o1 ^ shr_pair(o1, o0, 1) ^ shr_pair(o1, o0, 2) ^ shr_pair(o1, o0, 7);
    y2 = acc[2] ^ o2 ^ shr_pair(o2, o1, 1) ^ shr_pair(o2, o1, 2) ^ shr_pair(o2, o1, 7);
    y3 = acc[3] ^ o3 ^ shr_pair(o3, o2, 1) ^ shr_pair(o3, o2, 2) ^ shr_pair(o3, o2, 7);
}

// M[b] = b(x) * C as big-endian words (the layout ghash_mul_table reads)
__host__ __device__ inline uint4 ghash_table_entry(Gf C, uint32_t b)
{
    Gf acc{0, 0}, t = C;
    for (int j = 0; j < 8; ++j) {
        if (b & (0x80u >> j)) { acc.hi ^= t.hi; acc.lo ^= t.lo; }
        t = gf_mulx(t);
    }
    return make_uint4((uint32_t)(acc.hi >> 32), (uint32_t)acc.hi, (uint32_t)(acc.lo >> 32), (uint32_t)acc.lo);
}

}  // namespace uaes
