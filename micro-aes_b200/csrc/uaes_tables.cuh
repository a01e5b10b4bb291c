// uaes_tables.cuh -- AES lookup tables for sm_100a: generated at compile time, expanded
// into bank-conflict-free shared-memory layouts at kernel start.
//
// What the reference does per byte with sbox[] / rsbox[] (micro_aes.c:41-65), xtime
// (micro_aes.c:115-118) and MixColumns / InvMixColumns (micro_aes.c:221-239, 301-312) is
// folded into four 32-bit "T-tables" per direction:
//     Te0[x] = { 2*S(x), S(x), S(x), 3*S(x) }   (byte 0 = row 0), Te_k = Te0 rotated left 8k bits
//     Td0[x] = { 14*Si(x), 9*Si(x), 13*Si(x), 11*Si(x) },        Td_k = Td0 rotated left 8k bits
// so that one round of a column is four lookups and four XORs.
//
// Shared-memory layout (the part that matters on B200): shared memory has 32 banks x 4 B and
// serves one 4-byte word per bank per clock, i.e. at most 32 table lookups / clk / SM.  Random
// S-box indices hit random banks, so every table is stored 32 times, once per lane, with lane l's
// copy living entirely in bank l:
//     address(table t in 0..3, entry x, lane l) = base + t*32768 + x*128 + l*4
// A lookup is then ONE integer dot product (IDP.4A with a one-hot selector computes
// byte*128 + lanebase straight from the state word -- on the FMA pipe, which is otherwise idle,
// leaving the ALU pipe to the XORs and to the bitsliced co-runner of the CTR kernel) plus ONE
// conflict-free LDS with an immediate table offset.  -DUAES_LUT_PRMT selects the first-generation
// layout (x*256 + t*128 + l*4, address by one PRMT on the ALU pipe; same speed stand-alone,
// measured in profiles/r1_idp_vs_prmt.txt).  Kernels over-allocate and align `base` to 64 KiB at
// run time (needed by the PRMT variant; the window starts at a driver-chosen offset).
#pragma once
#include <stdint.h>

namespace uaes {

// ---------------------------------------------------------------- compile-time generation

struct ByteTable { uint8_t v[256]; };
struct WordTable { uint32_t v[256]; };

constexpr uint8_t gf_double(uint8_t a) { return (uint8_t)((a << 1) ^ ((a >> 7) * 0x1b)); }

constexpr uint8_t gf_mul(uint8_t a, uint8_t b)
{
    uint8_t r = 0;
    for (int i = 0; i < 8; ++i) {
        if (b & 1) r ^= a;
        a = gf_double(a);
        b >>= 1;
    }
    return r;
}

// FIPS-197 5.1.1: multiplicative inverse in GF(2^8) followed by the affine map
constexpr ByteTable make_sbox()
{
    ByteTable t{};
    // generator 3 walks the whole multiplicative group: p = 3^i, q = 3^-i
    uint8_t p = 1, q = 1;
    do {
        p = (uint8_t)(p ^ gf_double(p));                       // p *= 3
        q ^= (uint8_t)(q << 1); q ^= (uint8_t)(q << 2); q ^= (uint8_t)(q << 4);
        if (q & 0x80) q ^= 0x09;                               // q /= 3
        uint8_t x = (uint8_t)(q ^ (uint8_t)((q << 1) | (q >> 7)) ^ (uint8_t)((q << 2) | (q >> 6))
                              ^ (uint8_t)((q << 3) | (q >> 5)) ^ (uint8_t)((q << 4) | (q >> 4)));
        t.v[p] = (uint8_t)(x ^ 0x63);
    } while (p != 1);
    t.v[0] = 0x63;
    return t;
}

constexpr ByteTable make_inv_sbox()
{
    ByteTable s = make_sbox(), t{};
    for (int i = 0; i < 256; ++i) t.v[s.v[i]] = (uint8_t)i;
    return t;
}

constexpr WordTable make_te0()
{
    ByteTable s = make_sbox();
    WordTable t{};
    for (int i = 0; i < 256; ++i) {
        const uint32_t x = s.v[i], x2 = gf_double((uint8_t)x), x3 = x2 ^ x;
        t.v[i] = x2 | x << 8 | x << 16 | x3 << 24;
    }
    return t;
}

constexpr WordTable make_td0()
{
    ByteTable s = make_inv_sbox();
    WordTable t{};
    for (int i = 0; i < 256; ++i) {
        const uint8_t x = s.v[i];
        t.v[i] = (uint32_t)gf_mul(x, 14) | (uint32_t)gf_mul(x, 9) << 8
               | (uint32_t)gf_mul(x, 13) << 16 | (uint32_t)gf_mul(x, 11) << 24;
    }
    return t;
}

// inverse S-box replicated into all four bytes: the last decryption round picks bytes from it
constexpr WordTable make_td4()
{
    ByteTable s = make_inv_sbox();
    WordTable t{};
    for (int i = 0; i < 256; ++i) t.v[i] = 0x01010101u * s.v[i];
    return t;
}

// 1 KiB each, read once per CTA (warp-uniform index -> constant-cache broadcast)
__constant__ WordTable c_te0 = make_te0();
__constant__ WordTable c_td0 = make_td0();
__constant__ WordTable c_td4 = make_td4();
// byte tables for the one-off blocks (GCM subkeys, XTS stealing) that run on a single thread
__constant__ ByteTable c_sbox = make_sbox();
__constant__ ByteTable c_inv_sbox = make_inv_sbox();

// ---------------------------------------------------------------- shared-memory layout

constexpr uint32_t kTablePairBytes = 65536;            // 256 entries x (2 tables x 32 lanes x 4 B)
constexpr uint32_t kTableAlign     = 65536;
#ifndef UAES_LUT_PRMT
// IDP.4A addressing: entry stride 128 B, table A in the low 32 KiB of a pair region, B in the high
constexpr uint32_t kOffT0 = 0, kOffT1 = 32768, kOffT2 = kTablePairBytes, kOffT3 = kTablePairBytes + 32768;
#else
constexpr uint32_t kOffT0 = 0, kOffT1 = 128, kOffT2 = kTablePairBytes, kOffT3 = kTablePairBytes + 128;
#endif
constexpr uint32_t kEncTableBytes  = 2 * kTablePairBytes;   // Te0|Te1, Te2|Te3
#ifndef UAES_LUT_PRMT
// decryption keeps all four Td tables, the inverse S-box and (for XTS tweaks) Te0: three pair regions.
// They fit because IDP.4A addressing needs no 64 KiB alignment of the base (uaes_dec_table_base).
constexpr uint32_t kDecTableBytes  = 3 * kTablePairBytes;   // Td0|Td1, Td2|Td3, Td4|Te0
constexpr uint32_t kOffT4 = 2 * kTablePairBytes, kOffT5 = 2 * kTablePairBytes + 32768;
constexpr uint32_t kOffTd4 = kOffT4, kOffDecTe0 = kOffT5;
#else
constexpr uint32_t kDecTableBytes  = 2 * kTablePairBytes;   // Td0|Td1, Td4|Te0 (Td2/Td3 by a 16-bit rotate)
constexpr uint32_t kOffTd4 = kOffT2, kOffDecTe0 = kOffT3;
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int n)   // n in {0,8,16,24}
{
    return n ? __funnelshift_l(x, x, n) : x;
}

// Fill one 64 KiB pair region: table A in the low 128 B of each 256 B row, table B in the high
// 128 B, each word replicated for the 32 lanes.  src is a __constant__ table, rotA / rotB are
// left rotations in bits.  Called by all threads of the CTA.
__device__ __forceinline__ void fill_pair(uint32_t region, const WordTable &srcA, int rotA,
                                          const WordTable &srcB, int rotB)
{
    // 256 rows x 2 tables x 32 replicas = 4096 16-byte stores (four replicas each); consecutive
    // threads write consecutive 16-byte slots (conflict free).  The fill is the fixed cost of every
    // launch (a few microseconds), which is what short calls are made of.
    for (uint32_t q = threadIdx.x; q < 4096u; q += blockDim.x) {
        const uint32_t row = q >> 4, tbl = (q >> 3) & 1u, quad = q & 7u;
        const uint32_t v = tbl ? rotl32(srcB.v[row], rotB) : rotl32(srcA.v[row], rotA);
#ifndef UAES_LUT_PRMT
        const uint32_t ad = region + tbl * 32768u + row * 128u + quad * 16u;
#else
        const uint32_t ad = region + row * 256u + tbl * 128u + quad * 16u;
#endif
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(ad), "r"(v) : "memory");
    }
}

// returns the 64 KiB-aligned base of the encryption tables inside the dynamic smem window
__device__ __forceinline__ uint32_t align_table_base(const void *dyn_smem)
{
    return (smem_u32(dyn_smem) + (kTableAlign - 1)) & ~(kTableAlign - 1);
}

__device__ __forceinline__ void init_enc_tables(uint32_t base)
{
    fill_pair(base, c_te0, 0, c_te0, 8);
    fill_pair(base + kTablePairBytes, c_te0, 16, c_te0, 24);
}

// base of the decryption tables inside the dynamic window
__device__ __forceinline__ uint32_t dec_table_base(const void *dyn_smem)
{
#ifndef UAES_LUT_PRMT
    return (smem_u32(dyn_smem) + 127u) & ~127u;
#else
    return align_table_base(dyn_smem);
#endif
}

// Td0..Td3, the inverse S-box (Td4, last round) and Te0 (XTS decryption encrypts its tweaks).  With
// PRMT addressing only two pair regions fit behind the aligned base: Td2/Td3 are then Td0/Td1 rotated
// by 16 bits after the lookup (extra ALU work, which is why that layout is not the default).
__device__ __forceinline__ void init_dec_tables(uint32_t base)
{
    fill_pair(base, c_td0, 0, c_td0, 8);
#ifndef UAES_LUT_PRMT
    fill_pair(base + kTablePairBytes, c_td0, 16, c_td0, 24);
    fill_pair(base + 2 * kTablePairBytes, c_td4, 0, c_te0, 0);
#else
    fill_pair(base + kTablePairBytes, c_td4, 0, c_te0, 0);
#endif
}

// ---------------------------------------------------------------- the lookup primitive

// lanebase = table base + lane*4 (byte 0 = lane*4 < 128, byte 1 = 0, bytes 2..3 = base >> 16).
// PRMT builds  { lanebase.b0, w.b[BYTE], lanebase.b2, lanebase.b3 }  =  base + x*256 + lane*4.
template <int BYTE, uint32_t OFF>
__device__ __forceinline__ uint32_t lut(uint32_t lanebase, uint32_t w)
{
#ifndef UAES_LUT_PRMT
    // one integer dot product on the FMA pipe: w.b[BYTE] * 128 + lanebase (selector is one-hot)
    const uint32_t a = __dp4a(w, 0x80u << (8 * BYTE), lanebase);
#else
    const uint32_t a = __byte_perm(w, lanebase, 0x7604 | (BYTE << 4));
#endif
    uint32_t r;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(r) : "r"(a), "n"(OFF));
    return r;
}

// lookup by an already extracted index x (0..255) and a run-time table offset: rare, warp-uniform
// set-up work only
__device__ __forceinline__ uint32_t lut_index(uint32_t lanebase, uint32_t off, uint32_t x)
{
    uint32_t r;
#ifndef UAES_LUT_PRMT
    const uint32_t a = lanebase + off + x * 128u;
#else
    const uint32_t a = lanebase + off + x * 256u;
#endif
    asm("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}

}  // namespace uaes
