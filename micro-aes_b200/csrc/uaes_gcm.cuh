// uaes_gcm.cuh -- GCM kernels (included by uaes_kernels.cu).
//
// Restates AES_GCM_encrypt / gHash / xMac / mulGF128 (micro_aes.c:1164-1179, 1127-1137,
// 551-570, 476-493).  The reference makes two serial passes (CTR, then a bit-serial GHASH chain
// G <- H*(G ^ X_i)).  Here ONE pass reads the plaintext once and writes the ciphertext once:
//
//   GHASH(X_0..X_{n-1}) = sum_i X_i * H^(n-i)  is split into one chunk of CB = 32*R blocks per warp
//   of the persistent grid (R rows, sized so that every warp gets the same share), aligned to the
//   END of the message (a short first chunk is a full one with leading zeros).
//   Inside a chunk lane l of the warp owns blocks = l (mod 32) in counter space and runs
//   Horner with the fixed multiplier C = H^32:   y_l <- y_l * C ^ X.
//   The multiply-by-constant is 16 INDEPENDENT lookups of M[b] = b(x)*C, one per byte of y
//   (256 x 16 B, replicated 8x so the 8 lanes of a quarter-warp hit 8 different bank groups),
//   summed word-aligned into an unreduced 248-bit string and folded once (ghash_mul_const);
//   the block of row r is absorbed while the rounds of row r+1 run.
//   At the end of a chunk lane l scales y_l by H^(distance to the chunk end) in 1..32 (one
//   generic product per chunk) and the warp XOR-reduces by shuffle into one 16-byte partial.
//   A last small kernel folds the partials pairwise with P = H^CB, P^2, P^4, ... (carry-less
//   multiplies built from integer multiplies, gf_mul_fast), absorbs the
//   ragged tail block and the length block and writes  tag = E_K(J0) ^ GHASH.
//
//   The AAD enters as the initial GHASH state, which is simply XORed into block 0.
#pragma once

namespace uaes {

struct GcmWork {
    uint4 H;             // E_K(0), as the four memory words of the block
    uint4 EJ0;           // E_K(J0)
    uint4 C32;           // H^32
    uint4 aad_state;     // GHASH state after the (zero padded) AAD
    uint4 lanepow[32];   // lanepow[e-1] = H^e
    uint4 partials[1];   // NC entries, stored in reverse chunk order
};

__device__ __forceinline__ Gf gf_load(const uint4 &v) { return gf_from_words(v.x, v.y, v.z, v.w); }
__device__ __forceinline__ uint4 gf_store(Gf g)
{
    uint4 v;
    gf_to_words(g, v.x, v.y, v.z, v.w);
    return v;
}

// POLYVAL (GCM-SIV, micro_aes.c:498-528) is GHASH on byte-reversed blocks with the key
// mulX_GHASH(ByteReverse(H)) (RFC 8452 appendix A); ByteReverse of a block held as memory words:
__device__ __forceinline__ uint4 rev_block(uint4 v)
{
    return make_uint4(__byte_perm(v.w, 0, 0x0123), __byte_perm(v.z, 0, 0x0123),
                      __byte_perm(v.y, 0, 0x0123), __byte_perm(v.x, 0, 0x0123));
}

__device__ inline uint4 load_block_bytes(const uint8_t *p, uint32_t n)   // zero padded, n <= 16
{
    uint32_t w[4] = {0, 0, 0, 0};
    for (uint32_t i = 0; i < n; ++i) w[i >> 2] |= (uint32_t)p[i] << (8 * (i & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---------------------------------------------------------------- per-call constants

struct GcmSetupArgs {
    uaes_keysched ks;
    uint32_t j0[4];              // nonce || 00000001 as words (micro_aes.c:1150-1151)
    const uint8_t *aad;
    uint64_t aadlen;
    const uint4 *aad_state_in;   // non-null: GHASH state of the AAD computed by a bulk pass (large AAD)
    int polyval;                 // 1: H comes from auth[] (GCM-SIV message-authentication key)
    uint32_t auth[4];
    GcmWork *work;
};

// one warp: H and E_K(J0) on two lanes, the squarings H^2..H^32 on lane 0, then lane l
// assembles H^(l+1) from its binary digits; lane 0 finally absorbs the AAD (xMac over the AAD,
// micro_aes.c:1134)
__global__ void gcm_setup_kernel(const __grid_constant__ GcmSetupArgs a)
{
    __shared__ Gf sq[6];                                      // H^(2^i)
    const uint32_t lane = threadIdx.x;
    if (lane < 2) {
        uint32_t s[4] = {0, 0, 0, 0};
        if (lane == 1) { s[0] = a.j0[0]; s[1] = a.j0[1]; s[2] = a.j0[2]; s[3] = a.j0[3]; }
        small_encrypt(a.ks.w, a.ks.rounds, s);
        uint4 v = make_uint4(s[0], s[1], s[2], s[3]);
        if (lane == 0 && a.polyval)                           // H' = mulX_GHASH(ByteReverse(auth key))
            v = gf_store(gf_mulx(gf_load(rev_block(make_uint4(a.auth[0], a.auth[1], a.auth[2], a.auth[3])))));
        if (lane == 0) { a.work->H = v; sq[0] = gf_load(v); } else a.work->EJ0 = v;
    }
    __syncwarp();
    if (lane == 0)
        for (int i = 1; i < 6; ++i) sq[i] = gf_mul_fast(sq[i - 1], sq[i - 1]);
    __syncwarp();
    {
        const uint32_t e = lane + 1;                         // H^e
        Gf p{0x8000000000000000ull, 0};                      // the field's 1
        for (int i = 0; i < 6; ++i)
            if (e >> i & 1) p = gf_mul_fast(p, sq[i]);
        a.work->lanepow[lane] = gf_store(p);
        if (e == 32) a.work->C32 = gf_store(p);
    }
    if (lane == 0 && a.aad_state_in) {
        a.work->aad_state = *a.aad_state_in;
    } else if (lane == 0) {
        Gf g{0, 0};
        const Gf H = sq[0];
        for (uint64_t off = 0; off < a.aadlen; off += 16) {
            const uint64_t left = a.aadlen - off;
            uint4 blk = load_block_bytes(a.aad + off, left < 16 ? (uint32_t)left : 16);
            if (a.polyval) blk = rev_block(blk);
            const Gf x = gf_load(blk);
            g.hi ^= x.hi; g.lo ^= x.lo;
            g = gf_mul_fast(H, g);
        }
        a.work->aad_state = gf_store(g);
    }
}

// ---------------------------------------------------------------- the fused bulk pass

struct GcmBulkArgs {
    uaes_keysched ks;
    uint32_t w0, w1, b8;
    uint64_t v0;                 // counter of data block 0 = J0 + 1
    const uint4 *in;
    uint4 *out;
    uint64_t nblocks;            // FULL blocks only; the ragged tail is done by the finish kernel
    uint64_t chunk_blocks;       // CB = 32 * rows per chunk
    uint64_t nchunks;
    GcmWork *work;
};

constexpr uint32_t kGhashRegion = 32768;                 // M table: 256 x 16 B, replicated 8x
#ifndef UAES_GCM_THREADS
#define UAES_GCM_THREADS 640                             // 96 registers per thread, no spills (768 spills
#endif                                                   // and is slower); the lookup pipe saturates from 16 warps up
constexpr int kGcmThreads = UAES_GCM_THREADS;
constexpr int kGcmWarps = kGcmThreads / 32;

// y <- y * C on the device: ghash_mul_table (uaes_gf128.cuh) with the entry fetched from the 8x
// replicated shared-memory table.  mb = M base | (lane&7)*16; the index byte*128 + mb is one IDP.4A
// on the FMA pipe (one-hot selector), the fetch one LDS.128.
__device__ __forceinline__ void ghash_mul_const(uint32_t mb, uint32_t &y0, uint32_t &y1, uint32_t &y2, uint32_t &y3)
{
    ghash_mul_table([mb](uint32_t word, int byte) -> uint4 {
        const uint32_t ad = __dp4a(word, 0x80u << (8 * byte), mb);
        uint4 m;
        asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(m.x), "=r"(m.y), "=r"(m.z), "=r"(m.w) : "r"(ad));
        return m;
    }, y0, y1, y2, y3);
}

// The table-driven warps' part of the pass: chunks first_chunk, first_chunk + nwarps, ... of the range
// [0, region_blocks) (the whole message, or what is left in front of the bitsliced warps' region).
template <int NR, int MODE, bool REV>
__device__ __forceinline__ void gcm_tt_chunks(const GcmBulkArgs &a, uint64_t region_blocks, uint32_t lb, uint32_t mb,
                                              uint64_t first_chunk, uint64_t nwarps)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t *rk = a.ks.w;
    const uint64_t CB = a.chunk_blocks;
    uint4 aad_be = a.work->aad_state;                    // the GHASH state lives in big-endian words
    aad_be = make_uint4(__byte_perm(aad_be.x, 0, 0x0123), __byte_perm(aad_be.y, 0, 0x0123),
                        __byte_perm(aad_be.z, 0, 0x0123), __byte_perm(aad_be.w, 0, 0x0123));

    uint64_t cur_group = ~0ull;
    uint32_t s3 = 0, K0 = 0, D0 = 0, D1 = 0, D2 = 0, D3 = 0;

    for (uint64_t c = first_chunk; c < a.nchunks; c += nwarps) {
        // chunk c ends (NC-1-c) chunks before the end of the message
        const uint64_t b1 = region_blocks - (a.nchunks - 1 - c) * CB;
        const uint64_t b0 = b1 > CB ? b1 - CB : 0;
        const uint64_t vfirst = (a.v0 + b0) & ~31ull, vend = a.v0 + b1;      // counter space rows
        uint32_t y0 = 0, y1 = 0, y2 = 0, y3 = 0;
        uint64_t klast = 0;
        bool any = false;
        uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;             // previous row's block, not yet absorbed
        bool pend = false;

        auto valid_k = [&](uint64_t vrow, uint64_t &k) -> bool {
            const uint64_t v = vrow + lane;
            k = v - a.v0;
            return v >= a.v0 + b0 && v < vend;
        };
        uint64_t kk;
        uint4 cur = valid_k(vfirst, kk) ? ld_stream(a.in + kk) : make_uint4(0, 0, 0, 0);
        for (uint64_t vrow = vfirst; vrow < vend; vrow += 32) {
            uint64_t k, kn;
            const bool ok = valid_k(vrow, k);
            const uint4 nxt = (vrow + 32 < vend && valid_k(vrow + 32, kn)) ? ld_stream(a.in + kn) : make_uint4(0, 0, 0, 0);
            uint32_t o0 = cur.x, o1 = cur.y, o2 = cur.z, o3 = cur.w;
            // absorb the PREVIOUS row's block while this row's rounds run: the two are independent, so
            // their lookups interleave in one basic block (0 * C = 0: no branch on the first row)
            uint32_t m0 = y0, m1 = y1, m2 = y2, m3 = y3;
            ghash_mul_const(mb, m0, m1, m2, m3);
            y0 = pend ? m0 ^ p0 : y0; y1 = pend ? m1 ^ p1 : y1; y2 = pend ? m2 ^ p2 : y2; y3 = pend ? m3 ^ p3 : y3;
            if (MODE != 1) {
                if ((vrow >> 8) != cur_group) {              // same hoisting as ctr_kernel
                    cur_group = vrow >> 8;
                    uint32_t w2, w3;
                    ctr_words(a.b8, (cur_group << 8) & kMask56, w2, w3);
                    const uint32_t s0 = a.w0 ^ rk[0], s1 = a.w1 ^ rk[1], s2 = w2 ^ rk[2];
                    s3 = w3 ^ rk[3];
                    K0 = lut<0, kOffT0>(lb, s0) ^ lut<1, kOffT1>(lb, s1) ^ lut<2, kOffT2>(lb, s2) ^ rk[4];
                    const uint32_t C1 = lut<0, kOffT0>(lb, s1) ^ lut<1, kOffT1>(lb, s2) ^ lut<2, kOffT2>(lb, s3) ^ lut<3, kOffT3>(lb, s0) ^ rk[5];
                    const uint32_t C2 = lut<0, kOffT0>(lb, s2) ^ lut<1, kOffT1>(lb, s3) ^ lut<2, kOffT2>(lb, s0) ^ lut<3, kOffT3>(lb, s1) ^ rk[6];
                    const uint32_t C3 = lut<0, kOffT0>(lb, s3) ^ lut<1, kOffT1>(lb, s0) ^ lut<2, kOffT2>(lb, s1) ^ lut<3, kOffT3>(lb, s2) ^ rk[7];
                    D0 = lut<1, kOffT1>(lb, C1) ^ lut<2, kOffT2>(lb, C2) ^ lut<3, kOffT3>(lb, C3) ^ rk[8];
                    D1 = lut<0, kOffT0>(lb, C1) ^ lut<1, kOffT1>(lb, C2) ^ lut<2, kOffT2>(lb, C3) ^ rk[9];
                    D2 = lut<0, kOffT0>(lb, C2) ^ lut<1, kOffT1>(lb, C3) ^ lut<3, kOffT3>(lb, C1) ^ rk[10];
                    D3 = lut<0, kOffT0>(lb, C3) ^ lut<2, kOffT2>(lb, C1) ^ lut<3, kOffT3>(lb, C2) ^ rk[11];
                }
                const uint32_t c0 = K0 ^ lut<3, kOffT3>(lb, s3 ^ ((((uint32_t)vrow & 255u) + lane) << 24));
                uint32_t t0 = D0 ^ lut<0, kOffT0>(lb, c0), t1 = D1 ^ lut<3, kOffT3>(lb, c0);
                uint32_t t2 = D2 ^ lut<2, kOffT2>(lb, c0), t3 = D3 ^ lut<1, kOffT1>(lb, c0);
                enc_finish<NR, 3>(lb, t0, t1, t2, t3, rk, o0, o1, o2, o3);
                if (ok) st_stream(a.out + k, make_uint4(t0, t1, t2, t3));
                if (MODE == 0) { o0 = t0; o1 = t1; o2 = t2; o3 = t3; }     // GHASH runs over ciphertext
            }
            if (ok) {
                // to big-endian words; for POLYVAL the byte reversal of the block (rev_block) and the
                // word swap cancel into a plain reversal of the word order
                uint32_t e0, e1, e2, e3;
                if (REV) { e0 = o3; e1 = o2; e2 = o1; e3 = o0; }
                else {
                    e0 = __byte_perm(o0, 0, 0x0123); e1 = __byte_perm(o1, 0, 0x0123);
                    e2 = __byte_perm(o2, 0, 0x0123); e3 = __byte_perm(o3, 0, 0x0123);
                }
                if (k == 0) { e0 ^= aad_be.x; e1 ^= aad_be.y; e2 ^= aad_be.z; e3 ^= aad_be.w; }
                p0 = e0; p1 = e1; p2 = e2; p3 = e3;
                any = true;
                klast = k;
            }
            pend = ok;
            cur = nxt;
        }
        if (pend) { ghash_mul_const(mb, y0, y1, y2, y3); y0 ^= p0; y1 ^= p1; y2 ^= p2; y3 ^= p3; }

        // lane l holds sum_j X_(l+32j) * C^(J-j); scale by H^(b1 - klast) and reduce over the warp
        Gf z{0, 0};
        if (any) z = gf_mul_fast(gf_load(a.work->lanepow[(uint32_t)(b1 - klast) - 1]),
                                 Gf{(uint64_t)y0 << 32 | y1, (uint64_t)y2 << 32 | y3});
        for (int o = 16; o; o >>= 1) {
            z.hi ^= __shfl_xor_sync(0xffffffffu, z.hi, o);
            z.lo ^= __shfl_xor_sync(0xffffffffu, z.lo, o);
        }
        if (lane == 0) a.work->partials[a.nchunks - 1 - c] = gf_store(z);
    }
}


// MODE 0: encrypt (CTR, hash the OUTPUT)   MODE 1: hash only (GCM decrypt's verify pass)
// MODE 2: decrypt shard (CTR, hash the INPUT in the same pass; multi-GPU shards, where the caller
//         compares the combined tag afterwards)
// REV: absorb byte-reversed blocks (POLYVAL)
template <int NR, int MODE, bool REV = false>
__global__ void __launch_bounds__(kGcmThreads, 1) gcm_bulk_kernel(const __grid_constant__ GcmBulkArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    // ---- shared memory map: AES tables 64 KiB aligned, the GHASH table in a 32 KiB-aligned gap
    const uint32_t win0 = smem_u32(dyn), win1 = win0 + dyn_smem_size();
    const uint32_t tbase = align_table_base(dyn);
    const uint32_t mbase = tbase >= win0 + kGhashRegion ? tbase - kGhashRegion : tbase + kEncTableBytes;
    if (mbase + kGhashRegion > win1 || tbase + kEncTableBytes > win1) __trap();

    if (MODE != 1) init_enc_tables(tbase);
    {
        // M[b] = b(x) * C: bit 7 of b is the coefficient of x^0 (micro_aes.c:476-493 bit order)
        const Gf C = gf_load(a.work->C32);
        if (threadIdx.x < 256) {
            const uint4 v = ghash_table_entry(C, threadIdx.x);
            for (int rep = 0; rep < 8; ++rep) {
                const uint32_t ad = mbase + threadIdx.x * 128 + rep * 16;
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ad), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
        }
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    uint32_t lb = tbase + lane * 4, mb = mbase | ((lane & 7) << 4);
    asm volatile("" : "+r"(lb), "+r"(mb)::"memory");

    gcm_tt_chunks<NR, MODE, REV>(a, a.nblocks, lb, mb, (uint64_t)blockIdx.x * kGcmWarps + (threadIdx.x >> 5),
                                 (uint64_t)gridDim.x * kGcmWarps);
}

// ---- GCM with the co-runner (AES-GCM encryption / decrypting shards of >= 128 MiB) -----------------
// The bulk kernel above sits at the shared-memory lookup roof with 192 wavefronts per block (128 AES +
// 64 GHASH) and half of the ALU pipe idle.  Here one warpgroup of bitsliced warps (uaes_bitslice.cuh,
// the CTR-specialised form of ctr_kernel) produces the keystream of the LAST part of the message on
// the ALU pipe; those blocks cost the lookup pipe only their 64 GHASH wavefronts.  After the 32 x 32
// transposes lane l of a bitsliced warp holds the blocks = l (mod 32) of 1024 consecutive counters --
// exactly the lane ownership of the GHASH Horner with multiplier H^32 -- so the warp XORs, stores and
// absorbs its 32 slots in order with ghash_mul_const.  A bitsliced warp owns a contiguous run of
// passes (= one GHASH chunk); it scales its partial to the END of the message itself (one power of H
// per launch), so the finish kernel only XORs those partials onto the folded table-driven region.
// 3 table-driven warpgroups at 96 registers + 1 bitsliced warpgroup at 224 (setmaxnreg works on whole
// warpgroups only: a 2-warp co-runner next to 16 table-driven warps hangs, profiles/r2_setmaxnreg_partial.txt).
struct GcmHybridArgs {
    GcmBulkArgs b;               // b.nblocks = all full blocks; b.chunk_blocks / b.nchunks plan [0, a_blocks)
    uint64_t a_blocks;           // blocks [0, a_blocks): table-driven warps; v0 + a_blocks is a multiple of 1024
    uint64_t bs_passes;          // 1024-counter passes covering [a_blocks, nblocks)
    uint64_t bs_per;             // passes per bitsliced warp
    uint64_t nchunks_b;          // bitsliced chunks: partials[b.nchunks + j], already scaled to the end of the range
    BsKeyPlanes bs;
};

#ifndef UAES_GCM_BS_AHEAD
#define UAES_GCM_BS_AHEAD 1                       // rows a bitsliced warp loads ahead of its XOR / store / absorb loop
#endif
#ifndef UAES_GCM_TT
#define UAES_GCM_TT 384                           // table-driven threads of the co-runner kernels
#endif
#ifndef UAES_GCM_TT_REGS
#define UAES_GCM_TT_REGS 96
#endif
constexpr int kGcmHybTtThreads = UAES_GCM_TT;
constexpr int kGcmHybTtWarps = kGcmHybTtThreads / 32;

template <int NR, int MODE>
__device__ __forceinline__ void gcm_bs_chunk(const GcmHybridArgs &h, uint32_t lb, uint32_t mb, uint32_t *um, uint64_t j)
{
    const GcmBulkArgs &a = h.b;
    const uint32_t *rk = a.ks.w;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t p0 = j * h.bs_per, p1 = p0 + h.bs_per < h.bs_passes ? p0 + h.bs_per : h.bs_passes;
    const uint64_t ubase = a.v0 + h.a_blocks;                   // counter of the region's first block (not reduced mod 2^56)
    uint64_t tag16 = ~0ull, klast = 0;
    uint32_t y0 = 0, y1 = 0, y2 = 0, y3 = 0;
    bool any = false;

    for (uint64_t p = p0; p < p1; ++p) {
        const uint64_t u0 = ubase + (p << 10);
        const uint64_t vc = u0 & kMask56;
        const uint64_t kb = h.a_blocks + (p << 10);              // block index of (slot 0, lane 0)
#pragma unroll
        for (int i = 0; i < 4; ++i) {                            // this pass's 16 KiB towards L2 while the rounds run
            const uint64_t k = kb + (uint64_t)(lane * 4 + i) * 8;
            if (k < a.nblocks) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.in + k));
        }
        if ((vc >> 16) != tag16) {                               // counter bytes <= 13 changed: every 64 passes
            tag16 = vc >> 16;
            uint32_t w2, w3, uw[6];
            ctr_words(a.b8, vc, w2, w3);
            bs_uniform_words([&](int t, uint32_t x) { return lut_index(lb, t == 0 ? kOffT0 : t == 1 ? kOffT1 : t == 2 ? kOffT2 : kOffT3, x); },
                             a.w0 ^ rk[0], a.w1 ^ rk[1], w2 ^ rk[2], w3 ^ rk[3], rk, uw);
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 6; ++q) um[32 * q + lane] = bs_mask(uw[q], (int)lane);
            __syncwarp();
        }
        uint32_t s[128];
        bs_first_rounds(s, lane, (uint32_t)(vc >> 8) & 0xfcu, h.bs.k0, um);
#pragma unroll 1
        for (int r = 3; r <= NR; ++r) bs_round_or_last(s, h.bs.k[r - 3], r == NR);
        const uint64_t k0 = kb + lane;
        uint4 x = k0 < a.nblocks ? ld_stream(a.in + k0) : make_uint4(0, 0, 0, 0);
#if UAES_GCM_BS_AHEAD == 2
        uint4 x1 = k0 + 32 < a.nblocks ? ld_stream(a.in + k0 + 32) : make_uint4(0, 0, 0, 0);
#endif
#pragma unroll
        for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
#pragma unroll
        for (int t = 0; t < 32; ++t) {
            const uint64_t k = k0 + 32 * t;
#if UAES_GCM_BS_AHEAD == 2
            const uint4 nx = x1;
            x1 = (t + 2 < 32 && k + 64 < a.nblocks) ? ld_stream(a.in + k + 64) : make_uint4(0, 0, 0, 0);
#else
            const uint4 nx = (t + 1 < 32 && k + 32 < a.nblocks) ? ld_stream(a.in + k + 32) : make_uint4(0, 0, 0, 0);
#endif
            if (k < a.nblocks) {
                const uint4 o = make_uint4(x.x ^ s[t], x.y ^ s[32 + t], x.z ^ s[64 + t], x.w ^ s[96 + t]);
                st_stream(a.out + k, o);
                const uint4 g = MODE == 0 ? o : x;               // GHASH runs over the ciphertext
                ghash_mul_const(mb, y0, y1, y2, y3);
                y0 ^= __byte_perm(g.x, 0, 0x0123); y1 ^= __byte_perm(g.y, 0, 0x0123);
                y2 ^= __byte_perm(g.z, 0, 0x0123); y3 ^= __byte_perm(g.w, 0, 0x0123);
                klast = k; any = true;
            }
            x = nx;
        }
    }
    // chunk end b1; lane l holds sum_j X_(l+32j) * C^(J-j): scale by H^(b1 - klast), reduce over the warp,
    // then move the partial to the end of the range: * H^(nblocks - b1)
    const uint64_t b1 = h.a_blocks + (p1 << 10) < a.nblocks ? h.a_blocks + (p1 << 10) : a.nblocks;
    Gf z{0, 0};
    if (any) z = gf_mul_fast(gf_load(a.work->lanepow[(uint32_t)(b1 - klast) - 1]),
                             Gf{(uint64_t)y0 << 32 | y1, (uint64_t)y2 << 32 | y3});
    for (int o = 16; o; o >>= 1) {
        z.hi ^= __shfl_xor_sync(0xffffffffu, z.hi, o);
        z.lo ^= __shfl_xor_sync(0xffffffffu, z.lo, o);
    }
    if (lane == 0) {
        if (a.nblocks > b1) z = gf_mul_fast(z, gf_pow_fast(gf_load(a.work->H), a.nblocks - b1));
        a.work->partials[a.nchunks + j] = gf_store(z);
    }
}

template <int NR, int MODE>
__global__ void __launch_bounds__(kGcmHybTtThreads + kBsThreads, 1) gcm_bulk_hybrid_kernel(const __grid_constant__ GcmHybridArgs h)
{
    static_assert(MODE == 0 || MODE == 2, "the co-runner produces keystream: encrypt or decrypting shard");
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t win0 = smem_u32(dyn), win1 = win0 + dyn_smem_size();
    const uint32_t tbase = align_table_base(dyn);
    const uint32_t mbase = tbase >= win0 + kGhashRegion ? tbase - kGhashRegion : tbase + kEncTableBytes;
    const uint32_t after = mbase > tbase ? mbase + kGhashRegion : tbase + kEncTableBytes;   // uniform masks of the bitsliced warps
    if (mbase + kGhashRegion > win1 || tbase + kEncTableBytes > win1 || after + 4 * kBsUniformMasks * 4 > win1) __trap();

    init_enc_tables(tbase);
    {
        const Gf C = gf_load(h.b.work->C32);
        if (threadIdx.x < 256) {
            const uint4 v = ghash_table_entry(C, threadIdx.x);
            for (int rep = 0; rep < 8; ++rep) {
                const uint32_t ad = mbase + threadIdx.x * 128 + rep * 16;
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ad), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
        }
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    uint32_t lb = tbase + lane * 4, mb = mbase | ((lane & 7) << 4);
    asm volatile("" : "+r"(lb), "+r"(mb)::"memory");

    constexpr int kLaunchRegs = (65536 / (kGcmHybTtThreads + kBsThreads)) / 8 * 8;     // 128
    constexpr int kTtRegs = UAES_GCM_TT_REGS, kBsRegs = (kLaunchRegs + (kLaunchRegs - kTtRegs) * kGcmHybTtThreads / kBsThreads) / 8 * 8;   // 224
    if (threadIdx.x >= kGcmHybTtThreads) {
        reg_inc<kBsRegs>();
        const uint32_t bw = (threadIdx.x - kGcmHybTtThreads) >> 5;
        uint32_t *um = (uint32_t *)(dyn + (after - win0)) + bw * kBsUniformMasks;
        const uint64_t nw = (uint64_t)gridDim.x * (kBsThreads / 32);
        for (uint64_t j = (uint64_t)blockIdx.x * (kBsThreads / 32) + bw; j < h.nchunks_b; j += nw)
            gcm_bs_chunk<NR, MODE>(h, lb, mb, um, j);
        return;
    }
    reg_dec<kTtRegs>();
    gcm_tt_chunks<NR, MODE, false>(h.b, h.a_blocks, lb, mb, (uint64_t)blockIdx.x * kGcmHybTtWarps + (threadIdx.x >> 5),
                                   (uint64_t)gridDim.x * kGcmHybTtWarps);
}

// ---- the same with NARROW bitsliced warps (uaes_bitslice8.cuh; ctr_queue8_kernel in uaes_kernels.cu) ------
// 8 blocks per thread: a pass is one group of 256 counters, lane l holds the blocks = l (mod 32) of it, so
// the GHASH Horner with H^32 runs 8 steps per pass instead of 32 (the absorb code is a quarter of the
// wide form's 180 KB), the warps need 96 registers like the table-driven ones, and 8 of them fit per SM.
// The chunk bookkeeping (runs of 1024-counter passes, partial scaled to the end of the range) is unchanged.
struct GcmHybridArgs8 {
    GcmBulkArgs b;
    uint64_t a_blocks, bs_passes, bs_per, nchunks_b;     // as GcmHybridArgs
    BsKeyPlanes8 bs8;
};

#ifndef UAES_GCM8_BS
#define UAES_GCM8_BS 256
#endif
#ifndef UAES_GCM8_ROLL
#define UAES_GCM8_ROLL 1
#endif
constexpr uint32_t kGcm8StageWords = UAES_GCM8_ROLL ? 16 * 32 : 0;   // per bitsliced warp: four slots of keystream words, one column per lane
constexpr int kGcm8BsThreads = UAES_GCM8_BS;

template <int NR, int MODE>
__device__ __forceinline__ void gcm_bs8_chunk(const GcmHybridArgs8 &h, uint32_t lb, uint32_t mb, uint32_t *ws, uint32_t *stg, uint64_t j)
{
    const GcmBulkArgs &a = h.b;
    const uint32_t *rk = a.ks.w;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t *up = ws + 4 * lane, *dm = ws + 32 * 32;
    const uint64_t p0 = j * h.bs_per, p1 = p0 + h.bs_per < h.bs_passes ? p0 + h.bs_per : h.bs_passes;
    const uint64_t ubase = a.v0 + h.a_blocks;                   // counter of the region's first block (not reduced mod 2^56)
    uint64_t klast = 0;
    uint32_t y0 = 0, y1 = 0, y2 = 0, y3 = 0;
    bool any = false;
    Bs8Hoist hoist;

#pragma unroll 1
    for (uint64_t g = p0 << 2; g < p1 << 2; ++g) {               // groups of 256 counters
        const uint64_t kb = h.a_blocks + (g << 8);               // block index of (slot 0, lane 0)
        if (kb >= a.nblocks) break;
        {
            const uint64_t k = kb + 8 * (uint64_t)lane;          // this pass's 4 KiB towards L2 while the rounds run
            if (k < a.nblocks) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.in + k));
        }
        uint32_t s[32];
        hoist.group(s, lb, rk, a.w0, a.w1, a.b8, (ubase + (g << 8)) & kMask56, up, dm, lane);
        bs8_finish<NR>(s, h.bs8);
        const uint64_t k0 = kb + lane;
        uint4 x = k0 < a.nblocks ? ld_stream(a.in + k0) : make_uint4(0, 0, 0, 0);
        bs_transpose32(s);                                       // s[8 c + t] = word c of slot t
#if UAES_GCM8_ROLL
        // the absorb loop ROLLED: the keystream words go through a lane-private column of shared memory, four
        // slots at a time, so that the GHASH step exists twice in the code instead of eight times
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
            for (int tt = 0; tt < 4; ++tt)
#pragma unroll
                for (int c = 0; c < 4; ++c) stg[(4 * tt + c) * 32] = s[8 * c + 4 * hf + tt];
#pragma unroll 1
            for (int tt = 0; tt < 4; ++tt) {
                const uint64_t k = k0 + 32 * (4 * hf + tt);
                const uint4 nx = (4 * hf + tt + 1 < 8 && k + 32 < a.nblocks) ? ld_stream(a.in + k + 32) : make_uint4(0, 0, 0, 0);
                if (k < a.nblocks) {
                    const uint32_t *ksw = stg + 128 * tt;
                    const uint4 o = make_uint4(x.x ^ ksw[0], x.y ^ ksw[32], x.z ^ ksw[64], x.w ^ ksw[96]);
                    st_stream(a.out + k, o);
                    const uint4 gh = MODE == 0 ? o : x;          // GHASH runs over the ciphertext
                    ghash_mul_const(mb, y0, y1, y2, y3);
                    y0 ^= __byte_perm(gh.x, 0, 0x0123); y1 ^= __byte_perm(gh.y, 0, 0x0123);
                    y2 ^= __byte_perm(gh.z, 0, 0x0123); y3 ^= __byte_perm(gh.w, 0, 0x0123);
                    klast = k; any = true;
                }
                x = nx;
            }
        }
#else
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const uint64_t k = k0 + 32 * t;
            const uint4 nx = (t + 1 < 8 && k + 32 < a.nblocks) ? ld_stream(a.in + k + 32) : make_uint4(0, 0, 0, 0);
            if (k < a.nblocks) {
                const uint4 o = make_uint4(x.x ^ s[t], x.y ^ s[8 + t], x.z ^ s[16 + t], x.w ^ s[24 + t]);
                st_stream(a.out + k, o);
                const uint4 gh = MODE == 0 ? o : x;              // GHASH runs over the ciphertext
                ghash_mul_const(mb, y0, y1, y2, y3);
                y0 ^= __byte_perm(gh.x, 0, 0x0123); y1 ^= __byte_perm(gh.y, 0, 0x0123);
                y2 ^= __byte_perm(gh.z, 0, 0x0123); y3 ^= __byte_perm(gh.w, 0, 0x0123);
                klast = k; any = true;
            }
            x = nx;
        }
#endif
    }
    // chunk end b1; lane l holds sum_j X_(l+32j) * C^(J-j): scale by H^(b1 - klast), reduce over the warp,
    // then move the partial to the end of the range: * H^(nblocks - b1)
    const uint64_t b1 = h.a_blocks + (p1 << 10) < a.nblocks ? h.a_blocks + (p1 << 10) : a.nblocks;
    Gf z{0, 0};
    if (any) z = gf_mul_fast(gf_load(a.work->lanepow[(uint32_t)(b1 - klast) - 1]),
                             Gf{(uint64_t)y0 << 32 | y1, (uint64_t)y2 << 32 | y3});
    for (int o = 16; o; o >>= 1) {
        z.hi ^= __shfl_xor_sync(0xffffffffu, z.hi, o);
        z.lo ^= __shfl_xor_sync(0xffffffffu, z.lo, o);
    }
    if (lane == 0) {
        if (a.nblocks > b1) z = gf_mul_fast(z, gf_pow_fast(gf_load(a.work->H), a.nblocks - b1));
        a.work->partials[a.nchunks + j] = gf_store(z);
    }
}

template <int NR, int MODE>
__global__ void __launch_bounds__(kGcmHybTtThreads + kGcm8BsThreads, 1) gcm_bulk_hybrid8_kernel(const __grid_constant__ GcmHybridArgs8 h)
{
    static_assert(MODE == 0 || MODE == 2, "the co-runner produces keystream: encrypt or decrypting shard");
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t win0 = smem_u32(dyn), win1 = win0 + dyn_smem_size();
    const uint32_t tbase = align_table_base(dyn);
    const uint32_t mbase = tbase >= win0 + kGhashRegion ? tbase - kGhashRegion : tbase + kEncTableBytes;
    const uint32_t after = mbase > tbase ? mbase + kGhashRegion : tbase + kEncTableBytes;   // the bitsliced warps' U planes and D masks
    constexpr uint32_t kBsWarps = kGcm8BsThreads / 32;
    if (mbase + kGhashRegion > win1 || tbase + kEncTableBytes > win1 || after + kBsWarps * kBs8WarpWords * 4 > win1) __trap();
    // staging columns of the rolled absorb loop: in front of the GHASH table when that lies in the gap, else behind the warps' areas
    const uint32_t sbase = (win0 + 15u) & ~15u;
    const bool stage_front = mbase < tbase && sbase + kBsWarps * kGcm8StageWords * 4 <= mbase;
    const uint32_t stage0 = stage_front ? sbase : after + kBsWarps * kBs8WarpWords * 4;
    if (stage0 + kBsWarps * kGcm8StageWords * 4 > (stage_front ? mbase : win1)) __trap();

    init_enc_tables(tbase);
    {
        const Gf C = gf_load(h.b.work->C32);
        if (threadIdx.x < 256) {
            const uint4 v = ghash_table_entry(C, threadIdx.x);
            for (int rep = 0; rep < 8; ++rep) {
                const uint32_t ad = mbase + threadIdx.x * 128 + rep * 16;
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ad), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
        }
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t lb = tbase + lane * 4, mb = mbase | ((lane & 7) << 4);
    asm volatile("" : "+r"(lb), "+r"(mb)::"memory");

    if (w < kBsWarps) {                          // the bitsliced warps come first (ctr_queue8_kernel: UAES_Q8_MAP)
        uint32_t *ws = (uint32_t *)(dyn + (after - win0)) + w * kBs8WarpWords;
        const uint64_t nw = (uint64_t)gridDim.x * kBsWarps;
        uint32_t *stg = (uint32_t *)(dyn + (stage0 - win0)) + w * kGcm8StageWords + lane;
        for (uint64_t j = (uint64_t)blockIdx.x * kBsWarps + w; j < h.nchunks_b; j += nw)
            gcm_bs8_chunk<NR, MODE>(h, lb, mb, ws, stg, j);
        return;
    }
    gcm_tt_chunks<NR, MODE, false>(h.b, h.a_blocks, lb, mb, (uint64_t)blockIdx.x * kGcmHybTtWarps + (w - kBsWarps),
                                   (uint64_t)gridDim.x * kGcmHybTtWarps);
}

// ---------------------------------------------------------------- fold, tail, tag

struct GcmFinishArgs {
    uaes_keysched ks;
    uint32_t w0, w1, b8;
    uint64_t v0;
    const uint8_t *in;
    uint8_t *out;
    uint64_t len, aadlen;
    uint64_t nparts, chunk_rows; // partials left by the table-driven warps; neighbours are H^(32*rows) apart
    uint64_t nparts_b, blocks_b; // co-runner: partials [nparts, nparts + nparts_b), already scaled to the end of the
                                 // full-block range, which lies blocks_b blocks behind the table-driven region
    int mode;                    // as gcm_bulk_kernel's MODE
    int partial_only;            // write the GHASH state of this shard instead of a tag
    int siv;                     // POLYVAL + GCM-SIV tag (micro_aes.c:1454-1462); j0[] = nonce words
    uint8_t *tag_out;            // taglen bytes (16 for a shard's contribution), any alignment
    uint32_t taglen;             // GCM_TAG_LEN (micro_aes.c:1178): the leading bytes of the tag that are written
    GcmWork *work;
};

constexpr int kFinThreads = 256;     // latency-bound tree over <= one partial per warp of the grid

__global__ void __launch_bounds__(kFinThreads, 1) gcm_finish_kernel(const __grid_constant__ GcmFinishArgs a)
{
    uint4 *slot = a.work->partials;
    // multiplier between neighbouring partials: H^(chunk blocks) = (H^32)^rows; every thread
    // computes it for itself (no broadcast needed)
    Gf P = gf_pow_fast(gf_load(a.work->C32), a.chunk_rows);

    // slot[i] = partial of the i-th run counted from the end; total = sum_i slot[i] * P^i
    for (uint64_t n = a.nparts; n > 1; n = (n + 1) / 2) {
        const uint64_t half = (n + 1) / 2;
        for (uint64_t base = 0; base < half; base += kFinThreads) {
            const uint64_t m = base + threadIdx.x;
            Gf f{0, 0};
            if (m < half) {
                f = gf_load(slot[2 * m]);
                if (2 * m + 1 < n) {
                    const Gf g = gf_mul_fast(P, gf_load(slot[2 * m + 1]));
                    f.hi ^= g.hi; f.lo ^= g.lo;
                }
            }
            __syncthreads();
            if (m < half) slot[m] = gf_store(f);
            __syncthreads();
        }
        P = gf_mul_fast(P, P);
    }

    if (threadIdx.x != 0) return;
    const Gf H = gf_load(a.work->H);
    Gf S = a.nparts ? gf_load(slot[0]) : gf_load(a.work->aad_state);
    if (a.nparts_b) {                                         // the bitsliced warps' region follows
        S = gf_mul_fast(S, gf_pow_fast(H, a.blocks_b));
        for (uint64_t j = 0; j < a.nparts_b; ++j) {
            const Gf z = gf_load(slot[a.nparts + j]);
            S.hi ^= z.hi; S.lo ^= z.lo;
        }
    }

    const uint64_t nfull = a.len / 16;
    const uint32_t tail = (uint32_t)(a.len % 16);
    if (tail) {
        uint4 ct = load_block_bytes(a.in + 16 * nfull, tail);   // what GHASH absorbs (zero padded)
        if (a.siv) ct = rev_block(ct);
        if (a.mode != 1) {                                    // mixThenXor, micro_aes.c:949
            uint32_t w2, w3;
            ctr_words(a.b8, (a.v0 + nfull) & kMask56, w2, w3);
            uint32_t s[4] = {a.w0, a.w1, w2, w3};
            small_encrypt(a.ks.w, a.ks.rounds, s);
            const uint32_t keep[4] = {tail >= 4 ? 0xffffffffu : (1u << (8 * tail)) - 1,
                                      tail >= 8 ? 0xffffffffu : tail > 4 ? (1u << (8 * (tail - 4))) - 1 : 0,
                                      tail >= 12 ? 0xffffffffu : tail > 8 ? (1u << (8 * (tail - 8))) - 1 : 0,
                                      tail > 12 ? (1u << (8 * (tail - 12))) - 1 : 0};
            const uint32_t cw[4] = {(ct.x ^ s[0]) & keep[0], (ct.y ^ s[1]) & keep[1],
                                    (ct.z ^ s[2]) & keep[2], (ct.w ^ s[3]) & keep[3]};
            for (uint32_t i = 0; i < tail; ++i) a.out[16 * nfull + i] = (uint8_t)(cw[i >> 2] >> (8 * (i & 3)));
            if (a.mode == 0) ct = make_uint4(cw[0], cw[1], cw[2], cw[3]);
        }
        const Gf x = gf_load(ct);
        S.hi ^= x.hi; S.lo ^= x.lo;
        S = gf_mul_fast(H, S);
    }
    if (a.partial_only) {                                     // shard: hand the state to the combiner
        const uint4 sv = gf_store(S);
        const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
        for (uint32_t i = 0; i < 16; ++i) a.tag_out[i] = (uint8_t)(sw[i >> 2] >> (8 * (i & 3)));
        return;
    }
    if (a.siv) {
        // length block LE64(8*aadlen) || LE64(8*len) (micro_aes.c:1427-1429), byte-reversed
        S.hi ^= a.len * 8; S.lo ^= a.aadlen * 8;
        S = gf_mul_fast(H, S);
        const uint4 pv = rev_block(gf_store(S));              // = POLYVAL
        uint32_t t[4] = {pv.x ^ a.w0, pv.y ^ a.w1, pv.z ^ a.b8, pv.w & 0x7fffffffu};   // ^ nonce, clear bit 127
        small_encrypt(a.ks.w, a.ks.rounds, t);
        for (uint32_t i = 0; i < a.taglen; ++i) a.tag_out[i] = (uint8_t)(t[i >> 2] >> (8 * (i & 3)));
        return;
    }
    // length block: BE64(8*aadlen) || BE64(8*len)   (micro_aes.c:1130-1132)
    S.hi ^= a.aadlen * 8; S.lo ^= a.len * 8;
    S = gf_mul_fast(H, S);
    const uint4 ej0 = a.work->EJ0;
    uint4 tag = gf_store(S);
    tag.x ^= ej0.x; tag.y ^= ej0.y; tag.z ^= ej0.z; tag.w ^= ej0.w;
    const uint32_t tw[4] = {tag.x, tag.y, tag.z, tag.w};
    for (uint32_t i = 0; i < a.taglen; ++i) a.tag_out[i] = (uint8_t)(tw[i >> 2] >> (8 * (i & 3)));
}

static unsigned gcm_grid(uint64_t nchunks)
{
    const uint64_t need = (nchunks + kGcmWarps - 1) / kGcmWarps, sms = (uint64_t)sm_count();
    return (unsigned)(need < 1 ? 1 : need < sms ? need : sms);
}

template <int NR, int MODE, bool REV = false>
static cudaError_t launch_gcm_bulk_nr(const GcmBulkArgs &a, cudaStream_t st)
{
    cudaError_t e = opt_in_smem(gcm_bulk_kernel<NR, MODE, REV>);
    if (e != cudaSuccess) return e;
    gcm_bulk_kernel<NR, MODE, REV><<<gcm_grid(a.nchunks), kGcmThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

constexpr int kGcmDefaultShare = 170;    // of 1024: blocks given to the bitsliced warps; 632 / 661 / 671 / 680 / 648 / 568 GiB/s at
                                         // 0 / 120 / 150 / 180 / 200 / 250 for AES-128, 4 GiB (profiles/r2_sweep_gcm*.txt): the static split
                                         // falls off quickly once the bitsliced warps finish last, so stay left of the peak

#ifndef UAES_GCM_NARROW_DEFAULT
#define UAES_GCM_NARROW_DEFAULT 0           // measured: 655-666 GiB/s (8 warps, rolled absorb) and 665-670 (4 warps) against 676-678 for the wide form (profiles/r2_sweep_gcm8.txt)
#endif
#ifndef UAES_GCM8_DEFAULT_SHARE
#define UAES_GCM8_DEFAULT_SHARE 200
#endif
constexpr int kGcm8DefaultShare = UAES_GCM8_DEFAULT_SHARE;

template <int NR, int MODE>
static cudaError_t launch_gcm_hybrid8_nr(const GcmHybridArgs8 &h, cudaStream_t st)
{
    cudaError_t e = opt_in_smem(gcm_bulk_hybrid8_kernel<NR, MODE>);
    if (e != cudaSuccess) return e;
    gcm_bulk_hybrid8_kernel<NR, MODE><<<(unsigned)sm_count(), kGcmHybTtThreads + kGcm8BsThreads, kDynSmem, st>>>(h);
    ++g_launches;
    return cudaGetLastError();
}

template <int NR, int MODE>
static cudaError_t launch_gcm_hybrid_nr(const GcmHybridArgs &h, cudaStream_t st)
{
    cudaError_t e = opt_in_smem(gcm_bulk_hybrid_kernel<NR, MODE>);
    if (e != cudaSuccess) return e;
    gcm_bulk_hybrid_kernel<NR, MODE><<<(unsigned)sm_count(), kGcmHybTtThreads + kBsThreads, kDynSmem, st>>>(h);
    ++g_launches;
    return cudaGetLastError();
}

// One chunk per warp of the persistent grid: R rows each, so that all warps get the same share
// (the first chunk in message order is the short one).  Small messages get one row per chunk.
static void gcm_plan(uint64_t nblocks, uint64_t &rows_per_chunk, uint64_t &nchunks, int warps_per_cta = kGcmWarps)
{
    const uint64_t rows = (nblocks + 31) / 32;
    const uint64_t warps = (uint64_t)sm_count() * (uint64_t)warps_per_cta;
    rows_per_chunk = rows ? (rows + warps - 1) / warps : 1;
    nchunks = rows ? (nblocks + 32 * rows_per_chunk - 1) / (32 * rows_per_chunk) : 0;
}

// ---------------------------------------------------------------- GCM-SIV pieces

struct SivDeriveArgs {
    uaes_keysched master;
    uint32_t nonce[3];
    uint8_t *out;                // 8 bytes per derived half block, (2 + Nk/2) of them
};

// GCM_SIVsetup (micro_aes.c:1437-1451): E_K(LE32(i) || nonce), the first 8 bytes of each
__global__ void gcmsiv_derive_kernel(const __grid_constant__ SivDeriveArgs a)
{
    const uint32_t n = 2 + (a.master.rounds - 6) / 2;
    if (threadIdx.x >= n) return;
    uint32_t s[4] = {threadIdx.x, a.nonce[0], a.nonce[1], a.nonce[2]};
    small_encrypt(a.master.w, a.master.rounds, s);
    for (uint32_t i = 0; i < 8; ++i) a.out[8 * threadIdx.x + i] = (uint8_t)(s[i >> 2] >> (8 * (i & 3)));
}

struct Ctr32Args {
    uaes_keysched ks;
    const uint8_t *tag;          // 16 bytes in device memory: the initial counter block
    const uint4 *in;
    uint4 *out;
    uint64_t nblocks;
    uint32_t tail;
};

// CTR_cipher with mode SIVGCM_CTR (micro_aes.c:935-938): bit 7 of byte 15 forced, the counter is
// the little-endian 32-bit word in bytes 0..3 and wraps modulo 2^32.
//
// Same round-1/2 hoisting as ctr_kernel, for a counter that lives in COLUMN 0: inside a group of 256
// consecutive counter values only byte 0 changes, so round 1 column 0 = K0 ^ Te0[byte 0] and the
// other three columns are group constants (C3 follows counter byte 1, C2 byte 2, C1 byte 3); round 2
// column j = D_j ^ Te_x[a byte of column 0] with the Te_x terms a function of byte 0 alone, kept in
// registers per lane for the whole launch.  128 lookups per AES-128 block instead of 160.
constexpr int kCtr32Threads = 768;

template <int NR>
__global__ void __launch_bounds__(kCtr32Threads, 1) ctr32_kernel(const __grid_constant__ Ctr32Args a)
{
    constexpr int kWarps = kCtr32Threads / 32;
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_tables<true>(dyn);
    const uint32_t *rk = a.ks.w;
    const uint32_t lane = threadIdx.x & 31;
    const uint4 c = load_block_bytes(a.tag, 16);
    const uint32_t s1 = c.y ^ rk[1], s2 = c.z ^ rk[2], s3 = (c.w | 0x80000000u) ^ rk[3];

    // counter value of block k = (c.x + k) mod 2^32; u = c.x + k unreduced, group = u >> 8
    const uint64_t u0 = c.x;
    const uint32_t lowoff = c.x & 255u;
    const uint64_t g0 = u0 >> 8;
    const uint64_t ngroups = (lowoff + a.nblocks + 255) >> 8;
    const uint64_t wg = (uint64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const uint32_t half = (uint32_t)wg & 1u;
    const uint64_t npairs = (uint64_t)gridDim.x * (kWarps / 2);
    const uint64_t per = (ngroups + npairs - 1) / npairs;
    const uint64_t j0 = (wg >> 1) * per;
    const uint64_t j1 = j0 + per < ngroups ? j0 + per : ngroups;
    const uint32_t b0 = half * 128 + lane;                        // counter byte 0, + 32 * it

    auto kof = [&](uint64_t grp, int it) -> int64_t {
        return (int64_t)(grp << 8) + (int64_t)(b0 + 32 * it) - (int64_t)lowoff;
    };
    auto fetch = [&](uint64_t grp, int it) -> uint4 {
        const int64_t k = kof(grp, it);
        if (grp < j1 && k >= 0 && (uint64_t)k < a.nblocks) return ld_stream(a.in + k);
        return make_uint4(0, 0, 0, 0);
    };

    // launch constants: K0 and the byte-0 terms of round 2
    const uint32_t K0 = lut<1, kOffT1>(lb, s1) ^ lut<2, kOffT2>(lb, s2) ^ lut<3, kOffT3>(lb, s3) ^ rk[4];
    uint32_t U[4][4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const uint32_t c0 = K0 ^ lut<0, kOffT0>(lb, (b0 + 32 * it) ^ (rk[0] & 255u));
        U[it][0] = lut<0, kOffT0>(lb, c0); U[it][1] = lut<3, kOffT3>(lb, c0);
        U[it][2] = lut<2, kOffT2>(lb, c0); U[it][3] = lut<1, kOffT1>(lb, c0);
    }
    uint64_t tag16 = ~0ull;
    uint32_t E0 = 0, E1 = 0, E2 = 0, E3 = 0;

    uint4 cur = fetch(j0, 0);
    for (uint64_t j = j0; j < j1; ++j) {
        const uint32_t s0 = (uint32_t)((g0 + j) << 8) ^ rk[0];   // counter word with byte 0 = 0, wraps mod 2^32
        if (((g0 + j) >> 8) != tag16) {                           // counter bytes 2..3 changed: every 256 groups
            tag16 = (g0 + j) >> 8;
            const uint32_t C1 = lut<0, kOffT0>(lb, s1) ^ lut<1, kOffT1>(lb, s2) ^ lut<2, kOffT2>(lb, s3) ^ lut<3, kOffT3>(lb, s0) ^ rk[5];
            const uint32_t C2 = lut<0, kOffT0>(lb, s2) ^ lut<1, kOffT1>(lb, s3) ^ lut<2, kOffT2>(lb, s0) ^ lut<3, kOffT3>(lb, s1) ^ rk[6];
            E0 = lut<1, kOffT1>(lb, C1) ^ lut<2, kOffT2>(lb, C2) ^ rk[8];
            E1 = lut<0, kOffT0>(lb, C1) ^ lut<1, kOffT1>(lb, C2) ^ rk[9];
            E2 = lut<0, kOffT0>(lb, C2) ^ lut<3, kOffT3>(lb, C1) ^ rk[10];
            E3 = lut<2, kOffT2>(lb, C1) ^ lut<3, kOffT3>(lb, C2) ^ rk[11];
        }
        // per group: counter byte 1 enters through column 3 of round 1
        const uint32_t C3 = lut<0, kOffT0>(lb, s3) ^ lut<1, kOffT1>(lb, s0) ^ lut<2, kOffT2>(lb, s1) ^ lut<3, kOffT3>(lb, s2) ^ rk[7];
        const uint32_t D0 = E0 ^ lut<3, kOffT3>(lb, C3), D1 = E1 ^ lut<2, kOffT2>(lb, C3);
        const uint32_t D2 = E2 ^ lut<1, kOffT1>(lb, C3), D3 = E3 ^ lut<0, kOffT0>(lb, C3);
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const uint4 nxt = it < 3 ? fetch(j, it + 1) : fetch(j + 1, 0);
            const int64_t k = kof(j, it);
            uint32_t t0 = D0 ^ U[it][0], t1 = D1 ^ U[it][1], t2 = D2 ^ U[it][2], t3 = D3 ^ U[it][3];
            enc_finish<NR, 3>(lb, t0, t1, t2, t3, rk, cur.x, cur.y, cur.z, cur.w);
            if (k >= 0 && (uint64_t)k < a.nblocks) st_stream(a.out + k, make_uint4(t0, t1, t2, t3));
            cur = nxt;
        }
    }
    if (a.tail && blockIdx.x == 0 && threadIdx.x == 0) {
        uint32_t s[4] = {c.x + (uint32_t)a.nblocks, c.y, c.z, c.w | 0x80000000u};
        enc_block<NR>(lb, s[0], s[1], s[2], s[3], rk);
        const uint8_t *x = (const uint8_t *)(a.in + a.nblocks);
        uint8_t *y = (uint8_t *)(a.out + a.nblocks);
        for (uint32_t i = 0; i < a.tail; ++i) y[i] = x[i] ^ (uint8_t)(s[i >> 2] >> (8 * (i & 3)));
    }
}

template <int NR>
static cudaError_t launch_ctr32_nr(const Ctr32Args &a, cudaStream_t st)
{
    cudaError_t e = opt_in_smem(ctr32_kernel<NR>);
    if (e != cudaSuccess) return e;
    constexpr int kWarps = kCtr32Threads / 32;
    const uint64_t ngroups = (a.nblocks + 255 + 255) >> 8;      // upper bound (the offset is in device memory)
    const uint64_t ctas = (ngroups + 4 * (kWarps / 2) - 1) / (4 * (kWarps / 2)), sms = (uint64_t)sm_count();
    ctr32_kernel<NR><<<(unsigned)(ctas < 1 ? 1 : ctas < sms ? ctas : sms), kCtr32Threads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

// ---------------------------------------------------------------- multi-shard combine

// A message sharded over GPUs: shard r returns Z_r = sum over its blocks of
// X_i * H^(shard end - i).  GHASH of the whole = aad_state * H^(all blocks) ^ sum_r Z_r * H^(blocks
// after shard r), then the length block and E_K(J0) as usual.  Lane l takes contributions l, l+32, ...
// (a staged host-buffer call leaves one per chunk, a multi-GPU call one per rank).
struct GcmCombineArgs {
    uaes_keysched ks;
    uint32_t j0[4];
    const uint8_t *aad;
    uint64_t aadlen, len;
    const uint8_t *partials;     // nshards x 16 bytes (device)
    const uint64_t *after;       // GHASH blocks after the end of shard r (device)
    uint32_t nshards;
    uint64_t total_blocks;       // ceil(len / 16)
    uint8_t *tag_out;
    uint32_t taglen;             // bytes of the tag that are written (16 when fold_only)
    uint32_t fold_only;          // 1: write sum_r Z_r * H^after[r] and stop (streaming: many shards -> one)
};

__global__ void gcm_combine_kernel(const __grid_constant__ GcmCombineArgs a)
{
    __shared__ Gf sh_H, sh_ej0;
    const uint32_t lane = threadIdx.x;
    if (lane < 2) {
        uint32_t s[4] = {0, 0, 0, 0};
        if (lane == 1) { s[0] = a.j0[0]; s[1] = a.j0[1]; s[2] = a.j0[2]; s[3] = a.j0[3]; }
        small_encrypt(a.ks.w, a.ks.rounds, s);
        (lane == 0 ? sh_H : sh_ej0) = gf_from_words(s[0], s[1], s[2], s[3]);
    }
    __syncwarp();
    const Gf H = sh_H;
    Gf term{0, 0};
    for (uint32_t r = lane; r < a.nshards; r += 32) {
        const uint4 z = load_block_bytes(a.partials + 16 * (size_t)r, 16);
        const Gf t = gf_mul_fast(gf_load(z), gf_pow_fast(H, a.after[r]));
        term.hi ^= t.hi; term.lo ^= t.lo;
    }
    if (lane == 31 && !a.fold_only && a.aad && a.aadlen) {    // the AAD rides in front of block 0 (aad == NULL: it came as a shard)
        Gf g{0, 0};
        for (uint64_t off = 0; off < a.aadlen; off += 16) {
            const uint64_t left = a.aadlen - off;
            const Gf x = gf_load(load_block_bytes(a.aad + off, left < 16 ? (uint32_t)left : 16));
            g.hi ^= x.hi; g.lo ^= x.lo;
            g = gf_mul_fast(H, g);
        }
        const Gf t = gf_mul_fast(g, gf_pow_fast(H, a.total_blocks));
        term.hi ^= t.hi; term.lo ^= t.lo;
    }
    for (int o = 16; o; o >>= 1) {
        term.hi ^= __shfl_xor_sync(0xffffffffu, term.hi, o);
        term.lo ^= __shfl_xor_sync(0xffffffffu, term.lo, o);
    }
    if (lane) return;
    Gf S = term;
    if (!a.fold_only) {
        S.hi ^= a.aadlen * 8; S.lo ^= a.len * 8;
        S = gf_mul_fast(H, S);
        S.hi ^= sh_ej0.hi; S.lo ^= sh_ej0.lo;
    }
    const uint4 t = gf_store(S);
    const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
    for (uint32_t i = 0; i < a.taglen; ++i) a.tag_out[i] = (uint8_t)(tw[i >> 2] >> (8 * (i & 3)));
}

// J0 for a nonce that is not 12 bytes long (GCMsetup, micro_aes.c:1145-1149): GHASH_H({}, nonce),
// i.e. the zero-padded nonce blocks and the length block BE64(0) || BE64(8 * ivlen).  One thread.
struct GcmJ0Args {
    uaes_keysched ks;
    const uint8_t *iv;           // device memory
    uint64_t ivlen;
    uint8_t *out;                // 16 bytes, device memory
};

__global__ void gcm_j0_kernel(const __grid_constant__ GcmJ0Args a)
{
    if (threadIdx.x) return;
    uint32_t s[4] = {0, 0, 0, 0};
    small_encrypt(a.ks.w, a.ks.rounds, s);
    const Gf H = gf_from_words(s[0], s[1], s[2], s[3]);
    Gf g{0, 0};
    for (uint64_t off = 0; off < a.ivlen; off += 16) {
        const uint64_t left = a.ivlen - off;
        const Gf x = gf_load(load_block_bytes(a.iv + off, left < 16 ? (uint32_t)left : 16));
        g.hi ^= x.hi; g.lo ^= x.lo;
        g = gf_mul_fast(H, g);
    }
    g.lo ^= a.ivlen * 8;
    g = gf_mul_fast(H, g);
    const uint4 t = gf_store(g);
    const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
    for (uint32_t i = 0; i < 16; ++i) a.out[i] = (uint8_t)(tw[i >> 2] >> (8 * (i & 3)));
}

}  // namespace uaes

extern "C" size_t uaes_gcm_work_bytes(u64 len)
{
    using namespace uaes;
    uint64_t rows_per_chunk, nchunks;
    gcm_plan(len / 16, rows_per_chunk, nchunks);
    // + one partial per bitsliced warp of the co-runner forms (up to 8 per SM)
    return sizeof(GcmWork) + ((size_t)nchunks + 8 * (size_t)sm_count() + 8) * sizeof(uint4);
}

extern "C" int uaes_launch_gcm_j0(const uaes_keysched *ks, const void *iv_dev, u64 ivlen, void *out_dev,
                                  void *stream)
{
    using namespace uaes;
    GcmJ0Args a;
    a.ks = *ks; a.iv = (const uint8_t *)iv_dev; a.ivlen = ivlen; a.out = (uint8_t *)out_dev;
    gcm_j0_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    ++g_launches;
    return (int)cudaGetLastError();
}

extern "C" int uaes_launch_gcm(const uaes_keysched *ks, const unsigned char j0b[16], const void *aad_dev,
                               u64 aadlen, const void *aad_state_dev, const void *in, void *out, u64 len,
                               int mode, u64 first_block, int partial_only, void *tag_out, unsigned taglen,
                               void *work, void *stream)
{
    using namespace uaes;
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t nblocks = len / 16;
    uint64_t rows_per_chunk, nchunks;
    gcm_plan(nblocks, rows_per_chunk, nchunks);

    uint32_t j0[4];                                        // J0 = nonce || 00000001, or GHASH(nonce)
    for (int c = 0; c < 4; ++c)
        j0[c] = (uint32_t)j0b[4 * c] | (uint32_t)j0b[4 * c + 1] << 8 | (uint32_t)j0b[4 * c + 2] << 16 | (uint32_t)j0b[4 * c + 3] << 24;

    GcmSetupArgs s;
    s.ks = *ks;
    for (int c = 0; c < 4; ++c) s.j0[c] = j0[c];
    s.aad = (const uint8_t *)aad_dev; s.aadlen = aadlen; s.work = (GcmWork *)work;
    s.aad_state_in = (const uint4 *)aad_state_dev;
    s.polyval = 0;
    s.auth[0] = s.auth[1] = s.auth[2] = s.auth[3] = 0;
    gcm_setup_kernel<<<1, 32, 0, st>>>(s);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;

    // counter field of J0 = its bytes 9..15 as a 56-bit big-endian integer (micro_aes.c:421-427); data
    // starts at J0 + 1 (CCM_GCM pre-increment, micro_aes.c:939-941)
    uint64_t vj0 = 0;
    for (int i = 9; i < 16; ++i) vj0 = vj0 << 8 | j0b[i];
    const uint32_t b8 = j0b[8];

    uint64_t nparts_b = 0, blocks_b = 0;
    // the co-runner form: the last `share` of the blocks goes to bitsliced warps (region B starts on a
    // 1024-counter boundary); the on/off knobs are the CTR kernel's (uaes_ctr_tuning)
    ctr_tuning_init();
    const bool narrow = env_int("UAES_GCM_NARROW", UAES_GCM_NARROW_DEFAULT) != 0;
    const int share = g_ctr_share != kCtrDefaultShare ? g_ctr_share : env_int("UAES_GCM_BS_PERMILLE", narrow ? kGcm8DefaultShare : kGcmDefaultShare);
    if ((mode == 0 || mode == 2) && g_ctr_share > 0 && share > 0 && (long long)nblocks >= g_ctr_bs_min && nblocks >= 4096) {
        static thread_local GcmHybridArgs h;                 // 6 KB of planes: off the stack
        static thread_local GcmHybridArgs8 h8;
        const uint64_t v0 = (vj0 + 1 + first_block) & kMask56;
        const uint64_t uend = v0 + nblocks, want = nblocks / 1024 * (uint64_t)share;
        const uint64_t ub = (uend - want) & ~1023ull;
        if (ub > v0 + 1024 && ub < uend) {
            h.a_blocks = ub - v0;
            h.bs_passes = (uend - ub + 1023) >> 10;
            const uint64_t nw = (uint64_t)sm_count() * (narrow ? kGcm8BsThreads / 32 : kBsThreads / 32);
            h.bs_per = (h.bs_passes + nw - 1) / nw;
            h.nchunks_b = (h.bs_passes + h.bs_per - 1) / h.bs_per;
            gcm_plan(h.a_blocks, rows_per_chunk, nchunks, kGcmHybTtWarps);
            GcmBulkArgs &b = h.b;
            b.ks = *ks;
            b.w0 = j0[0]; b.w1 = j0[1]; b.b8 = b8; b.v0 = v0;
            b.in = (const uint4 *)in; b.out = (uint4 *)out;
            b.nblocks = nblocks; b.chunk_blocks = 32 * rows_per_chunk; b.nchunks = nchunks; b.work = (GcmWork *)work;
            nparts_b = h.nchunks_b; blocks_b = nblocks - h.a_blocks;
            if (narrow) {
                h8.b = h.b; h8.a_blocks = h.a_blocks; h8.bs_passes = h.bs_passes; h8.bs_per = h.bs_per; h8.nchunks_b = h.nchunks_b;
                bs8_make_key_planes(ks->w, ks->rounds, &h8.bs8);
                switch (ks->rounds * 4 + mode) {
                case 40: e = launch_gcm_hybrid8_nr<10, 0>(h8, st); break;
                case 42: e = launch_gcm_hybrid8_nr<10, 2>(h8, st); break;
                case 48: e = launch_gcm_hybrid8_nr<12, 0>(h8, st); break;
                case 50: e = launch_gcm_hybrid8_nr<12, 2>(h8, st); break;
                case 56: e = launch_gcm_hybrid8_nr<14, 0>(h8, st); break;
                case 58: e = launch_gcm_hybrid8_nr<14, 2>(h8, st); break;
                default: e = cudaErrorInvalidValue;
                }
            } else {
                bs_make_key_planes(ks->w, ks->rounds, &h.bs);
                switch (ks->rounds * 4 + mode) {
                case 40: e = launch_gcm_hybrid_nr<10, 0>(h, st); break;
                case 42: e = launch_gcm_hybrid_nr<10, 2>(h, st); break;
                case 48: e = launch_gcm_hybrid_nr<12, 0>(h, st); break;
                case 50: e = launch_gcm_hybrid_nr<12, 2>(h, st); break;
                case 56: e = launch_gcm_hybrid_nr<14, 0>(h, st); break;
                case 58: e = launch_gcm_hybrid_nr<14, 2>(h, st); break;
                default: e = cudaErrorInvalidValue;
                }
            }
            if (e != cudaSuccess) return (int)e;
        }
    }
    if (nchunks && !nparts_b) {
        GcmBulkArgs b;
        b.ks = *ks;
        b.w0 = j0[0]; b.w1 = j0[1]; b.b8 = b8; b.v0 = (vj0 + 1 + first_block) & kMask56;
        b.in = (const uint4 *)in; b.out = (uint4 *)out;
        b.nblocks = nblocks; b.chunk_blocks = 32 * rows_per_chunk; b.nchunks = nchunks; b.work = (GcmWork *)work;
        switch (ks->rounds * 4 + mode) {
        case 40: e = launch_gcm_bulk_nr<10, 0>(b, st); break;
        case 41: e = launch_gcm_bulk_nr<10, 1>(b, st); break;
        case 42: e = launch_gcm_bulk_nr<10, 2>(b, st); break;
        case 48: e = launch_gcm_bulk_nr<12, 0>(b, st); break;
        case 49: e = launch_gcm_bulk_nr<12, 1>(b, st); break;
        case 50: e = launch_gcm_bulk_nr<12, 2>(b, st); break;
        case 56: e = launch_gcm_bulk_nr<14, 0>(b, st); break;
        case 57: e = launch_gcm_bulk_nr<14, 1>(b, st); break;
        case 58: e = launch_gcm_bulk_nr<14, 2>(b, st); break;
        default: e = cudaErrorInvalidValue;
        }
        if (e != cudaSuccess) return (int)e;
    }

    GcmFinishArgs f;
    f.ks = *ks;
    f.w0 = j0[0]; f.w1 = j0[1]; f.b8 = b8; f.v0 = (vj0 + 1 + first_block) & kMask56;
    f.in = (const uint8_t *)in; f.out = (uint8_t *)out;
    f.len = len; f.aadlen = aadlen; f.nparts = nchunks; f.chunk_rows = rows_per_chunk; f.mode = mode; f.partial_only = partial_only; f.siv = 0;
    f.nparts_b = nparts_b; f.blocks_b = blocks_b;
    f.tag_out = (uint8_t *)tag_out; f.taglen = partial_only ? 16 : taglen; f.work = (GcmWork *)work;
    gcm_finish_kernel<<<1, kFinThreads, 0, st>>>(f);
    ++g_launches;
    return (int)cudaGetLastError();
}

extern "C" int uaes_launch_gcm_combine(const uaes_keysched *ks, const unsigned char j0b[16], const void *aad_dev,
                                       u64 aadlen, u64 len, const void *partials_dev, const void *after_dev,
                                       unsigned nshards, void *tag_out, unsigned taglen, void *stream)
{
    using namespace uaes;
    GcmCombineArgs a;
    a.ks = *ks;
    for (int c = 0; c < 4; ++c)
        a.j0[c] = (uint32_t)j0b[4 * c] | (uint32_t)j0b[4 * c + 1] << 8 | (uint32_t)j0b[4 * c + 2] << 16 | (uint32_t)j0b[4 * c + 3] << 24;
    a.aad = (const uint8_t *)aad_dev; a.aadlen = aadlen; a.len = len;
    a.partials = (const uint8_t *)partials_dev; a.after = (const uint64_t *)after_dev;
    a.nshards = nshards; a.total_blocks = (len + 15) / 16; a.tag_out = (uint8_t *)tag_out;
    a.taglen = taglen; a.fold_only = 0;
    gcm_combine_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    ++g_launches;
    return (int)cudaGetLastError();
}

// streaming: fold up to 31 shard contributions into the contribution of their union
extern "C" int uaes_launch_gcm_fold(const uaes_keysched *ks, const void *partials_dev, const void *after_dev,
                                    unsigned nshards, void *out_dev, void *stream)
{
    using namespace uaes;
    GcmCombineArgs a;
    a.ks = *ks;
    a.j0[0] = a.j0[1] = a.j0[2] = a.j0[3] = 0;
    a.aad = nullptr; a.aadlen = 0; a.len = 0;
    a.partials = (const uint8_t *)partials_dev; a.after = (const uint64_t *)after_dev;
    a.nshards = nshards; a.total_blocks = 0; a.tag_out = (uint8_t *)out_dev;
    a.taglen = 16; a.fold_only = 1;
    gcm_combine_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    ++g_launches;
    return (int)cudaGetLastError();
}

extern "C" int uaes_launch_gcmsiv_derive(const uaes_keysched *master, const unsigned char nonce[12],
                                         void *out_dev, void *stream)
{
    using namespace uaes;
    SivDeriveArgs a;
    a.master = *master;
    for (int c = 0; c < 3; ++c)
        a.nonce[c] = (uint32_t)nonce[4 * c] | (uint32_t)nonce[4 * c + 1] << 8 | (uint32_t)nonce[4 * c + 2] << 16 | (uint32_t)nonce[4 * c + 3] << 24;
    a.out = (uint8_t *)out_dev;
    gcmsiv_derive_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    ++g_launches;
    return (int)cudaGetLastError();
}

// POLYVAL(auth; aad, data) -> GCM-SIV tag = E_enc((POLYVAL ^ nonce) with bit 127 cleared)
extern "C" int uaes_launch_gcmsiv_tag(const uaes_keysched *enc, const unsigned char auth[16],
                                      const unsigned char nonce[12], const void *aad_dev, u64 aadlen,
                                      const void *aad_state_dev, const void *data, u64 len, int partial_only,
                                      void *tag_out, void *work, void *stream)
{
    using namespace uaes;
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t nblocks = len / 16;
    uint64_t rows_per_chunk, nchunks;
    gcm_plan(nblocks, rows_per_chunk, nchunks);
    uint32_t nw[3];
    for (int c = 0; c < 3; ++c)
        nw[c] = (uint32_t)nonce[4 * c] | (uint32_t)nonce[4 * c + 1] << 8 | (uint32_t)nonce[4 * c + 2] << 16 | (uint32_t)nonce[4 * c + 3] << 24;

    GcmSetupArgs s;
    s.ks = *enc;
    s.j0[0] = s.j0[1] = s.j0[2] = s.j0[3] = 0;
    s.aad = (const uint8_t *)aad_dev; s.aadlen = aadlen; s.work = (GcmWork *)work;
    s.aad_state_in = (const uint4 *)aad_state_dev;
    s.polyval = 1;
    for (int c = 0; c < 4; ++c)
        s.auth[c] = (uint32_t)auth[4 * c] | (uint32_t)auth[4 * c + 1] << 8 | (uint32_t)auth[4 * c + 2] << 16 | (uint32_t)auth[4 * c + 3] << 24;
    gcm_setup_kernel<<<1, 32, 0, st>>>(s);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;

    if (nchunks) {
        GcmBulkArgs b;
        b.ks = *enc;
        b.w0 = b.w1 = b.b8 = 0; b.v0 = 0;
        b.in = (const uint4 *)data; b.out = nullptr;
        b.nblocks = nblocks; b.chunk_blocks = 32 * rows_per_chunk; b.nchunks = nchunks; b.work = (GcmWork *)work;
        e = launch_gcm_bulk_nr<10, 1, true>(b, st);          // hash only: NR is irrelevant
        if (e != cudaSuccess) return (int)e;
    }
    GcmFinishArgs f;
    f.ks = *enc;
    f.w0 = nw[0]; f.w1 = nw[1]; f.b8 = nw[2]; f.v0 = 0;      // the nonce words ride in the counter fields
    f.in = (const uint8_t *)data; f.out = nullptr;
    f.len = len; f.aadlen = aadlen; f.nparts = nchunks; f.chunk_rows = rows_per_chunk;
    f.nparts_b = 0; f.blocks_b = 0;
    f.mode = 1; f.partial_only = partial_only; f.siv = 1;
    f.tag_out = (uint8_t *)tag_out; f.taglen = 16; f.work = (GcmWork *)work;
    gcm_finish_kernel<<<1, kFinThreads, 0, st>>>(f);
    ++g_launches;
    return (int)cudaGetLastError();
}

extern "C" int uaes_launch_ctr32(const uaes_keysched *enc, const void *tag_dev, const void *in, void *out,
                                 u64 len, void *stream)
{
    using namespace uaes;
    if (len == 0) return 0;
    Ctr32Args a;
    a.ks = *enc;
    a.tag = (const uint8_t *)tag_dev;
    a.in = (const uint4 *)in; a.out = (uint4 *)out;
    a.nblocks = len / 16; a.tail = (uint32_t)(len % 16);
    cudaStream_t st = (cudaStream_t)stream;
    switch (enc->rounds) {
    case 10: return (int)launch_ctr32_nr<10>(a, st);
    case 12: return (int)launch_ctr32_nr<12>(a, st);
    case 14: return (int)launch_ctr32_nr<14>(a, st);
    }
    return (int)cudaErrorInvalidValue;
}
