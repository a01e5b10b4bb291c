// uaes_xts.cuh -- XTS kernels (included by uaes_kernels.cu).
//
// Restates XTS_cipher (micro_aes.c:1008-1055): Y_j = T_j ^ Cipher_K1(T_j ^ X_j) with
// T_0 = E_K2(tweak), T_{j+1} = alpha * T_j, plus ciphertext stealing for ragged units.  The
// reference's serial tweak chain becomes T_0 * alpha^j computed from the block index.
#pragma once

namespace uaes {

// geometry of the kernels with the bitsliced co-runner (table-driven threads; share of the work, per 1024)
constexpr int kXtsTtThreads = 384;
constexpr int kXtsDefaultShare = 165;
constexpr int kXtsUnitDefaultShare = 140;    // one large data unit (static split)
constexpr int kXtsDecDefaultShare = 60;      // decryption; only its being > 0 matters when the work queue is on
#ifndef UAES_XTS_QUEUE_DEFAULT
#define UAES_XTS_QUEUE_DEFAULT 1
#endif
constexpr int kXtsQueueDefault = UAES_XTS_QUEUE_DEFAULT;   // 1: work queue instead of the static share (encryption with the co-runner)

// Encryption with ONLY Te0 available at table offset OFF (decrypt kernels keep Te0 next to the
// inverse tables so that they can still encrypt sector tweaks): Te_k = Te0 rotated by 8k bits.
template <int NR, uint32_t OFF>
__device__ __forceinline__ void enc_block_te0(uint32_t lb, uint32_t &s0, uint32_t &s1, uint32_t &s2,
                                              uint32_t &s3, const uint32_t *rk)
{
    s0 ^= rk[0]; s1 ^= rk[1]; s2 ^= rk[2]; s3 ^= rk[3];
#pragma unroll
    for (int r = 1; r < NR; ++r) {
        const uint32_t t0 = lut<0, OFF>(lb, s0) ^ rotl32(lut<1, OFF>(lb, s1), 8) ^ rotl32(lut<2, OFF>(lb, s2), 16) ^ rotl32(lut<3, OFF>(lb, s3), 24) ^ rk[4 * r + 0];
        const uint32_t t1 = lut<0, OFF>(lb, s1) ^ rotl32(lut<1, OFF>(lb, s2), 8) ^ rotl32(lut<2, OFF>(lb, s3), 16) ^ rotl32(lut<3, OFF>(lb, s0), 24) ^ rk[4 * r + 1];
        const uint32_t t2 = lut<0, OFF>(lb, s2) ^ rotl32(lut<1, OFF>(lb, s3), 8) ^ rotl32(lut<2, OFF>(lb, s0), 16) ^ rotl32(lut<3, OFF>(lb, s1), 24) ^ rk[4 * r + 2];
        const uint32_t t3 = lut<0, OFF>(lb, s3) ^ rotl32(lut<1, OFF>(lb, s0), 8) ^ rotl32(lut<2, OFF>(lb, s1), 16) ^ rotl32(lut<3, OFF>(lb, s2), 24) ^ rk[4 * r + 3];
        s0 = t0; s1 = t1; s2 = t2; s3 = t3;
    }
    // Te0 = {2S, S, S, 3S}: S(x) sits in bytes 1 and 2
    auto last = [&](uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
        const uint32_t lo = __byte_perm(lut<0, OFF>(lb, a), lut<1, OFF>(lb, b), 0x0051);
        const uint32_t hi = __byte_perm(lut<2, OFF>(lb, c), lut<3, OFF>(lb, d), 0x6200);
        return __byte_perm(lo, hi, 0x7610);
    };
    const uint32_t o0 = last(s0, s1, s2, s3) ^ rk[4 * NR + 0], o1 = last(s1, s2, s3, s0) ^ rk[4 * NR + 1];
    const uint32_t o2 = last(s2, s3, s0, s1) ^ rk[4 * NR + 2], o3 = last(s3, s0, s1, s2) ^ rk[4 * NR + 3];
    s0 = o0; s1 = o1; s2 = o2; s3 = o3;
}

// decrypt kernels use init_dec_tables (uaes_tables.cuh): the Td tables, Td4 and Te0 for the tweaks
template <bool ENC>
__device__ __forceinline__ uint32_t setup_xts_tables(const void *dyn)
{
    const uint32_t base = ENC ? align_table_base(dyn) : dec_table_base(dyn);
    if (base + (ENC ? kEncTableBytes : kDecTableBytes) > smem_u32(dyn) + dyn_smem_size()) __trap();
    if (ENC) init_enc_tables(base); else init_dec_tables(base);
    __syncthreads();
    uint32_t lanebase = base + (threadIdx.x & 31) * 4;
    asm volatile("" : "+r"(lanebase)::"memory");
    return lanebase;
}

// one XEX block: s = Cipher(s ^ T) ^ T
template <int NR, bool ENC>
__device__ __forceinline__ void xex_block(uint32_t lb, uint32_t &s0, uint32_t &s1, uint32_t &s2,
                                          uint32_t &s3, const uint32_t *k1, Tweak t)
{
    uint32_t t0, t1, t2, t3;
    tweak_words(t, t0, t1, t2, t3);
    s0 ^= t0; s1 ^= t1; s2 ^= t2; s3 ^= t3;
    if (ENC) enc_block<NR>(lb, s0, s1, s2, s3, k1, t0, t1, t2, t3);
    else     dec_block<NR>(lb, s0, s1, s2, s3, k1, t0, t1, t2, t3);
}

// ---------------------------------------------------------------- batched sectors

struct XtsSectorArgs {
    uaes_keysched k1, k2;
    uint64_t first_sector, sector_blocks, nsectors;
    const uint4 *in;
    uint4 *out;
};

// A warp takes a tile of 32 consecutive sectors: lane l encrypts the tweak of sector l (one AES
// per 32 sectors per lane), then the warp walks the sectors; for each one the sector's T_0 is
// broadcast by shuffle and lane l takes block 32*row + l with tweak T_0 * alpha^(32*row + l).
// A 512-byte sector is exactly one coalesced row.
template <int NR, bool ENC>
__global__ void __launch_bounds__(kThreads, 1) xts_sectors_kernel(const __grid_constant__ XtsSectorArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_xts_tables<ENC>(dyn);
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t ntiles = (a.nsectors + 31) / 32;
    const uint64_t nwarps = (uint64_t)gridDim.x * kWarpsPerCta;
    const uint64_t sb = a.sector_blocks;

    for (uint64_t tile = (uint64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); tile < ntiles; tile += nwarps) {
        // T_0 of sector tile*32 + lane: E_K2(LE128(sector)), micro_aes.c:1017-1027
        const uint64_t sec = a.first_sector + tile * 32 + lane;
        uint32_t e0 = (uint32_t)sec, e1 = (uint32_t)(sec >> 32), e2 = 0, e3 = 0;
        if (ENC) enc_block<NR>(lb, e0, e1, e2, e3, a.k2.w);
        else     enc_block_te0<NR, kOffDecTe0>(lb, e0, e1, e2, e3, a.k2.w);

        const uint64_t left = a.nsectors - tile * 32;
        const int nsec = left < 32 ? (int)left : 32;
        for (int s = 0; s < nsec; ++s) {
            Tweak t0;
            t0.lo = (uint64_t)__shfl_sync(0xffffffffu, e1, s) << 32 | __shfl_sync(0xffffffffu, e0, s);
            t0.hi = (uint64_t)__shfl_sync(0xffffffffu, e3, s) << 32 | __shfl_sync(0xffffffffu, e2, s);
            Tweak t = xts_shl(t0, lane);                        // alpha^lane
            const uint64_t base = (tile * 32 + s) * sb;
            for (uint64_t j = lane; j < sb; j += 32) {
                uint4 v = ld_stream(a.in + base + j);
                xex_block<NR, ENC>(lb, v.x, v.y, v.z, v.w, a.k1.w, t);
                st_stream(a.out + base + j, v);
                t = xts_shl(t, 32);                             // next row of this sector
            }
        }
    }
}

// ---------------------------------------------------------------- one data unit

struct XtsUnitArgs {
    uaes_keysched k1, k1e, k2;   // bulk-direction schedule, K1 encryption schedule, K2 schedule
    uint32_t tweak[4];
    const uint4 *in;
    uint4 *out;
    uint64_t nblocks;            // blocks handled by the plain loop: len/16 - (len%16 != 0)
    uint64_t first_block;        // position of in[0] inside the data unit (a range of a larger unit: a
                                 // staged chunk, or one GPU's share); tweak of block k = T_0 * alpha^(first_block + k)
    uint32_t tail;               // len % 16 (stealing when non-zero; only the range that ends the unit has one)
};

// ciphertext stealing (micro_aes.c:1037-1053) on one thread with the byte-wise cipher
template <bool ENC>
__device__ inline void xts_steal_tail(const XtsUnitArgs &a, const Tweak &T0)
{
    const uint64_t m = a.nblocks;                              // index of the last full block
    const Tweak Tm = xts_jump(T0, a.first_block + m), Tn = xts_shl(Tm, 1);
    const uint8_t *x = (const uint8_t *)(a.in + m);
    uint8_t *y = (uint8_t *)(a.out + m);
    uint8_t first[16], part[16];
    for (int i = 0; i < 16; ++i) first[i] = x[i];
    for (uint32_t i = 0; i < a.tail; ++i) part[i] = x[16 + i];
    auto xex = [&](uint8_t b[16], Tweak t) {
        uint32_t tw[4], s[4];
        tweak_words(t, tw[0], tw[1], tw[2], tw[3]);
        for (int c = 0; c < 4; ++c)
            s[c] = ((uint32_t)b[4 * c] | (uint32_t)b[4 * c + 1] << 8 | (uint32_t)b[4 * c + 2] << 16 | (uint32_t)b[4 * c + 3] << 24) ^ tw[c];
        if (ENC) small_encrypt(a.k1e.w, a.k1e.rounds, s); else small_decrypt(a.k1e.w, a.k1e.rounds, s);
        for (int c = 0; c < 4; ++c) {
            s[c] ^= tw[c];
            for (int i = 0; i < 4; ++i) b[4 * c + i] = (uint8_t)(s[c] >> (8 * i));
        }
    };
    // encrypt: block m under T_m, the stolen block under alpha*T_m; decrypt: the other way round
    xex(first, ENC ? Tm : Tn);
    for (uint32_t i = a.tail; i < 16; ++i) part[i] = first[i];
    xex(part, ENC ? Tn : Tm);
    for (int i = 0; i < 16; ++i) y[i] = part[i];
    for (uint32_t i = 0; i < a.tail; ++i) y[16 + i] = first[i];
}

// The unit is cut into rows of 32 blocks; every warp owns a contiguous run of rows, jumps to
// T_0 * alpha^(first block) once (ladder of squarings of alpha^128) and then steps alpha^32 per row.
template <int NR, bool ENC>
__global__ void __launch_bounds__(kThreads, 1) xts_unit_kernel(const __grid_constant__ XtsUnitArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    // T_0 is handed to the other warps through the last 16 bytes of the dynamic window (the tables
    // end at least 35 KB earlier); static shared memory would push the launch over the 227 KB limit
    volatile uint32_t *t0s = (volatile uint32_t *)(dyn + dyn_smem_size() - 16);
    if (threadIdx.x == 0) {                                   // T_0 = E_K2(tweak), micro_aes.c:1026-1027
        uint32_t s[4] = {a.tweak[0], a.tweak[1], a.tweak[2], a.tweak[3]};
        small_encrypt(a.k2.w, a.k2.rounds, s);
        t0s[0] = s[0]; t0s[1] = s[1]; t0s[2] = s[2]; t0s[3] = s[3];
    }
    const uint32_t lb = setup_xts_tables<ENC>(dyn);           // contains __syncthreads()
    const uint32_t lane = threadIdx.x & 31;
    Tweak T0;
    T0.lo = (uint64_t)t0s[1] << 32 | t0s[0];
    T0.hi = (uint64_t)t0s[3] << 32 | t0s[2];

    const uint64_t rows = (a.nblocks + 31) / 32;
    const uint64_t nwarps = (uint64_t)gridDim.x * kWarpsPerCta;
    const uint64_t rpw = (rows + nwarps - 1) / nwarps;
    const uint64_t warp = (uint64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const uint64_t r0 = warp * rpw, r1 = r0 + rpw < rows ? r0 + rpw : rows;

    if (r0 < r1) {
        Tweak t = xts_jump(T0, a.first_block + r0 * 32 + lane);
        uint64_t k = r0 * 32 + lane;
        uint4 cur = k < a.nblocks ? ld_stream(a.in + k) : make_uint4(0, 0, 0, 0);
        for (uint64_t r = r0; r < r1; ++r, k += 32) {
            const uint4 nxt = (r + 1 < r1 && k + 32 < a.nblocks) ? ld_stream(a.in + k + 32) : make_uint4(0, 0, 0, 0);
            uint4 v = cur;
            xex_block<NR, ENC>(lb, v.x, v.y, v.z, v.w, a.k1.w, t);
            if (k < a.nblocks) st_stream(a.out + k, v);
            t = xts_shl(t, 32);
            cur = nxt;
        }
    }

    if (a.tail && blockIdx.x == 0 && threadIdx.x == 0) xts_steal_tail<ENC>(a, T0);
}

// ---- one data unit, encryption, with the co-runner (the reference's AES_XTS_encrypt on a large
// buffer): rows of 32 blocks; table-driven warps own runs of row pairs, bitsliced warps runs of
// 32-row tiles; every warp jumps to its first tweak once and then steps alpha^32 per row.
struct XtsUnitHybridArgs {
    XtsUnitArgs u;
    uint64_t tt_blocks;          // blocks [0, tt_blocks): table-driven warps; a multiple of 1024
    BsKeyPlanesFull bs;
};

// table-driven side of the single-unit kernel with a co-runner: as in xts_sectors_hybrid_kernel
#ifndef UAES_XTSU_TT
#define UAES_XTSU_TT 512
#endif
#ifndef UAES_XTSU_ILP
#define UAES_XTSU_ILP 1
#endif
#ifndef UAES_XTSU_TT_REGS
#define UAES_XTSU_TT_REGS 64
#endif
constexpr int kXtsUnitTt = UAES_XTSU_TT;

template <int NR>
__global__ void __launch_bounds__(kXtsUnitTt + kBsThreads, 1) xts_unit_hybrid_kernel(const __grid_constant__ XtsUnitHybridArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    volatile uint32_t *t0s = (volatile uint32_t *)(dyn + dyn_smem_size() - 16);
    if (threadIdx.x == 0) {                                   // T_0 = E_K2(tweak), micro_aes.c:1026-1027
        uint32_t s[4] = {a.u.tweak[0], a.u.tweak[1], a.u.tweak[2], a.u.tweak[3]};
        small_encrypt(a.u.k2.w, a.u.k2.rounds, s);
        t0s[0] = s[0]; t0s[1] = s[1]; t0s[2] = s[2]; t0s[3] = s[3];
    }
    const uint32_t lb = setup_xts_tables<true>(dyn);          // contains __syncthreads()
    const uint32_t lane = threadIdx.x & 31;
    Tweak T0;
    T0.lo = (uint64_t)t0s[1] << 32 | t0s[0];
    T0.hi = (uint64_t)t0s[3] << 32 | t0s[2];
    constexpr int kTtWarps = kXtsUnitTt / 32;
    constexpr int kLaunchRegs = (65536 / (kXtsUnitTt + kBsThreads)) / 8 * 8;
    constexpr int kTtRegs = UAES_XTSU_TT_REGS, kBsRegs = (kLaunchRegs + (kLaunchRegs - kTtRegs) * kXtsUnitTt / kBsThreads) / 8 * 8;
    const uint64_t nblocks = a.u.nblocks;

    if (threadIdx.x >= kXtsUnitTt) {
        reg_inc<kBsRegs>();
        const uint64_t ntiles = (nblocks - a.tt_blocks + 1023) / 1024;
        const uint64_t gw = (uint64_t)blockIdx.x * (kBsThreads / 32) + ((threadIdx.x - kXtsUnitTt) >> 5);
        const uint64_t nw = (uint64_t)gridDim.x * (kBsThreads / 32);
        const uint64_t per = (ntiles + nw - 1) / nw;
        const uint64_t p0 = gw * per < ntiles ? gw * per : ntiles;
        const uint64_t p1 = p0 + per < ntiles ? p0 + per : ntiles;
        if (p0 >= p1) return;
        Tweak tstart = xts_jump(T0, a.u.first_block + a.tt_blocks + p0 * 1024 + lane);
        for (uint64_t tile = p0; tile < p1; ++tile) {
            const uint64_t kb = a.tt_blocks + tile * 1024 + lane;
            Tweak t = tstart;                                  // only tstart stays live across the rounds
            uint32_t s[128];
#pragma unroll
            for (int tb = 0; tb < 32; tb += 8) {
                uint4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = kb + 32 * (tb + i) < nblocks ? ld_stream(a.u.in + kb + 32 * (tb + i)) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint32_t w0, w1, w2, w3;
                    tweak_words(t, w0, w1, w2, w3);
                    s[tb + i] = v[i].x ^ w0; s[32 + tb + i] = v[i].y ^ w1; s[64 + tb + i] = v[i].z ^ w2; s[96 + tb + i] = v[i].w ^ w3;
                    t = xts_shl(t, 32);
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
            bs_encrypt_planes<NR>(s, a.bs);
#pragma unroll
            for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
            Tweak u = tstart;
#pragma unroll
            for (int tt = 0; tt < 32; ++tt) {
                uint32_t w0, w1, w2, w3;
                tweak_words(u, w0, w1, w2, w3);
                if (kb + 32 * tt < nblocks)
                    st_stream(a.u.out + kb + 32 * tt, make_uint4(s[tt] ^ w0, s[32 + tt] ^ w1, s[64 + tt] ^ w2, s[96 + tt] ^ w3));
                u = xts_shl(u, 32);
            }
            tstart = u;                                        // alpha^1024 further: the next tile's first tweak
        }
        return;
    }
    reg_dec<kTtRegs>();

    const uint32_t *k1 = a.u.k1.w;
    constexpr int ILP = UAES_XTSU_ILP;                           // rows in flight per thread
    const uint64_t nsteps = a.tt_blocks / (32 * ILP);            // tt_blocks is a multiple of 1024
    const uint64_t gw = (uint64_t)blockIdx.x * kTtWarps + (threadIdx.x >> 5);
    const uint64_t nw = (uint64_t)gridDim.x * kTtWarps;
    const uint64_t per = (nsteps + nw - 1) / nw;
    const uint64_t q0 = gw * per < nsteps ? gw * per : nsteps;
    const uint64_t q1 = q0 + per < nsteps ? q0 + per : nsteps;
    if (q0 < q1) {
        Tweak t = xts_jump(T0, a.u.first_block + q0 * (32 * ILP) + lane);
        uint4 cur[ILP], nxt[ILP];
#pragma unroll
        for (int i = 0; i < ILP; ++i) cur[i] = ld_stream(a.u.in + q0 * (32 * ILP) + 32 * i + lane);
        for (uint64_t q = q0; q < q1; ++q) {
            const uint64_t k = q * (32 * ILP) + lane;
            uint32_t st[ILP][4];
            uint4 tw[ILP];
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (q + 1 < q1) nxt[i] = ld_stream(a.u.in + k + 32 * ILP + 32 * i);
                tweak_words(t, tw[i].x, tw[i].y, tw[i].z, tw[i].w);
                t = xts_shl(t, 32);
                st[i][0] = cur[i].x ^ tw[i].x ^ k1[0]; st[i][1] = cur[i].y ^ tw[i].y ^ k1[1];
                st[i][2] = cur[i].z ^ tw[i].z ^ k1[2]; st[i][3] = cur[i].w ^ tw[i].w ^ k1[3];
            }
            enc_finish_n<NR, 1, ILP>(lb, st, k1, tw);
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                st_stream(a.u.out + k + 32 * i, make_uint4(st[i][0], st[i][1], st[i][2], st[i][3]));
                cur[i] = nxt[i];
            }
        }
    }
    if (a.u.tail && blockIdx.x == 0 && threadIdx.x == 0) xts_steal_tail<true>(a.u, T0);
}

template <int NR, bool ENC>
static cudaError_t launch_xts_sectors_nr(const XtsSectorArgs &a, cudaStream_t st)
{
    cudaError_t e = opt_in_smem(xts_sectors_kernel<NR, ENC>);
    if (e != cudaSuccess) return e;
    xts_sectors_kernel<NR, ENC><<<grid_for((a.nsectors + 31) / 32), kThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

// ---------------------------------------------------------------- batched sectors, with the co-runner
//
// 512-byte sectors (32 blocks), both directions: the same two-kinds-of-warps kernel as ctr_kernel.
// 12 table-driven warps (two sectors in flight each) keep the lookup pipe at its roof; one warpgroup
// of bitsliced warps (uaes_bitslice.cuh, general form) encrypts tiles of 32 sectors on the ALU pipe:
// lane l, slot t <-> block l of sector t, so every load / store of a slot is the sector's coalesced
// 512-byte row.  The tweaks stay in the ordinary layout: T_0 of sector t is one table-driven
// encryption on lane t, T_0 * alpha^l is a shift, and the XEX whitening happens on the blocks'
// words before the transposes into planes and after the transposes back.
struct XtsHybridArgs {
    XtsSectorArgs x;             // sector_blocks == 32
    uint64_t tt_tiles;           // static split: tiles [0, tt_tiles) of 32 sectors: table-driven warps
    uint64_t ntiles;             // the rest: bitsliced warps
    unsigned long long *q;       // non-null: no static split -- the two-ended work queue of ctr_queue_kernel, unit = one tile
    uint32_t q_zero;             // 0 (see q_post)
    BsKeyPlanesFull bs;          // x.k1 (encryption schedule, or the inverse schedule) as planes
};

template <int NR, bool ENC>
__device__ __forceinline__ void xts_bitsliced_warp(const XtsHybridArgs &a, uint32_t lb, uint64_t t0, uint64_t t1)
{
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t tile = t0; tile < t1; ++tile) {
        const uint64_t sec0 = tile * 32;
        const uint64_t left = a.x.nsectors - sec0;
        const int nsec = left < 32 ? (int)left : 32;
        // T_0 of sector sec0 + lane (micro_aes.c:1017-1027)
        const uint64_t sec = a.x.first_sector + sec0 + lane;
        uint32_t e0 = (uint32_t)sec, e1 = (uint32_t)(sec >> 32), e2 = 0, e3 = 0;
        if (ENC) enc_block<NR>(lb, e0, e1, e2, e3, a.x.k2.w);
        else     enc_block_te0<NR, kOffDecTe0>(lb, e0, e1, e2, e3, a.x.k2.w);
        auto tweak_of = [&](int t, uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3) {
            Tweak tw;
            tw.lo = (uint64_t)__shfl_sync(0xffffffffu, e1, t) << 32 | __shfl_sync(0xffffffffu, e0, t);
            tw.hi = (uint64_t)__shfl_sync(0xffffffffu, e3, t) << 32 | __shfl_sync(0xffffffffu, e2, t);
            tweak_words(xts_shl(tw, lane), w0, w1, w2, w3);
        };
        const uint4 *src = a.x.in + sec0 * 32 + lane;
        uint4 *dst = a.x.out + sec0 * 32 + lane;
        uint32_t s[128];
#pragma unroll
        for (int tb = 0; tb < 32; tb += kBsLoadBatch) {          // loads in batches of rows
            uint4 v[kBsLoadBatch];
#pragma unroll
            for (int i = 0; i < kBsLoadBatch; ++i) v[i] = tb + i < nsec ? ld_stream(src + (tb + i) * 32) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int i = 0; i < kBsLoadBatch; ++i) {
                uint32_t w0, w1, w2, w3;
                tweak_of(tb + i, w0, w1, w2, w3);
                s[tb + i] = v[i].x ^ w0; s[32 + tb + i] = v[i].y ^ w1;
                s[64 + tb + i] = v[i].z ^ w2; s[96 + tb + i] = v[i].w ^ w3;
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
        if (ENC) bs_encrypt_planes<NR>(s, a.bs); else bs_decrypt_planes<NR>(s, a.bs);
#pragma unroll
        for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
#pragma unroll
        for (int t = 0; t < 32; ++t) {
            uint32_t w0, w1, w2, w3;
            tweak_of(t, w0, w1, w2, w3);
            if (t < nsec) st_stream(dst + t * 32, make_uint4(s[t] ^ w0, s[32 + t] ^ w1, s[64 + t] ^ w2, s[96 + t] ^ w3));
        }
    }
}


// the table-driven role of the sector kernels with a co-runner; tt_warp = this warp's number among the CTA's table-driven warps
template <int NR, bool ENC, int ILP, int TTW, class ARGS>
__device__ __forceinline__ void xts_hybrid_tt_role(const ARGS &a, uint32_t lb, uint32_t tt_warp)
{
    const uint32_t lane = threadIdx.x & 31;
    constexpr int kTtWarps = TTW;                // table-driven warps per CTA
    // table-driven warps: a contiguous run of tiles each (static split) or tiles claimed from the FRONT of
    // the work queue, ILP sectors in flight
    const uint64_t gw = (uint64_t)blockIdx.x * kTtWarps + tt_warp;
    const uint64_t nw = (uint64_t)gridDim.x * kTtWarps;
    const uint64_t per = (a.tt_tiles + nw - 1) / nw;
    const uint64_t q0 = a.q ? q_front(q_post(a.q, 1ull, a.q_zero), a.ntiles) : gw * per < a.tt_tiles ? gw * per : a.tt_tiles;
    const uint64_t q1 = a.q ? kQNone : q0 + per < a.tt_tiles ? q0 + per : a.tt_tiles;
    const uint32_t *k1 = a.x.k1.w;
    unsigned long long posted = 0;
    for (uint64_t tile = q0; tile < q1; tile = a.q ? q_front(posted, a.ntiles) : tile + 1) {
        if (a.q) posted = q_post(a.q, 1ull, a.q_zero);       // the next tile, a tile ahead
        const uint64_t sec0 = tile * 32;
        const uint64_t left = a.x.nsectors - sec0;
        const int nsec = left < 32 ? (int)left : 32;
        const uint64_t sec = a.x.first_sector + sec0 + lane;
        const uint4 *src = a.x.in + sec0 * 32 + lane;
        uint4 *dst = a.x.out + sec0 * 32 + lane;
        uint4 cur[ILP], nxt[ILP];
        // the tile's first two sectors are requested BEFORE the 32 tweak encryptions: their latency hides
        // behind those (ncu had the table-driven warps at 12 % long-scoreboard stalls, one cold load per tile):
        // 591-596 -> 613 GiB/s; an L2 prefetch of the NEXT tile from the middle of this one made it worse (608),
        // profiles/r2_sweep_xts_loads_first.txt
#pragma unroll
        for (int i = 0; i < ILP; ++i) cur[i] = i < nsec ? ld_stream(src + i * 32) : make_uint4(0, 0, 0, 0);
        uint32_t e0 = (uint32_t)sec, e1 = (uint32_t)(sec >> 32), e2 = 0, e3 = 0;
        if (ENC) enc_block<NR>(lb, e0, e1, e2, e3, a.x.k2.w);
        else     enc_block_te0<NR, kOffDecTe0>(lb, e0, e1, e2, e3, a.x.k2.w);
        for (int sct = 0; sct < nsec; sct += ILP) {
            uint32_t st[ILP][4];
            uint4 tw[ILP];
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                nxt[i] = sct + ILP + i < nsec ? ld_stream(src + (sct + ILP + i) * 32) : make_uint4(0, 0, 0, 0);
                Tweak t;
                t.lo = (uint64_t)__shfl_sync(0xffffffffu, e1, (sct + i) & 31) << 32 | __shfl_sync(0xffffffffu, e0, (sct + i) & 31);
                t.hi = (uint64_t)__shfl_sync(0xffffffffu, e3, (sct + i) & 31) << 32 | __shfl_sync(0xffffffffu, e2, (sct + i) & 31);
                tweak_words(xts_shl(t, lane), tw[i].x, tw[i].y, tw[i].z, tw[i].w);
                st[i][0] = cur[i].x ^ tw[i].x; st[i][1] = cur[i].y ^ tw[i].y;
                st[i][2] = cur[i].z ^ tw[i].z; st[i][3] = cur[i].w ^ tw[i].w;
                if (ENC) { st[i][0] ^= k1[0]; st[i][1] ^= k1[1]; st[i][2] ^= k1[2]; st[i][3] ^= k1[3]; }
            }
            if (ENC) enc_finish_n<NR, 1, ILP>(lb, st, k1, tw); else dec_block_n<NR, ILP>(lb, st, k1, tw);
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (sct + i < nsec) st_stream(dst + (sct + i) * 32, make_uint4(st[i][0], st[i][1], st[i][2], st[i][3]));
                cur[i] = nxt[i];
            }
        }
    }
}

// geometry of the table-driven side (profiles/r2_sweep_xts_ilp.txt): UAES_XTS_TT threads with UAES_XTS_ILP sectors in flight
// 16 warps with ONE sector in flight at 64 registers (+ the 4 bitsliced warps at 224): 651.8 GiB/s for AES-256 against 614.7 for
// 12 warps with two sectors in flight at 96; 20 warps at 56 registers: 646.5; 16 at 72 (co-runner 192): 650.2
#ifndef UAES_XTS_TT
#define UAES_XTS_TT 512
#endif
#ifndef UAES_XTS_ILP
#define UAES_XTS_ILP 1
#endif
#ifndef UAES_XTS_TT_REGS
#define UAES_XTS_TT_REGS 64
#endif
constexpr int kXtsHybTt = UAES_XTS_TT;

template <int NR, bool ENC>
__global__ void __launch_bounds__(kXtsHybTt + kBsThreads, 1) xts_sectors_hybrid_kernel(const __grid_constant__ XtsHybridArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_xts_tables<ENC>(dyn);
    constexpr int kTtWarps = kXtsHybTt / 32;
    constexpr int kLaunchRegs = (65536 / (kXtsHybTt + kBsThreads)) / 8 * 8;          // 128 for 384 + 128 threads
    // 96 / 224: 582 -> 596 GiB/s for AES-256 (profiles/r2_sweep_hybrid_regs.txt); ECB and OCB keep 104 / 200
    constexpr int kTtRegs = UAES_XTS_TT_REGS, kBsRegs = (kLaunchRegs + (kLaunchRegs - kTtRegs) * kXtsHybTt / kBsThreads) / 8 * 8;

    if (threadIdx.x >= kXtsHybTt) {
        reg_inc<kBsRegs>();
        if (a.q) {                                   // work queue: tiles from the BACK, the next one claimed a tile ahead
            uint64_t u = q_back(q_post(a.q, 1ull << 32, a.q_zero), a.ntiles);
            while (u != kQNone) {
                const unsigned long long posted = q_post(a.q, 1ull << 32, a.q_zero);
                xts_bitsliced_warp<NR, ENC>(a, lb, u, u + 1);
                u = q_back(posted, a.ntiles);
            }
            return;
        }
        const uint64_t nbs = a.ntiles - a.tt_tiles;
        const uint64_t gw = (uint64_t)blockIdx.x * (kBsThreads / 32) + ((threadIdx.x - kXtsHybTt) >> 5);
        const uint64_t nw = (uint64_t)gridDim.x * (kBsThreads / 32);
        const uint64_t per = (nbs + nw - 1) / nw;
        const uint64_t p0 = gw * per < nbs ? gw * per : nbs;
        const uint64_t p1 = p0 + per < nbs ? p0 + per : nbs;
        xts_bitsliced_warp<NR, ENC>(a, lb, a.tt_tiles + p0, a.tt_tiles + p1);
        return;
    }
    reg_dec<kTtRegs>();

    xts_hybrid_tt_role<NR, ENC, UAES_XTS_ILP, kTtWarps>(a, lb, threadIdx.x >> 5);
}

// ---- the same with NARROW bitsliced warps (uaes_bitslice8.cuh, general form): 8 sectors per pass, lane l, slot t
// <-> block l of sector t.  32 state registers instead of 128: the warps run in 96 registers like the table-driven
// ones (8 of them per SM, no setmaxnreg) and one round is 440 (forward) / 520 (inverse) instructions, so the
// round loop -- also the INVERSE one, whose wide form did not fit the instruction cache -- stays cached.
// MEASURED SLOWER than the wide form (522 vs 612 GiB/s encrypt, 490 vs 558 decrypt, profiles/r2_sweep_xts8.txt): without
// CTR's hoisted rounds a narrow block costs 1.5 x the instructions of a table-driven one.  Selectable (UAES_XTS_NARROW=1),
// parity-tested, off by default.
struct XtsHybridArgs8 {
    XtsSectorArgs x;             // sector_blocks == 32
    uint64_t tt_tiles, ntiles;   // as XtsHybridArgs
    unsigned long long *q;
    uint32_t q_zero;
    BsKeyPlanes8Full bs;
};

#ifndef UAES_XTS8_BS
#define UAES_XTS8_BS 256
#endif
#ifndef UAES_XTS8_DEC_DEFAULT_SHARE
#define UAES_XTS8_DEC_DEFAULT_SHARE 165          // decryption with the narrow co-runner: on (work queue; the value only has to be > 0)
#endif
constexpr int kXts8BsThreads = UAES_XTS8_BS;

template <int NR, bool ENC>
__device__ __forceinline__ void xts_bs8_tile(const XtsHybridArgs8 &a, uint32_t lb, uint64_t tile)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t sec0 = tile * 32;
    const uint64_t left = a.x.nsectors - sec0;
    const int nsec = left < 32 ? (int)left : 32;
    const uint4 *src = a.x.in + sec0 * 32 + lane;
    uint4 *dst = a.x.out + sec0 * 32 + lane;
    if ((int)(lane >> 2) < nsec) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.x.in + sec0 * 32 + 8 * lane));   // the first 8 sectors
    // T_0 of sector sec0 + lane (micro_aes.c:1017-1027)
    const uint64_t sec = a.x.first_sector + sec0 + lane;
    uint32_t e0 = (uint32_t)sec, e1 = (uint32_t)(sec >> 32), e2 = 0, e3 = 0;
    if (ENC) enc_block<NR>(lb, e0, e1, e2, e3, a.x.k2.w);
    else     enc_block_te0<NR, kOffDecTe0>(lb, e0, e1, e2, e3, a.x.k2.w);
    auto tweak_of = [&](int t, uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3) {
        Tweak tw;
        tw.lo = (uint64_t)__shfl_sync(0xffffffffu, e1, t) << 32 | __shfl_sync(0xffffffffu, e0, t);
        tw.hi = (uint64_t)__shfl_sync(0xffffffffu, e3, t) << 32 | __shfl_sync(0xffffffffu, e2, t);
        tweak_words(xts_shl(tw, lane), w0, w1, w2, w3);
    };
#pragma unroll 1
    for (int q = 0; q < 32; q += 8) {
        if (q >= nsec) break;
        if (q + 8 + (int)(lane >> 2) < nsec)                     // the next pass's 4 KiB towards L2 while this one's rounds run
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.x.in + (sec0 + q + 8) * 32 + 8 * lane));
        uint32_t s[32];
        {
            uint4 v[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = q + t < nsec ? ld_stream(src + (q + t) * 32) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                uint32_t w0, w1, w2, w3;
                tweak_of(q + t, w0, w1, w2, w3);
                s[t] = v[t].x ^ w0; s[8 + t] = v[t].y ^ w1; s[16 + t] = v[t].z ^ w2; s[24 + t] = v[t].w ^ w3;
            }
        }
        bs_transpose32(s);
        if (ENC) bs8_encrypt_planes<NR>(s, a.bs); else bs8_decrypt_planes<NR>(s, a.bs);
        bs_transpose32(s);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            uint32_t w0, w1, w2, w3;
            tweak_of(q + t, w0, w1, w2, w3);
            if (q + t < nsec) st_stream(dst + (q + t) * 32, make_uint4(s[t] ^ w0, s[8 + t] ^ w1, s[16 + t] ^ w2, s[24 + t] ^ w3));
        }
    }
}

template <int NR, bool ENC>
__global__ void __launch_bounds__(kXtsTtThreads + kXts8BsThreads, 1) xts_sectors_hybrid8_kernel(const __grid_constant__ XtsHybridArgs8 a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_xts_tables<ENC>(dyn);
    constexpr uint32_t kBsWarps = kXts8BsThreads / 32;
    const uint32_t w = threadIdx.x >> 5;
    if (w < kBsWarps) {                              // the bitsliced warps take the low warp numbers (ctr_queue8_kernel: UAES_Q8_MAP)
        if (a.q) {                                   // work queue: tiles from the BACK, the next one claimed a tile ahead
            uint64_t u = q_back(q_post(a.q, 1ull << 32, a.q_zero), a.ntiles);
            while (u != kQNone) {
                const unsigned long long posted = q_post(a.q, 1ull << 32, a.q_zero);
                xts_bs8_tile<NR, ENC>(a, lb, u);
                u = q_back(posted, a.ntiles);
            }
            return;
        }
        const uint64_t nbs = a.ntiles - a.tt_tiles;
        const uint64_t gw = (uint64_t)blockIdx.x * kBsWarps + w, nw = (uint64_t)gridDim.x * kBsWarps;
        const uint64_t per = (nbs + nw - 1) / nw;
        const uint64_t p0 = gw * per < nbs ? gw * per : nbs;
        const uint64_t p1 = p0 + per < nbs ? p0 + per : nbs;
        for (uint64_t tile = a.tt_tiles + p0; tile < a.tt_tiles + p1; ++tile) xts_bs8_tile<NR, ENC>(a, lb, tile);
        return;
    }
    xts_hybrid_tt_role<NR, ENC, 2, kXtsTtThreads / 32>(a, lb, w - kBsWarps);
}

#ifndef UAES_XTS_NARROW_DEFAULT
#define UAES_XTS_NARROW_DEFAULT 0           // measured: 522 vs 612 GiB/s (encrypt), 490 vs 558 (decrypt against no co-runner): profiles/r2_sweep_xts8.txt
#endif

template <int NR, bool ENC>
static cudaError_t launch_xts_hybrid8_nr(const XtsSectorArgs &x, uint64_t bs_tiles, cudaStream_t st)
{
    cudaError_t e = opt_in_smem(xts_sectors_hybrid8_kernel<NR, ENC>);
    if (e != cudaSuccess) return e;
    static thread_local XtsHybridArgs8 a;
    a.x = x;
    a.ntiles = (x.nsectors + 31) / 32;
    a.tt_tiles = a.ntiles - bs_tiles;
    a.q = nullptr; a.q_zero = 0;
    if (env_int("UAES_XTS_QUEUE", kXtsQueueDefault)) {                 // dynamic split, both directions (the static share is ignored)
        if ((e = q_slot(st, &a.q)) != cudaSuccess) return e;
    }
    bs8_make_key_planes_full(x.k1.w, NR, &a.bs);
    const uint64_t need = (a.ntiles + 15) / 16, sms = (uint64_t)sm_count();
    xts_sectors_hybrid8_kernel<NR, ENC><<<(unsigned)(need < sms ? need : sms), kXtsTtThreads + kXts8BsThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

template <int NR, bool ENC>
static cudaError_t launch_xts_hybrid_nr(const XtsSectorArgs &x, uint64_t bs_tiles, cudaStream_t st)
{
    cudaError_t e = opt_in_smem(xts_sectors_hybrid_kernel<NR, ENC>);
    if (e != cudaSuccess) return e;
    static thread_local XtsHybridArgs a;                 // 8 KB of planes: off the stack, one per calling thread
    a.x = x;
    a.ntiles = (x.nsectors + 31) / 32;
    a.tt_tiles = a.ntiles - bs_tiles;
    a.q = nullptr; a.q_zero = 0;
    if (env_int(ENC ? "UAES_XTS_QUEUE" : "UAES_XTS_DEC_QUEUE", kXtsQueueDefault)) {   // dynamic split (the static share is ignored)
        if ((e = q_slot(st, &a.q)) != cudaSuccess) return e;
    }
    bs_make_key_planes_full(x.k1.w, NR, &a.bs);
    const uint64_t need = (a.ntiles + 15) / 16, sms = (uint64_t)sm_count();
    xts_sectors_hybrid_kernel<NR, ENC><<<(unsigned)(need < sms ? need : sms), kXtsHybTt + kBsThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

template <int NR>
static cudaError_t launch_xts_unit_hybrid_nr(const XtsUnitArgs &u, uint64_t bs_blocks, cudaStream_t st)
{
    cudaError_t e = opt_in_smem(xts_unit_hybrid_kernel<NR>);
    if (e != cudaSuccess) return e;
    static thread_local XtsUnitHybridArgs a;                 // 8 KB of planes: off the stack, one per calling thread
    a.u = u;
    a.tt_blocks = (u.nblocks - bs_blocks) & ~1023ull;
    bs_make_key_planes_full(u.k1.w, NR, &a.bs);
    const uint64_t need = (u.nblocks + 32 * 16 - 1) / (32 * 16), sms = (uint64_t)sm_count();
    xts_unit_hybrid_kernel<NR><<<(unsigned)(need < sms ? need : sms), kXtsUnitTt + kBsThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

template <int NR, bool ENC>
static cudaError_t launch_xts_unit_nr(const XtsUnitArgs &a, cudaStream_t st)
{
    if (ENC) {
        // A large unit with the co-runner.  Next to 12 table-driven warps with two rows in flight it lost (574 GiB/s
        // without, 535 with, profiles/r1_xts_hybrid_sweep.txt); next to 16 warps with one row in flight at 64 registers it
        // pays: 574 -> 619 / 640 / 618 GiB/s at 100 / 140 / 180 per 1024 (AES-256, profiles/r2_sweep_xtsunit_ilp.txt).
        ctr_tuning_init();
        const int share = g_ctr_share != kCtrDefaultShare ? g_ctr_share : env_int("UAES_XTS_UNIT_BS_PERMILLE", kXtsUnitDefaultShare);
        if (g_ctr_share > 0 && share > 0 && (long long)a.nblocks >= g_ctr_bs_min && a.nblocks >= 2048)
            return launch_xts_unit_hybrid_nr<NR>(a, a.nblocks / 1024 * (uint64_t)share, st);
    }
    cudaError_t e = opt_in_smem(xts_unit_kernel<NR, ENC>);
    if (e != cudaSuccess) return e;
    // at least 8 rows per warp so that the jump-ahead amortises
    xts_unit_kernel<NR, ENC><<<grid_for((a.nblocks + 255) / 256), kThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace uaes

extern "C" int uaes_launch_xts_sectors(const uaes_keysched *ks1, const uaes_keysched *ks2, int encrypt,
                                       u64 first_sector, u64 sector_blocks, u64 nsectors,
                                       const void *in, void *out, void *stream)
{
    using namespace uaes;
    if (nsectors == 0 || sector_blocks == 0) return 0;
    XtsSectorArgs a;
    a.k1 = *ks1; a.k2 = *ks2;
    a.first_sector = first_sector; a.sector_blocks = sector_blocks; a.nsectors = nsectors;
    a.in = (const uint4 *)in; a.out = (uint4 *)out;
    cudaStream_t st = (cudaStream_t)stream;
    // 512-byte sectors, enough of them: table-driven warps + bitsliced co-runner
    // (threshold and on/off are the CTR kernel's knobs, uaes_ctr_tuning; the share is XTS's own:
    // measured 549 / 565 / 577 / 584 / 566 / 519 GiB/s at 0 / 60 / 100 / 140 / 180 / 220 per 1024 for
    // AES-256, profiles/r1_xts_hybrid_sweep.txt; a non-default CTR share, as the tests set, wins)
    ctr_tuning_init();
    if (sector_blocks == 32 && g_ctr_share > 0 && (long long)(nsectors * 32) >= g_ctr_bs_min) {
        const uint64_t ntiles = (nsectors + 31) / 32;
        // Decryption: the bitsliced inverse cipher as co-runner did not pay next to 12 table-driven warps with two
        // sectors in flight (558 -> 523 / 480 GiB/s at 30 / 100 per 1024); next to 16 warps with one sector in flight
        // it does: 557 -> 589 / 583 at 60 / 100 per 1024 (static split), profiles/r2_sweep_xtsdec_ilp.txt.  On by default.
        const bool narrow = env_int("UAES_XTS_NARROW", UAES_XTS_NARROW_DEFAULT) != 0;
        const int dflt = encrypt ? env_int("UAES_XTS_BS_PERMILLE", kXtsDefaultShare)
                                 : env_int("UAES_XTS_DEC_BS_PERMILLE", narrow ? UAES_XTS8_DEC_DEFAULT_SHARE : kXtsDecDefaultShare);
        const int share = g_ctr_share != kCtrDefaultShare ? g_ctr_share : dflt;
        const uint64_t bs_tiles = ntiles * (uint64_t)share / 1024;
        if (bs_tiles > 0 && narrow) {
            switch (ks1->rounds * 2 + (encrypt ? 1 : 0)) {
            case 21: return (int)launch_xts_hybrid8_nr<10, true>(a, bs_tiles, st);
            case 20: return (int)launch_xts_hybrid8_nr<10, false>(a, bs_tiles, st);
            case 25: return (int)launch_xts_hybrid8_nr<12, true>(a, bs_tiles, st);
            case 24: return (int)launch_xts_hybrid8_nr<12, false>(a, bs_tiles, st);
            case 29: return (int)launch_xts_hybrid8_nr<14, true>(a, bs_tiles, st);
            case 28: return (int)launch_xts_hybrid8_nr<14, false>(a, bs_tiles, st);
            }
        }
        if (bs_tiles > 0) {
            switch (ks1->rounds * 2 + (encrypt ? 1 : 0)) {
            case 21: return (int)launch_xts_hybrid_nr<10, true>(a, bs_tiles, st);
            case 20: return (int)launch_xts_hybrid_nr<10, false>(a, bs_tiles, st);
            case 25: return (int)launch_xts_hybrid_nr<12, true>(a, bs_tiles, st);
            case 24: return (int)launch_xts_hybrid_nr<12, false>(a, bs_tiles, st);
            case 29: return (int)launch_xts_hybrid_nr<14, true>(a, bs_tiles, st);
            case 28: return (int)launch_xts_hybrid_nr<14, false>(a, bs_tiles, st);
            }
        }
    }
    switch (ks1->rounds * 2 + (encrypt ? 1 : 0)) {
    case 21: return (int)launch_xts_sectors_nr<10, true>(a, st);
    case 20: return (int)launch_xts_sectors_nr<10, false>(a, st);
    case 25: return (int)launch_xts_sectors_nr<12, true>(a, st);
    case 24: return (int)launch_xts_sectors_nr<12, false>(a, st);
    case 29: return (int)launch_xts_sectors_nr<14, true>(a, st);
    case 28: return (int)launch_xts_sectors_nr<14, false>(a, st);
    }
    return (int)cudaErrorInvalidValue;
}

extern "C" int uaes_launch_xts_unit(const uaes_keysched *ks1, const uaes_keysched *ks1e,
                                    const uaes_keysched *ks2, int encrypt, const unsigned char tweak[16],
                                    u64 first_block, const void *in, void *out, u64 len, void *stream)
{
    using namespace uaes;
    if (len < 16) return (int)cudaErrorInvalidValue;
    XtsUnitArgs a;
    a.k1 = *ks1; a.k1e = *ks1e; a.k2 = *ks2;
    for (int c = 0; c < 4; ++c)
        a.tweak[c] = (uint32_t)tweak[4 * c] | (uint32_t)tweak[4 * c + 1] << 8 | (uint32_t)tweak[4 * c + 2] << 16 | (uint32_t)tweak[4 * c + 3] << 24;
    a.in = (const uint4 *)in; a.out = (uint4 *)out;
    a.tail = (uint32_t)(len % 16);
    a.nblocks = len / 16 - (a.tail ? 1 : 0);
    a.first_block = first_block;
    cudaStream_t st = (cudaStream_t)stream;
    switch (ks1->rounds * 2 + (encrypt ? 1 : 0)) {
    case 21: return (int)launch_xts_unit_nr<10, true>(a, st);
    case 20: return (int)launch_xts_unit_nr<10, false>(a, st);
    case 25: return (int)launch_xts_unit_nr<12, true>(a, st);
    case 24: return (int)launch_xts_unit_nr<12, false>(a, st);
    case 29: return (int)launch_xts_unit_nr<14, true>(a, st);
    case 28: return (int)launch_xts_unit_nr<14, false>(a, st);
    }
    return (int)cudaErrorInvalidValue;
}
