// uaes_kernels.cu -- sm_100a kernels for the AES bulk path and their C launchers.
//
// Kernel shape shared by every mode ("Kernels"):
//   * persistent grid: one 768- or 1024-thread CTA per SM (148 on B200), 227 KB of dynamic shared
//     memory holding the lane-replicated T-tables (uaes_tables.cuh);
//   * one 16-byte block per thread per step, a warp covers 32 consecutive blocks = 512
//     contiguous bytes, moved with one 128-bit load and one 128-bit store per thread;
//   * the next step's input is requested before the current step's rounds start, so the warps
//     keep 12-16 KB of loads in flight per SM;
//   * round keys are kernel arguments (constant bank), no per-launch symbol copies.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <atomic>

#include "uaes_core.cuh"
#include "uaes_gf128.cuh"
#include "uaes_bitslice.cuh"
#include "uaes_bitslice8.cuh"

namespace uaes {

constexpr int kThreads = 1024;
constexpr int kWarpsPerCta = kThreads / 32;
constexpr uint32_t kDynSmem = 227 * 1024;          // everything an SM has; tables are aligned inside

static std::atomic<unsigned long long> g_launches{0};     // host threads of several devices launch concurrently

// ---------------------------------------------------------------- small device helpers

// cache operators of the bulk loads/stores (every byte is touched once); overridable for tuning
#ifndef UAES_LD
#define UAES_LD "ld.global.cs.v4.u32"
#endif
#ifndef UAES_ST
#define UAES_ST "st.global.cs.v4.u32"
#endif

__device__ __forceinline__ uint4 ld_stream(const uint4 *p)
{
    uint4 v;
    asm volatile(UAES_LD " {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ void st_stream(uint4 *p, uint4 v)
{
    asm volatile(UAES_ST " [%0], {%1,%2,%3,%4};"
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t dyn_smem_size()
{
    uint32_t v;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(v));
    return v;
}

// Builds the tables and returns lanebase = table base + lane*4.  Traps if the aligned tables do
// not fit (cannot happen with kDynSmem on sm_100a, but a silent overrun would corrupt results).
template <bool ENC>
__device__ __forceinline__ uint32_t setup_tables(const void *dyn)
{
    const uint32_t base = ENC ? align_table_base(dyn) : dec_table_base(dyn);
    if (base + (ENC ? kEncTableBytes : kDecTableBytes) > smem_u32(dyn) + dyn_smem_size()) __trap();
    if (ENC) init_enc_tables(base); else init_dec_tables(base);
    __syncthreads();
    uint32_t lanebase = base + (threadIdx.x & 31) * 4;
    // lookups are plain (non-volatile) asm so the compiler may schedule them freely; this
    // barrier keeps them from being hoisted above the table fill
    asm volatile("" : "+r"(lanebase)::"memory");
    return lanebase;
}

// counter-block words 2 and 3 (bytes 8..15) for 56-bit counter value v (micro_aes.c:421-427:
// big-endian in bytes 9..15)
__device__ __forceinline__ void ctr_words(uint32_t b8, uint64_t v, uint32_t &w2, uint32_t &w3)
{
    w2 = b8 | __byte_perm((uint32_t)(v >> 32) & 0x00ffffffu, 0, 0x0123);
    w3 = __byte_perm((uint32_t)v, 0, 0x0123);
}

constexpr uint64_t kMask56 = (1ull << 56) - 1;

// ---------------------------------------------------------------- CTR (micro_aes.c:919-950)

struct CtrArgsBase {
    uaes_keysched ks;
    uint32_t w0, w1, b8;
    uint64_t v0;                 // counter of block 0
    const uint4 *in;
    uint4 *out;
    uint64_t nblocks;            // full blocks
    uint32_t tail;               // len % 16
    // split between the table-driven warps and the bitsliced ALU co-runner warps (below)
    uint64_t tt_blocks;          // blocks [0, tt_blocks) belong to the table-driven warps
    uint64_t bs_u0;              // v0 + tt_blocks: first counter (not reduced mod 2^56) of the
                                 // bitsliced range, a multiple of 1024 unless bs_passes == 0
    uint64_t bs_passes;          // 1024-counter passes covering blocks [tt_blocks, nblocks)
    // work-queue form (ctr_queue_kernel): the counter range in units of 2^q_shift blocks, handed out
    // from the front to the table-driven warps and from the back to the bitsliced warps
    unsigned long long *q;       // device: [0] front | back << 32, [1] units done by table-driven warps, [2] by bitsliced warps
    uint64_t q_u0;               // counter (not reduced mod 2^56) where unit 0 starts: v0 rounded down to a unit
    uint64_t q_units;            // units covering [v0, v0 + nblocks)
    uint32_t q_bs_on;            // 0: the bitsliced warps take no work (short calls)
    uint32_t q_zero;             // 0 (see q_post)
    uint32_t q_shift;            // log2(blocks per unit): 11, or 10 for calls short enough that the last unit shows
};
struct CtrArgs : CtrArgsBase {
    BsKeyPlanes bs;              // wide bitsliced form (uaes_bitslice.cuh): 32 blocks per thread
};
struct CtrArgs8 : CtrArgsBase {
    BsKeyPlanes8 bs8;            // narrow form (uaes_bitslice8.cuh): 8 blocks per thread
};

// Work unit = a "group": the 256 counter values that share bytes 0..14 of the counter block.
// Inside a group only byte 15 changes, and the cipher's first two rounds factor accordingly:
//
//   round 1: column 0 = K0 ^ Te3[S-box input of byte 15]; columns 1..3 (C1..C3) do not see byte 15.
//            K0 depends on counter bytes 0, 5, 10 only -> constant for 2^40 consecutive blocks.
//   round 2: column j = D_j ^ Te_x[one byte of column 0].  The Te_x term is a function of byte 15
//            (and K0) alone, so a lane that always serves the same byte-15 values keeps those four
//            words per value in REGISTERS for the whole launch (U[it][0..3] below).
//            D_j = E_j ^ Te_y[a byte of C1]; C1 follows counter byte 14 (one lookup per group), the
//            E_j follow bytes 12..13 (recomputed every 256 groups).
//
// So rounds 1-2 cost 5 lookups per GROUP HALF (4 rows) instead of 32 per block, and every block
// pays only the 16 x (NR-2) lookups of rounds 3..NR -- the shared-memory pipe is the roof
// (profiles/), so lookups are what is worth saving.  A warp serves one half of a group (rows
// it = 0..3: byte 15 = 128*half + 32*it + lane) and walks a CONTIGUOUS run of groups so that the
// slow-changing constants really are constant; its partner warp serves the other half.
// ---- the ALU co-runner: one warpgroup (4 warps, one per scheduler) of bitsliced AES -------------
// The table-driven warps keep the shared-memory pipe at its roof and, with their lookup addresses
// built on the FMA pipe, use less than half of the ALU pipe; these four warps turn the rest into
// keystream with no lookups at all (uaes_bitslice.cuh).  Register budgets differ by 2x, so the two
// roles re-balance the CTA's register file with setmaxnreg right after the table fill.
constexpr int kBsThreads = 128;
#ifndef UAES_HYB_TT_REGS
#define UAES_HYB_TT_REGS 104
#endif
#ifndef UAES_BS_LOAD_BATCH
#define UAES_BS_LOAD_BATCH 8
#endif
constexpr int kBsLoadBatch = UAES_BS_LOAD_BATCH;   // rows a general-form bitsliced warp loads at a time
constexpr int kHybridTtRegs = UAES_HYB_TT_REGS;  // register budget of the table-driven threads in the ECB / XTS / OCB kernels with a co-runner

template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// one pass = 1024 consecutive counters starting at u0 (a multiple of 1024, not reduced mod 2^56);
// tag16 remembers for which counter bytes <= 13 the uniform masks `um` were built
template <int NR, int BATCH>
__device__ __forceinline__ void ctr_bs_pass(const CtrArgs &a, uint32_t lb, uint32_t *um, uint64_t u0, uint64_t &tag16)
{
    const uint32_t *rk = a.ks.w;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t vc = u0 & kMask56;
    {   // pull this pass's 16 KiB of input towards L2 while the rounds run (4 lines per lane)
        const uint64_t kb = u0 - a.v0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint64_t k = kb + (uint64_t)(lane * 4 + i) * 8;
            if (k < a.nblocks) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.in + k));
        }
    }
    if ((vc >> 16) != tag16) {              // counter bytes <= 13 changed: every 64 passes
        tag16 = vc >> 16;
        uint32_t w2, w3, uw[6];
        ctr_words(a.b8, vc, w2, w3);
        bs_uniform_words([&](int t, uint32_t x) { return lut_index(lb, t == 0 ? kOffT0 : t == 1 ? kOffT1 : t == 2 ? kOffT2 : kOffT3, x); },
                         a.w0 ^ rk[0], a.w1 ^ rk[1], w2 ^ rk[2], w3 ^ rk[3], rk, uw);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 6; ++j) um[32 * j + lane] = bs_mask(uw[j], (int)lane);
        __syncwarp();
    }
    uint32_t s[128];
    bs_first_rounds(s, lane, (uint32_t)(vc >> 8) & 0xfcu, a.bs.k0, um);
    // all rounds in ONE loop body (the last one skips MixColumns): a second copy of the S-box layer
    // for the last round cost 1.2 % through the instruction cache (1008 -> 1020 GiB/s)
#pragma unroll 1
    for (int r = 3; r <= NR; ++r) bs_round_or_last(s, a.bs.k[r - 3], r == NR);
    // XOR with the data: slot t of all lanes = one coalesced 512-byte row.  The loads are
    // software-pipelined one batch ahead (the first batch goes out before the transposes).
    const int64_t k0 = (int64_t)(u0 - a.v0) + lane;          // block index of slot 0 (may be < 0 in the first unit)
    auto load_batch = [&](int t0, uint4 (&x)[BATCH]) {
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
            const int64_t k = k0 + 32 * (t0 + i);
            x[i] = (uint64_t)k < a.nblocks ? ld_stream(a.in + k) : make_uint4(0, 0, 0, 0);
        }
    };
    uint4 x[BATCH], y[BATCH];
    load_batch(0, x);
#pragma unroll
    for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
#pragma unroll
    for (int t0 = 0; t0 < 32; t0 += BATCH) {
        if (t0 + BATCH < 32) load_batch(t0 + BATCH, y);
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
            const int64_t k = k0 + 32 * (t0 + i);
            const int t = t0 + i;
            x[i].x ^= s[t]; x[i].y ^= s[32 + t]; x[i].z ^= s[64 + t]; x[i].w ^= s[96 + t];
            if ((uint64_t)k < a.nblocks) st_stream(a.out + k, x[i]);
            x[i] = y[i];
        }
    }
}

template <int NR, int BATCH>
__device__ __forceinline__ void ctr_bitsliced_warp(const CtrArgs &a, uint32_t lb, uint32_t *um,
                                                   uint64_t p0, uint64_t p1)
{
    uint64_t tag16 = ~0ull;
    for (uint64_t p = p0; p < p1; ++p) ctr_bs_pass<NR, BATCH>(a, lb, um, a.bs_u0 + (p << 10), tag16);
}

template <int NR, int kCtrThreads, bool BS, int ILP>
__global__ void __launch_bounds__(kCtrThreads + (BS ? kBsThreads : 0), 1) ctr_kernel(const __grid_constant__ CtrArgs a)
{
    constexpr int kCtrWarps = kCtrThreads / 32;
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_tables<true>(dyn);
    const uint32_t *rk = a.ks.w;
    const uint32_t lane = threadIdx.x & 31;

    if (BS) {
        // launch allocation is 65536 / threads rounded down to 8; hand the table warps' surplus to
        // the bitsliced warpgroup (whose 128-plane state needs it)
        constexpr int kLaunchRegs = (65536 / (kCtrThreads + kBsThreads)) / 8 * 8;
        constexpr int kTtRegs = ILP == 1 ? 80 : 104;       // 96 and 112 measured worse (profiles/r1_ctr_hybrid_sweep.txt)
        constexpr int kBsRegs0 = kLaunchRegs + (kLaunchRegs - kTtRegs) * kCtrThreads / kBsThreads;
        constexpr int kBsRegs = (kBsRegs0 > 232 ? 232 : kBsRegs0) / 8 * 8;
        if (threadIdx.x >= kCtrThreads) {
            reg_inc<kBsRegs>();
            const uint32_t tbase = align_table_base(dyn);
            const uint32_t bw = (threadIdx.x - kCtrThreads) >> 5;
            const uint32_t off = tbase + kEncTableBytes + bw * (kBsUniformMasks * 4) - smem_u32(dyn);
            if (off + kBsUniformMasks * 4 > dyn_smem_size()) __trap();
            const uint64_t gw = (uint64_t)blockIdx.x * (kBsThreads / 32) + bw;
            const uint64_t nw = (uint64_t)gridDim.x * (kBsThreads / 32);
            const uint64_t per = (a.bs_passes + nw - 1) / nw;
            const uint64_t p0 = gw * per < a.bs_passes ? gw * per : a.bs_passes;
            const uint64_t p1 = p0 + per < a.bs_passes ? p0 + per : a.bs_passes;
            ctr_bitsliced_warp<NR, 4>(a, lb, (uint32_t *)(dyn + off), p0, p1);
            return;
        }
        reg_dec<kTtRegs>();
    }

    const uint64_t tt_blocks = BS ? a.tt_blocks : a.nblocks;
    const uint32_t lowoff = (uint32_t)a.v0 & 255u;
    const uint64_t g0 = a.v0 >> 8;
    const uint64_t ngroups = (lowoff + tt_blocks + 255) >> 8;
    const uint64_t wg = (uint64_t)blockIdx.x * kCtrWarps + (threadIdx.x >> 5);
    const uint32_t half = (uint32_t)wg & 1u;
    const uint64_t npairs = (uint64_t)gridDim.x * (kCtrWarps / 2);
    const uint64_t per = (ngroups + npairs - 1) / npairs;
    const uint64_t j0 = (wg >> 1) * per;
    const uint64_t j1 = j0 + per < ngroups ? j0 + per : ngroups;
    const uint32_t b15 = half * 128 + lane;                   // + 32 * it

    // block index handled by this lane in row `it` of group j; valid iff 0 <= k < nblocks
    auto kof = [&](uint64_t grp, int it) -> int64_t {
        return (int64_t)(grp << 8) + (int64_t)(b15 + 32 * it) - (int64_t)lowoff;
    };
    auto fetch = [&](uint64_t grp, int it) -> uint4 {
        const int64_t k = kof(grp, it);
        if (grp < j1 && k >= 0 && (uint64_t)k < tt_blocks) return ld_stream(a.in + k);
        return make_uint4(0, 0, 0, 0);
    };

    uint64_t tag40 = ~0ull, tag16 = ~0ull;
    uint32_t U[4][4], Cp1 = 0, E0 = 0, E1 = 0, E2 = 0, E3 = 0;
    const uint32_t s0 = a.w0 ^ rk[0], s1 = a.w1 ^ rk[1];

    uint4 cur[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) cur[i] = fetch(j0, i);
    for (uint64_t j = j0; j < j1; ++j) {
        const uint64_t vg = ((g0 + j) << 8) & kMask56;
        uint32_t w2, w3;
        ctr_words(a.b8, vg, w2, w3);
        const uint32_t s2 = w2 ^ rk[2], s3 = w3 ^ rk[3];         // byte 15 of the counter is 0 here
        if ((vg >> 40) != tag40) {                               // once per launch in practice
            tag40 = vg >> 40;
            const uint32_t K0 = lut<0, kOffT0>(lb, s0) ^ lut<1, kOffT1>(lb, s1) ^ lut<2, kOffT2>(lb, s2) ^ rk[4];
            Cp1 = lut<0, kOffT0>(lb, s1) ^ lut<1, kOffT1>(lb, s2) ^ lut<3, kOffT3>(lb, s0) ^ rk[5];
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const uint32_t c0 = K0 ^ lut<3, kOffT3>(lb, s3 ^ ((b15 + 32 * it) << 24));
                U[it][0] = lut<0, kOffT0>(lb, c0); U[it][1] = lut<3, kOffT3>(lb, c0);
                U[it][2] = lut<2, kOffT2>(lb, c0); U[it][3] = lut<1, kOffT1>(lb, c0);
            }
            tag16 = ~0ull;
        }
        if ((vg >> 16) != tag16) {                               // every 256 groups
            tag16 = vg >> 16;
            const uint32_t C2 = lut<0, kOffT0>(lb, s2) ^ lut<1, kOffT1>(lb, s3) ^ lut<2, kOffT2>(lb, s0) ^ lut<3, kOffT3>(lb, s1) ^ rk[6];
            const uint32_t C3 = lut<0, kOffT0>(lb, s3) ^ lut<1, kOffT1>(lb, s0) ^ lut<2, kOffT2>(lb, s1) ^ lut<3, kOffT3>(lb, s2) ^ rk[7];
            E0 = lut<2, kOffT2>(lb, C2) ^ lut<3, kOffT3>(lb, C3) ^ rk[8];
            E1 = lut<1, kOffT1>(lb, C2) ^ lut<2, kOffT2>(lb, C3) ^ rk[9];
            E2 = lut<0, kOffT0>(lb, C2) ^ lut<1, kOffT1>(lb, C3) ^ rk[10];
            E3 = lut<0, kOffT0>(lb, C3) ^ lut<3, kOffT3>(lb, C2) ^ rk[11];
        }
        // per group: counter byte 14 enters through column 1 of round 1
        const uint32_t C1 = Cp1 ^ lut<2, kOffT2>(lb, s3);
        const uint32_t D0 = E0 ^ lut<1, kOffT1>(lb, C1), D1 = E1 ^ lut<0, kOffT0>(lb, C1);
        const uint32_t D2 = E2 ^ lut<3, kOffT3>(lb, C1), D3 = E3 ^ lut<2, kOffT2>(lb, C1);

#pragma unroll
        for (int it = 0; it < 4; it += ILP) {
            uint4 nxt[ILP];
            uint32_t t[ILP][4];
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                const int n = it + ILP + i;
                nxt[i] = n < 4 ? fetch(j, n) : fetch(j + 1, n - 4);
                t[i][0] = D0 ^ U[it + i][0]; t[i][1] = D1 ^ U[it + i][1];
                t[i][2] = D2 ^ U[it + i][2]; t[i][3] = D3 ^ U[it + i][3];
            }
            enc_finish_n<NR, 3, ILP>(lb, t, rk, cur);
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                const int64_t k = kof(j, it + i);
                if (k >= 0 && (uint64_t)k < tt_blocks) st_stream(a.out + k, make_uint4(t[i][0], t[i][1], t[i][2], t[i][3]));
                cur[i] = nxt[i];
            }
        }
    }

    // ragged tail: Y[0..n) = E(ctr)[0..n) ^ X[0..n)  (mixThenXor, micro_aes.c:534-544)
    if (a.tail && blockIdx.x == 0 && threadIdx.x == 0) {
        uint32_t w2, w3;
        ctr_words(a.b8, (a.v0 + a.nblocks) & kMask56, w2, w3);
        uint32_t s0 = a.w0, s1 = a.w1, s2 = w2, s3 = w3;
        enc_block<NR>(lb, s0, s1, s2, s3, rk);
        const uint32_t ksw[4] = {s0, s1, s2, s3};
        const uint8_t *x = (const uint8_t *)(a.in + a.nblocks);
        uint8_t *y = (uint8_t *)(a.out + a.nblocks);
        for (uint32_t i = 0; i < a.tail; ++i) y[i] = x[i] ^ (uint8_t)(ksw[i >> 2] >> (8 * (i & 3)));
    }
}

// ---- CTR with a two-ended work queue ---------------------------------------------------------------
// Same two kinds of warps as ctr_kernel, but no static split: the counter range is cut into units of
// 2^q_shift blocks; table-driven warps claim units from the FRONT, bitsliced warps from the BACK, through
// one 64-bit atomic add on a packed (front, back) word.  A claim sees both counts at its own place in
// the atomic order, so "front + back < units" decides exactly and without a retry loop whether the
// unit is still free: the two kinds of warps meet wherever their actual speeds on THIS GPU put the
// border, nobody waits for the other kind (the static 195/1024 split was tuned on one box; the pool
// spreads 982..1026 GiB/s), and short calls need no special case.
// A table-driven warp serves both halves of its unit one after the other (rows 0..3, then rows 4..7 of
// every group), re-deriving the 16 register-resident round-2 words when the half changes; its next
// unit is claimed a unit ahead so that neither the atomic nor the first loads of that unit are
// waited for.
#ifndef UAES_BS_BATCH
#define UAES_BS_BATCH 2                          // rows of a bitsliced pass loaded ahead of the XOR / store (4: 16 more
                                                 // registers, spills: 996 -> 1036 GiB/s going from 4 to 2, profiles/r2_ctr_queue_sweep*.txt)
#endif
#ifndef UAES_Q_UNIT_SHIFT
#define UAES_Q_UNIT_SHIFT 11                     // 2048 blocks = 32 KiB = 8 groups = 2 bitsliced passes
#endif
constexpr int kQUnitShift = UAES_Q_UNIT_SHIFT;
constexpr uint64_t kQNone = ~0ull;

// A claim is split in two so that nobody waits for the atomic: q_post() issues it from lane 0 (the
// result is not touched), q_front() / q_back() broadcast and decode it a unit later.
__device__ __forceinline__ unsigned long long q_post(unsigned long long *q, unsigned long long inc, uint32_t zero)
{
    // ptxas turns an atomic on a warp-uniform address into its aggregated form (VOTE, POPC, ATOMG,
    // SHFL), whose shuffle reads the result straight away.  An address it cannot prove uniform
    // (q + 0 * %laneid, the zero being a kernel argument) keeps the plain ATOMG, issued and left alone.
    unsigned long long old = 0;
    if ((threadIdx.x & 31) == 0) {
        uint32_t lid;
        asm volatile("mov.u32 %0, %%laneid;" : "=r"(lid));
        asm volatile("atom.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(q + zero * lid), "l"(inc) : "memory");
    }
    return old;
}

// front claim (table-driven): the unit or kQNone; all lanes get the same answer
__device__ __forceinline__ uint64_t q_front(unsigned long long posted, uint64_t units)
{
    const unsigned long long old = __shfl_sync(0xffffffffu, posted, 0);
    const uint64_t f = old & 0xffffffffull, b = old >> 32;
    return f + b < units ? f : kQNone;
}

__device__ __forceinline__ uint64_t q_back(unsigned long long posted, uint64_t units)
{
    const unsigned long long old = __shfl_sync(0xffffffffu, posted, 0);
    const uint64_t f = old & 0xffffffffull, b = old >> 32;
    return f + b < units ? units - 1 - b : kQNone;
}

#ifndef UAES_TT_L2_PREFETCH
#define UAES_TT_L2_PREFETCH 0
#endif
#ifndef UAES_TT_ROLL_ROWS
#define UAES_TT_ROLL_ROWS 0                     // 1: one copy of the round code for both row pairs of a group; measured -13 % (profiles/r2_sweep_q8.txt)
#endif
// the table-driven role of the work-queue kernels: claims units from the front until none is left,
// then (one thread of the grid) the ragged tail
template <int NR, int ILP>
__device__ __forceinline__ void ctr_queue_table_role(const CtrArgsBase &a, uint32_t lb, uint32_t tail_thread = 0)
{
    const uint32_t *rk = a.ks.w;
    const uint32_t lane = threadIdx.x & 31;
    // absolute group index G = counter >> 8 (counter space not reduced mod 2^56); block index of
    // (G, row r = 4 * half + it, lane) = 256 G + 32 r + lane - v0
    const uint32_t kGroupsPerUnit = 1u << (a.q_shift - 8);
    const uint64_t Gfirst = a.q_u0 >> 8;
    auto kof = [&](uint64_t G, uint32_t r) -> int64_t {
        return (int64_t)((G << 8) + 32 * r + lane) - (int64_t)a.v0;
    };
    auto fetch = [&](bool ok, uint64_t G, uint32_t r) -> uint4 {
        const int64_t k = kof(G, r);
        if (ok && k >= 0 && (uint64_t)k < a.nblocks) return ld_stream(a.in + k);
        return make_uint4(0, 0, 0, 0);
    };

    uint64_t tag40 = ~0ull, tag16 = ~0ull, done = 0;
    uint32_t U[4][4], K0 = 0, Cp1 = 0, E0 = 0, E1 = 0, E2 = 0, E3 = 0;
    const uint32_t s0 = a.w0 ^ rk[0], s1 = a.w1 ^ rk[1];

    uint64_t ucur = q_front(q_post(a.q, 1ull, a.q_zero), a.q_units);
    uint4 cur[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) cur[i] = fetch(ucur != kQNone, Gfirst + ucur * kGroupsPerUnit, (uint32_t)i);
    while (ucur != kQNone) {
        // the next unit is claimed a unit ahead: the atomic is issued here, its answer is read when the
        // second half starts, the first loads of that unit go out with the last rows of this one
        const unsigned long long posted = q_post(a.q, 1ull, a.q_zero);
        uint64_t unext = kQNone;
        const uint64_t Gu = Gfirst + ucur * kGroupsPerUnit;
#pragma unroll 1
        for (uint32_t half = 0; half < 2; ++half) {
            const uint32_t b15 = half * 128 + lane;                   // + 32 * it
            bool fresh = true;                                        // U belongs to (K0, half)
            if (half) unext = q_front(posted, a.q_units);
#pragma unroll 1
            for (uint32_t jj = 0; jj < kGroupsPerUnit; ++jj) {
                const uint64_t G = Gu + jj;
                const uint64_t vg = (G << 8) & kMask56;
                uint32_t w2, w3;
                ctr_words(a.b8, vg, w2, w3);
                const uint32_t s2 = w2 ^ rk[2], s3 = w3 ^ rk[3];     // byte 15 of the counter is 0 here
                if ((vg >> 40) != tag40) {                            // once per launch in practice
                    tag40 = vg >> 40;
                    K0 = lut<0, kOffT0>(lb, s0) ^ lut<1, kOffT1>(lb, s1) ^ lut<2, kOffT2>(lb, s2) ^ rk[4];
                    Cp1 = lut<0, kOffT0>(lb, s1) ^ lut<1, kOffT1>(lb, s2) ^ lut<3, kOffT3>(lb, s0) ^ rk[5];
                    tag16 = ~0ull;
                    fresh = true;
                }
                if (fresh) {                                          // per unit half: 20 lookups for 8192
                    fresh = false;
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const uint32_t c0 = K0 ^ lut<3, kOffT3>(lb, s3 ^ ((b15 + 32 * it) << 24));
                        U[it][0] = lut<0, kOffT0>(lb, c0); U[it][1] = lut<3, kOffT3>(lb, c0);
                        U[it][2] = lut<2, kOffT2>(lb, c0); U[it][3] = lut<1, kOffT1>(lb, c0);
                    }
                }
                if ((vg >> 16) != tag16) {                            // every 256 groups, and on every new unit
                    tag16 = vg >> 16;
                    const uint32_t C2 = lut<0, kOffT0>(lb, s2) ^ lut<1, kOffT1>(lb, s3) ^ lut<2, kOffT2>(lb, s0) ^ lut<3, kOffT3>(lb, s1) ^ rk[6];
                    const uint32_t C3 = lut<0, kOffT0>(lb, s3) ^ lut<1, kOffT1>(lb, s0) ^ lut<2, kOffT2>(lb, s1) ^ lut<3, kOffT3>(lb, s2) ^ rk[7];
                    E0 = lut<2, kOffT2>(lb, C2) ^ lut<3, kOffT3>(lb, C3) ^ rk[8];
                    E1 = lut<1, kOffT1>(lb, C2) ^ lut<2, kOffT2>(lb, C3) ^ rk[9];
                    E2 = lut<0, kOffT0>(lb, C2) ^ lut<1, kOffT1>(lb, C3) ^ rk[10];
                    E3 = lut<0, kOffT0>(lb, C3) ^ lut<3, kOffT3>(lb, C2) ^ rk[11];
                }
                // per group: counter byte 14 enters through column 1 of round 1
                const uint32_t C1 = Cp1 ^ lut<2, kOffT2>(lb, s3);
                const uint32_t D0 = E0 ^ lut<1, kOffT1>(lb, C1), D1 = E1 ^ lut<0, kOffT0>(lb, C1);
                const uint32_t D2 = E2 ^ lut<3, kOffT3>(lb, C1), D3 = E3 ^ lut<2, kOffT2>(lb, C1);
                // where this warp's rows continue after the group (warp-uniform selects, no branches): the next
                // group of this half; the first group of the other half; the first group of the unit claimed ahead
                const bool last_group = jj + 1 == kGroupsPerUnit;
#if UAES_TT_L2_PREFETCH
                // the next group's four rows of this half (2 KiB = 16 lines) towards L2: the loads that fetch a row ahead
                // into registers then find it there (one row takes a warp about as long as a DRAM access under load)
                if (!last_group && lane < 16) {
                    const int64_t kp = (int64_t)(((G + 1) << 8) + 128 * half + 8 * lane) - (int64_t)a.v0;
                    if (kp >= 0 && (uint64_t)kp < a.nblocks) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.in + kp));
                }
#endif
                const uint64_t Gn = !last_group ? G + 1 : half == 0 ? Gu : Gfirst + unext * kGroupsPerUnit;
                const uint32_t rn = !last_group ? 4 * half : half == 0 ? 4u : 0u;
                const bool okn = !last_group || half == 0 || unext != kQNone;
#if UAES_TT_ROLL_ROWS
                // the two row pairs of the group share ONE copy of the round code (the instruction working set of
                // the two roles together has to fit the 32 KB instruction cache): the pair's U words are selected
                // into place (8 selects per pair), the pointers of the rows fetched ahead likewise
                static_assert(ILP == 2, "rolled form: two row pairs");
#pragma unroll 1
                for (int it = 0; it < 4; it += ILP) {
                    uint4 nxt[ILP];
                    uint32_t t[ILP][4];
                    const bool second = it != 0;
                    const uint64_t Gf = second ? Gn : G;
                    const uint32_t rf = second ? rn : 4 * half + ILP;
                    const bool okf = second ? okn : true;
#pragma unroll
                    for (int i = 0; i < ILP; ++i) {
                        nxt[i] = fetch(okf, Gf, rf + i);
                        t[i][0] = D0 ^ (second ? U[2 + i][0] : U[i][0]); t[i][1] = D1 ^ (second ? U[2 + i][1] : U[i][1]);
                        t[i][2] = D2 ^ (second ? U[2 + i][2] : U[i][2]); t[i][3] = D3 ^ (second ? U[2 + i][3] : U[i][3]);
                    }
                    enc_finish_n<NR, 3, ILP>(lb, t, rk, cur);
#pragma unroll
                    for (int i = 0; i < ILP; ++i) {
                        const int64_t k = kof(G, 4 * half + it + i);
                        if (k >= 0 && (uint64_t)k < a.nblocks) st_stream(a.out + k, make_uint4(t[i][0], t[i][1], t[i][2], t[i][3]));
                        cur[i] = nxt[i];
                    }
                }
#else
#pragma unroll
                for (int it = 0; it < 4; it += ILP) {
                    uint4 nxt[ILP];
                    uint32_t t[ILP][4];
#pragma unroll
                    for (int i = 0; i < ILP; ++i) {
                        if (it + ILP < 4) nxt[i] = fetch(true, G, 4 * half + it + ILP + i);
                        else              nxt[i] = fetch(okn, Gn, rn + i);
                        t[i][0] = D0 ^ U[it + i][0]; t[i][1] = D1 ^ U[it + i][1];
                        t[i][2] = D2 ^ U[it + i][2]; t[i][3] = D3 ^ U[it + i][3];
                    }
                    enc_finish_n<NR, 3, ILP>(lb, t, rk, cur);
#pragma unroll
                    for (int i = 0; i < ILP; ++i) {
                        const int64_t k = kof(G, 4 * half + it + i);
                        if (k >= 0 && (uint64_t)k < a.nblocks) st_stream(a.out + k, make_uint4(t[i][0], t[i][1], t[i][2], t[i][3]));
                        cur[i] = nxt[i];
                    }
                }
#endif
            }
        }
        ++done;
        ucur = unext;
    }
    if (lane == 0 && done) atomicAdd(a.q + 1, (unsigned long long)done);

    // ragged tail: Y[0..n) = E(ctr)[0..n) ^ X[0..n)  (mixThenXor, micro_aes.c:534-544)
    if (a.tail && blockIdx.x == 0 && threadIdx.x == tail_thread) {
        uint32_t w2, w3;
        ctr_words(a.b8, (a.v0 + a.nblocks) & kMask56, w2, w3);
        uint32_t t0 = a.w0, t1 = a.w1, t2 = w2, t3 = w3;
        enc_block<NR>(lb, t0, t1, t2, t3, rk);
        const uint32_t ksw[4] = {t0, t1, t2, t3};
        const uint8_t *x = (const uint8_t *)(a.in + a.nblocks);
        uint8_t *y = (uint8_t *)(a.out + a.nblocks);
        for (uint32_t i = 0; i < a.tail; ++i) y[i] = x[i] ^ (uint8_t)(ksw[i >> 2] >> (8 * (i & 3)));
    }
}

template <int NR, int kCtrThreads, int ILP>
__global__ void __launch_bounds__(kCtrThreads + kBsThreads, 1) ctr_queue_kernel(const __grid_constant__ CtrArgs a)
{
    static_assert(ILP == 2, "two rows in flight per table-driven thread");
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_tables<true>(dyn);
    const uint32_t lane = threadIdx.x & 31;
    constexpr int kLaunchRegs = (65536 / (kCtrThreads + kBsThreads)) / 8 * 8;
#ifndef UAES_Q_TT_REGS
#define UAES_Q_TT_REGS 96                        // table-driven threads; the co-runner gets 224 (104 / 200: -0.8 .. -1.8 %)
#endif
    constexpr int kTtRegs = UAES_Q_TT_REGS;
    constexpr int kBsRegs0 = kLaunchRegs + (kLaunchRegs - kTtRegs) * kCtrThreads / kBsThreads;
    constexpr int kBsRegs = (kBsRegs0 > 232 ? 232 : kBsRegs0) / 8 * 8;

    if (threadIdx.x >= kCtrThreads) {
        reg_inc<kBsRegs>();
        const uint32_t tbase = align_table_base(dyn);
        const uint32_t bw = (threadIdx.x - kCtrThreads) >> 5;
        const uint32_t off = tbase + kEncTableBytes + bw * (kBsUniformMasks * 4) - smem_u32(dyn);
        if (off + kBsUniformMasks * 4 > dyn_smem_size()) __trap();
        if (!a.q_bs_on) return;
        uint32_t *um = (uint32_t *)(dyn + off);
        uint64_t tag16 = ~0ull, done = 0;
        uint64_t u = q_back(q_post(a.q, 1ull << 32, a.q_zero), a.q_units);
        while (u != kQNone) {
            const unsigned long long posted = q_post(a.q, 1ull << 32, a.q_zero);      // the next unit, a unit ahead
            const uint64_t c0 = a.q_u0 + (u << a.q_shift);
#pragma unroll 1
            for (uint32_t p = 0; p < (1u << (a.q_shift - 10)); ++p)
                ctr_bs_pass<NR, UAES_BS_BATCH>(a, lb, um, c0 + ((uint64_t)p << 10), tag16);
            ++done;
            u = q_back(posted, a.q_units);
        }
        if (lane == 0 && done) atomicAdd(a.q + 2, (unsigned long long)done);
        return;
    }
    reg_dec<kTtRegs>();

    ctr_queue_table_role<NR, ILP>(a, lb);
}

// ---- CTR work-queue kernel with NARROW bitsliced warps (uaes_bitslice8.cuh) --------------------------
// Same queue, same table-driven role; the co-runner warps hold 8 blocks per thread in 32 registers
// instead of 32 blocks in 128, so they need 96-112 registers instead of 224 and two of them fit per
// scheduler (more warps to pick from when the table-driven ones wait for the lookup pipe), and
// their round loop is 440 instructions instead of 1 600 (instruction cache).
// One pass = one group (256 counters, 8 rows): the state entering round 3 is U(byte 15) ^ D(group);
// the thread's U planes are computed once per 2^40 blocks and parked in shared memory (lane-private
// words, conflict free), D is expanded into 32 mask words by the 32 lanes of the warp.
// Geometry (profiles/r2_sweep_q8.txt): 16 table-driven warps with ONE row in flight at 64 registers + 8 bitsliced warps
// at 112 (setmaxnreg from a launch allocation of 80): 1100-1115 GiB/s; 12 table-driven warps with two rows in flight +
// 8 bitsliced, 96 registers each, no setmaxnreg: 1080-1090 on the same boxes; 20 + 8 (64 / 88): 1109; 20 + 4: 1078;
// 16 + 12: 1069; table-driven warps at 56 registers: 1073 (16 + 8), 1105 (20 + 8).
#ifndef UAES_Q8_TT
#define UAES_Q8_TT 512
#endif
#ifndef UAES_Q8_BS
#define UAES_Q8_BS 256
#endif
#ifndef UAES_Q8_TT_REGS
#define UAES_Q8_TT_REGS 64                       // 0: both roles keep the launch allocation (no setmaxnreg)
#endif
#ifndef UAES_Q8_MAP
#define UAES_Q8_MAP 1                            // bitsliced warps first: 1080.8 vs 1076.3 GiB/s; segregated by scheduler: 952-957
#endif
#ifndef UAES_Q8_ILP
#define UAES_Q8_ILP 1                            // rows in flight per table-driven thread
#endif
#ifndef UAES_Q8_BATCH
#define UAES_Q8_BATCH 2                          // rows loaded ahead of the XOR / store
#endif
// AES-192 / 256 keep 12 table-driven warps with two rows in flight + 8 bitsliced warps at 96 registers each (no
// setmaxnreg): AES-256 759 GiB/s against 740 in the AES-128 geometry (the longer round loop spills there)
#ifndef UAES_Q8_TT_LONG
#define UAES_Q8_TT_LONG 384
#define UAES_Q8_BS_LONG 256
#define UAES_Q8_ILP_LONG 2
#define UAES_Q8_TT_REGS_LONG 0
#endif
template <int NR> struct Q8Geom {
    static constexpr int TT = NR == 10 ? UAES_Q8_TT : UAES_Q8_TT_LONG, BS = NR == 10 ? UAES_Q8_BS : UAES_Q8_BS_LONG;
    static constexpr int ILP = NR == 10 ? UAES_Q8_ILP : UAES_Q8_ILP_LONG, TTREGS = NR == 10 ? UAES_Q8_TT_REGS : UAES_Q8_TT_REGS_LONG;
};
constexpr uint32_t kBs8WarpWords = 32 * 32 + 2 * 32;   // U planes (32 per lane) + two D-mask buffers

// Rounds 0-2 of one group for the narrow bitsliced warps: s[32] = planes of the state entering round 3.
// K0 / Cp1 / E0..E3 follow the slow counter bytes exactly as in the table-driven role; the thread's U
// planes (a function of counter byte 15 alone) are rebuilt when counter bits >= 40 change and parked in
// shared memory: planes 4q..4q+3 of this thread as one uint4 at ((uint4 *)up)[32 q] (up = warp area + 4 * lane words);
// dm = two buffers of 32 D-mask words per warp.
struct Bs8Hoist {
    uint64_t tag40 = ~0ull, tag16 = ~0ull;
    uint32_t K0 = 0, Cp1 = 0, E0 = 0, E1 = 0, E2 = 0, E3 = 0, flip = 0;

    // vg = counter of the group's first block (byte 15 = 0), reduced mod 2^56
    __device__ __forceinline__ void group(uint32_t (&s)[32], uint32_t lb, const uint32_t *rk, uint32_t w0, uint32_t w1,
                                          uint32_t b8, uint64_t vg, uint32_t *up, uint32_t *dm, uint32_t lane)
    {
        const uint32_t s0 = w0 ^ rk[0], s1 = w1 ^ rk[1];
        uint32_t w2, w3;
        ctr_words(b8, vg, w2, w3);
        const uint32_t s2 = w2 ^ rk[2], s3 = w3 ^ rk[3];                      // byte 15 of the counter is 0 here
        if ((vg >> 40) != tag40) {                                            // once per launch in practice
            tag40 = vg >> 40;
            K0 = lut<0, kOffT0>(lb, s0) ^ lut<1, kOffT1>(lb, s1) ^ lut<2, kOffT2>(lb, s2) ^ rk[4];
            Cp1 = lut<0, kOffT0>(lb, s1) ^ lut<1, kOffT1>(lb, s2) ^ lut<3, kOffT3>(lb, s0) ^ rk[5];
            tag16 = ~0ull;
            uint32_t m[32];                                                   // m[8 c + t] = U word of column c, slot t
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const uint32_t c0 = K0 ^ lut<3, kOffT3>(lb, s3 ^ ((32 * t + lane) << 24));
                m[t] = lut<0, kOffT0>(lb, c0);      m[8 + t] = lut<3, kOffT3>(lb, c0);
                m[16 + t] = lut<2, kOffT2>(lb, c0); m[24 + t] = lut<1, kOffT1>(lb, c0);
            }
            bs_transpose32(m);
#pragma unroll
            for (int j = 0; j < 32; j += 4) ((uint4 *)up)[8 * j] = make_uint4(m[j], m[j + 1], m[j + 2], m[j + 3]);
        }
        if ((vg >> 16) != tag16) {                                            // every 256 groups
            tag16 = vg >> 16;
            const uint32_t C2 = lut<0, kOffT0>(lb, s2) ^ lut<1, kOffT1>(lb, s3) ^ lut<2, kOffT2>(lb, s0) ^ lut<3, kOffT3>(lb, s1) ^ rk[6];
            const uint32_t C3 = lut<0, kOffT0>(lb, s3) ^ lut<1, kOffT1>(lb, s0) ^ lut<2, kOffT2>(lb, s1) ^ lut<3, kOffT3>(lb, s2) ^ rk[7];
            E0 = lut<2, kOffT2>(lb, C2) ^ lut<3, kOffT3>(lb, C3) ^ rk[8];
            E1 = lut<1, kOffT1>(lb, C2) ^ lut<2, kOffT2>(lb, C3) ^ rk[9];
            E2 = lut<0, kOffT0>(lb, C2) ^ lut<1, kOffT1>(lb, C3) ^ rk[10];
            E3 = lut<0, kOffT0>(lb, C3) ^ lut<3, kOffT3>(lb, C2) ^ rk[11];
        }
        const uint32_t C1 = Cp1 ^ lut<2, kOffT2>(lb, s3);                     // counter byte 14 enters through column 1 of round 1
        const uint32_t D0 = E0 ^ lut<1, kOffT1>(lb, C1), D1 = E1 ^ lut<0, kOffT0>(lb, C1);
        const uint32_t D2 = E2 ^ lut<3, kOffT3>(lb, C1), D3 = E3 ^ lut<2, kOffT2>(lb, C1);
        uint32_t *d = dm + flip;                                              // double buffered: one __syncwarp per group is enough
        flip ^= 32;
        d[lane] = bs8_spread(D0, D1, D2, D3, (int)lane);                      // lane j owns plane j of D
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {                                     // 128-bit reads: 16 instead of 64 instructions
            const uint4 u = ((const uint4 *)up)[8 * j], m = ((const uint4 *)d)[j >> 2];
            s[j] = u.x ^ m.x; s[j + 1] = u.y ^ m.y; s[j + 2] = u.z ^ m.z; s[j + 3] = u.w ^ m.w;
        }
    }
};

template <int NR>
__device__ __forceinline__ void ctr_bs8_role(const CtrArgs8 &a, uint32_t lb, uint32_t *ws)
{
    constexpr int BATCH = UAES_Q8_BATCH;
    const uint32_t *rk = a.ks.w;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t *up = ws + 4 * lane;                // planes 4q..4q+3 of this thread's U: ((uint4 *)up)[32 q]
    uint32_t *dm = ws + 32 * 32;                 // D masks, double buffered
    const uint32_t kGroupsPerUnit = 1u << (a.q_shift - 8);
    const uint64_t Gfirst = a.q_u0 >> 8;
    uint64_t done = 0;
    Bs8Hoist hoist;

    uint64_t u = q_back(q_post(a.q, 1ull << 32, a.q_zero), a.q_units);
    while (u != kQNone) {
        const unsigned long long posted = q_post(a.q, 1ull << 32, a.q_zero);      // the next unit, a unit ahead
        const uint64_t Gu = Gfirst + u * kGroupsPerUnit;
#pragma unroll 1
        for (uint32_t jj = 0; jj < kGroupsPerUnit; ++jj) {
            const uint64_t G = Gu + jj;
            const int64_t k0 = (int64_t)((G << 8) + lane) - (int64_t)a.v0;        // block of slot 0 (may be < 0 in the first unit)
            {   // pull the pass's 4 KiB of input towards L2 while the rounds run: one 128-byte line per lane
                const int64_t kp = (int64_t)(G << 8) - (int64_t)a.v0 + 8 * (int64_t)lane;
                if (kp >= 0 && (uint64_t)kp < a.nblocks) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.in + kp));
            }
            uint32_t s[32];
            hoist.group(s, lb, rk, a.w0, a.w1, a.b8, (G << 8) & kMask56, up, dm, lane);
            bs8_finish<NR>(s, a.bs8);
            // XOR with the data: slot t of all lanes = one coalesced 512-byte row, loaded BATCH rows ahead
            auto load_batch = [&](int t0, uint4 (&x)[BATCH]) {
#pragma unroll
                for (int i = 0; i < BATCH; ++i) {
                    const int64_t k = k0 + 32 * (t0 + i);
                    x[i] = (uint64_t)k < a.nblocks ? ld_stream(a.in + k) : make_uint4(0, 0, 0, 0);
                }
            };
            uint4 x[BATCH], y[BATCH];
            load_batch(0, x);
            bs_transpose32(s);                                                    // s[8 c + t] = word c of slot t
#pragma unroll
            for (int t0 = 0; t0 < 8; t0 += BATCH) {
                if (t0 + BATCH < 8) load_batch(t0 + BATCH, y);
#pragma unroll
                for (int i = 0; i < BATCH; ++i) {
                    const int64_t k = k0 + 32 * (t0 + i);
                    const int t = t0 + i;
                    x[i].x ^= s[t]; x[i].y ^= s[8 + t]; x[i].z ^= s[16 + t]; x[i].w ^= s[24 + t];
                    if ((uint64_t)k < a.nblocks) st_stream(a.out + k, x[i]);
                    x[i] = y[i];
                }
            }
        }
        ++done;
        u = q_back(posted, a.q_units);
    }
    if (lane == 0 && done) atomicAdd(a.q + 2, (unsigned long long)done);
}

template <int NR, int TT, int BS, int ILP, int TTREGS>
__global__ void __launch_bounds__(TT + BS, 1) ctr_queue8_kernel(const __grid_constant__ CtrArgs8 a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_tables<true>(dyn);
    constexpr int kLaunchRegs = (65536 / (TT + BS)) / 8 * 8 > 255 ? 248 : (65536 / (TT + BS)) / 8 * 8;
    constexpr int kTtRegs = TTREGS;
    constexpr int kBsRegs = kTtRegs ? (kLaunchRegs + (kLaunchRegs - kTtRegs) * TT / BS) / 8 * 8 : 0;
    // which warps play which role (UAES_Q8_MAP): 0 = the first TT / 32 warps are table-driven; 1 = the LAST ones are;
    // 2 = segregated by scheduler (warp w runs on scheduler w % 4): schedulers 0 and 1 host table-driven warps only,
    // the bitsliced warps share schedulers 2 and 3 with one table-driven warp each (needs TT = 384, BS = 256)
    const uint32_t w = threadIdx.x >> 5;
#if UAES_Q8_MAP == 1
    const bool is_bs = w < BS / 32;
    const uint32_t bw = w;
#elif UAES_Q8_MAP == 2
    static_assert(TT == 384 && BS == 256, "segregated map: 12 + 8 warps");
    const bool is_bs = (w & 3u) >= 2u && w >= 4u;
    const uint32_t bw = ((w >> 2) - 1u) * 2u + ((w & 3u) - 2u);
#else
    const bool is_bs = w >= TT / 32;
    const uint32_t bw = w - TT / 32;
#endif
    if (is_bs) {
        if (kTtRegs) { if (kBsRegs > kLaunchRegs) reg_inc<kBsRegs ? kBsRegs : 24>(); else reg_dec<kBsRegs ? kBsRegs : 24>(); }
        // the warps' areas: in the gap in front of the 64 KiB-aligned tables if it is large enough, else behind them
        const uint32_t d0 = smem_u32(dyn), tb = align_table_base(dyn);
        constexpr uint32_t kNeed = (BS / 32) * kBs8WarpWords * 4;
        uint32_t off = (d0 + 15u & ~15u) - d0;
        if (off + kNeed > tb - d0) off = tb + kEncTableBytes - d0;
        if (off + kNeed > dyn_smem_size()) __trap();
        if (!a.q_bs_on) return;
        ctr_bs8_role<NR>(a, lb, (uint32_t *)(dyn + off) + bw * kBs8WarpWords);
        return;
    }
    if (kTtRegs) { if (kTtRegs < kLaunchRegs) reg_dec<kTtRegs ? kTtRegs : 24>(); else reg_inc<kTtRegs ? kTtRegs : 24>(); }
    ctr_queue_table_role<NR, ILP>(a, lb, UAES_Q8_MAP == 1 ? BS : 0);
}

// ---------------------------------------------------------------- ECB (micro_aes.c:636-680)

struct EcbArgs {
    uaes_keysched ks;
    const uint4 *in;
    uint4 *out;
    uint64_t nblocks;
    uint32_t tail;               // encrypt: zero-padded extra block; decrypt: bytes copied through
    uint32_t pad;                // encrypt: 0 = zero padding of a ragged tail only, 1 = PKCS#7, 2 = ISO/IEC 7816-4
                                 // (the reference's AES_PADDING: the last two ALWAYS add a block, micro_aes.c:610-621)
};

// padBlock (micro_aes.c:610-621): the last block of an ECB encryption from `tail` input bytes
__device__ inline void ecb_pad_block(const uint8_t *x, uint32_t tail, uint32_t pad, uint32_t s[4])
{
    const uint32_t n = 16 - tail;
    s[0] = s[1] = s[2] = s[3] = 0;
    for (uint32_t i = 0; i < 16; ++i) {
        const uint32_t b = i < tail ? x[i] : pad == 1 ? n : (pad == 2 && i == tail) ? 0x80u : 0u;
        s[i >> 2] |= b << (8 * (i & 3));
    }
}

template <int NR, bool ENC>
__global__ void __launch_bounds__(kThreads, 1) ecb_kernel(const __grid_constant__ EcbArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_tables<ENC>(dyn);
    const uint32_t *rk = a.ks.w;
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    uint64_t k = (uint64_t)blockIdx.x * kThreads + threadIdx.x;

    uint4 cur = k < a.nblocks ? ld_stream(a.in + k) : make_uint4(0, 0, 0, 0);
    for (; k < a.nblocks; k += stride) {
        const uint4 nxt = k + stride < a.nblocks ? ld_stream(a.in + k + stride) : make_uint4(0, 0, 0, 0);
        uint32_t s0 = cur.x, s1 = cur.y, s2 = cur.z, s3 = cur.w;
        if (ENC) enc_block<NR>(lb, s0, s1, s2, s3, rk); else dec_block<NR>(lb, s0, s1, s2, s3, rk);
        st_stream(a.out + k, make_uint4(s0, s1, s2, s3));
        cur = nxt;
    }

    if ((a.tail || (ENC && a.pad)) && blockIdx.x == 0 && threadIdx.x == 0) {
        const uint8_t *x = (const uint8_t *)(a.in + a.nblocks);
        uint8_t *y = (uint8_t *)(a.out + a.nblocks);
        if (ENC) {                                    // padBlock, micro_aes.c:610-621
            uint32_t s[4];
            ecb_pad_block(x, a.tail, a.pad, s);
            enc_block<NR>(lb, s[0], s[1], s[2], s[3], rk);
            for (uint32_t i = 0; i < 16; ++i) y[i] = (uint8_t)(s[i >> 2] >> (8 * (i & 3)));
        } else {                                      // the memcpy of micro_aes.c:667 leaves them as is
            for (uint32_t i = 0; i < a.tail; ++i) y[i] = x[i];
        }
    }
}

// ---- ECB encryption with the co-runner: 12 table-driven warps (two rows in flight) + one warpgroup
// of bitsliced warps in the general form (uaes_bitslice.cuh); tile = 1024 blocks, lane l, slot t <->
// block 32 t + l of the tile, so every access is a coalesced 512-byte row.
// CFB = true turns it into CFB decryption (micro_aes.c:799-845): P_k = E(C_(k-1)) ^ C_k, C_(-1) = IV --
// the same cipher calls on the input shifted by one block, the ciphertext XORed in at the end.
struct EcbHybridArgs {
    EcbArgs e;
    uint64_t tt_blocks;          // blocks [0, tt_blocks): table-driven warps; a multiple of 1024
    uint32_t iv[4];              // CFB / CBC only
    unsigned long long *q;       // ecb_dec_hybrid_kernel: non-null = no static split, the two-ended work queue of ctr_queue_kernel, unit = a tile of 1024 blocks
    uint32_t q_zero;             // 0 (see q_post)
    BsKeyPlanesFull bs;
};

// table-driven side: UAES_ECB_TT threads with UAES_ECB_ILP rows in flight at UAES_ECB_TT_REGS registers (profiles/r2_sweep_ecb_ilp.txt)
#ifndef UAES_ECB_TT
#define UAES_ECB_TT 384
#endif
#ifndef UAES_ECB_ILP
#define UAES_ECB_ILP 2
#endif
#ifndef UAES_ECB_TT_REGS
#define UAES_ECB_TT_REGS kHybridTtRegs
#endif
// CFB decryption (the same kernel, CFB = true) does gain from the 16-warp geometry: 884 against 854 GiB/s
#ifndef UAES_CFB_TT
#define UAES_CFB_TT 512
#endif
#ifndef UAES_CFB_ILP
#define UAES_CFB_ILP 1
#endif
#ifndef UAES_CFB_TT_REGS
#define UAES_CFB_TT_REGS 64
#endif
template <bool CFB> struct EcbGeom {
    static constexpr int TT = CFB ? UAES_CFB_TT : UAES_ECB_TT, ILP = CFB ? UAES_CFB_ILP : UAES_ECB_ILP;
    static constexpr int TTREGS = CFB ? UAES_CFB_TT_REGS : UAES_ECB_TT_REGS;
};

template <int NR, bool CFB>
__global__ void __launch_bounds__(EcbGeom<CFB>::TT + kBsThreads, 1) ecb_hybrid_kernel(const __grid_constant__ EcbHybridArgs a)
{
    const uint4 iv = make_uint4(a.iv[0], a.iv[1], a.iv[2], a.iv[3]);
    // cipher input of block k: the block itself (ECB) or its predecessor (CFB)
    auto cin = [&](uint64_t k) -> uint4 { return !CFB ? ld_stream(a.e.in + k) : k ? a.e.in[k - 1] : iv; };
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_tables<true>(dyn);
    const uint32_t *rk = a.e.ks.w;
    const uint32_t lane = threadIdx.x & 31;
    constexpr int kTt = EcbGeom<CFB>::TT, kTtWarps = kTt / 32;
    constexpr int kLaunchRegs = (65536 / (kTt + kBsThreads)) / 8 * 8;
    constexpr int kTtRegs = EcbGeom<CFB>::TTREGS, kBsRegs = (kLaunchRegs + (kLaunchRegs - kTtRegs) * kTt / kBsThreads) / 8 * 8;

    if (threadIdx.x >= kTt) {
        reg_inc<kBsRegs>();
        const uint64_t ntiles = (a.e.nblocks - a.tt_blocks + 1023) / 1024;
        const uint64_t gw = (uint64_t)blockIdx.x * (kBsThreads / 32) + ((threadIdx.x - kTt) >> 5);
        const uint64_t nw = (uint64_t)gridDim.x * (kBsThreads / 32);
        const uint64_t per = (ntiles + nw - 1) / nw;
        const uint64_t p0 = gw * per < ntiles ? gw * per : ntiles;
        const uint64_t p1 = p0 + per < ntiles ? p0 + per : ntiles;
        for (uint64_t tile = p0; tile < p1; ++tile) {
            const uint64_t kb = a.tt_blocks + tile * 1024 + lane;
            uint32_t s[128];
#pragma unroll
            for (int tb = 0; tb < 32; tb += kBsLoadBatch) {
                uint4 v[kBsLoadBatch];
#pragma unroll
                for (int i = 0; i < kBsLoadBatch; ++i) v[i] = kb + 32 * (tb + i) < a.e.nblocks ? cin(kb + 32 * (tb + i)) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int i = 0; i < kBsLoadBatch; ++i) { s[tb + i] = v[i].x; s[32 + tb + i] = v[i].y; s[64 + tb + i] = v[i].z; s[96 + tb + i] = v[i].w; }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
            bs_encrypt_planes<NR>(s, a.bs);
#pragma unroll
            for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
#pragma unroll
            for (int tb = 0; tb < 32; tb += 4) {
                uint4 x[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    x[i] = CFB && kb + 32 * (tb + i) < a.e.nblocks ? ld_stream(a.e.in + kb + 32 * (tb + i)) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int t = tb + i;
                    if (kb + 32 * t < a.e.nblocks)
                        st_stream(a.e.out + kb + 32 * t, make_uint4(s[t] ^ x[i].x, s[32 + t] ^ x[i].y, s[64 + t] ^ x[i].z, s[96 + t] ^ x[i].w));
                }
            }
        }
        return;
    }
    reg_dec<kTtRegs>();

    constexpr int ILP = EcbGeom<CFB>::ILP;                       // 32-block rows in flight per thread
    const uint64_t nsteps = a.tt_blocks / (32 * ILP);            // tt_blocks is a multiple of 1024
    const uint64_t gw = (uint64_t)blockIdx.x * kTtWarps + (threadIdx.x >> 5);
    const uint64_t nw = (uint64_t)gridDim.x * kTtWarps;
    const uint64_t per = (nsteps + nw - 1) / nw;
    const uint64_t q0 = gw * per < nsteps ? gw * per : nsteps;
    const uint64_t q1 = q0 + per < nsteps ? q0 + per : nsteps;
    uint4 cur[ILP], nxt[ILP];
    if (q0 < q1) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) cur[i] = cin(q0 * (32 * ILP) + 32 * i + lane);
    }
    for (uint64_t q = q0; q < q1; ++q) {
        const uint64_t k = q * (32 * ILP) + lane;
        uint4 zero[ILP];
        uint32_t st[ILP][4];
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (q + 1 < q1) nxt[i] = cin(k + 32 * ILP + 32 * i);
            zero[i] = CFB ? ld_stream(a.e.in + k + 32 * i) : make_uint4(0, 0, 0, 0);             // the ciphertext XORed in at the end
            st[i][0] = cur[i].x ^ rk[0]; st[i][1] = cur[i].y ^ rk[1]; st[i][2] = cur[i].z ^ rk[2]; st[i][3] = cur[i].w ^ rk[3];
        }
        enc_finish_n<NR, 1, ILP>(lb, st, rk, zero);
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            st_stream(a.e.out + k + 32 * i, make_uint4(st[i][0], st[i][1], st[i][2], st[i][3]));
            cur[i] = nxt[i];
        }
    }

    if ((a.e.tail || (!CFB && a.e.pad)) && blockIdx.x == 0 && threadIdx.x == 0) {
        const uint8_t *x = (const uint8_t *)(a.e.in + a.e.nblocks);
        uint8_t *y = (uint8_t *)(a.e.out + a.e.nblocks);
        if (!CFB) {                                              // padBlock, micro_aes.c:610-621
            uint32_t s[4];
            ecb_pad_block(x, a.e.tail, a.e.pad, s);
            enc_block<NR>(lb, s[0], s[1], s[2], s[3], rk);
            for (uint32_t i = 0; i < 16; ++i) y[i] = (uint8_t)(s[i >> 2] >> (8 * (i & 3)));
        } else {                                                 // ragged last block: E(C_(m-1)) ^ C_m, micro_aes.c:840-845
            const uint4 ch = cin(a.e.nblocks);
            uint32_t s[4] = {ch.x, ch.y, ch.z, ch.w};
            enc_block<NR>(lb, s[0], s[1], s[2], s[3], rk);
            for (uint32_t i = 0; i < a.e.tail; ++i) y[i] = x[i] ^ (uint8_t)(s[i >> 2] >> (8 * (i & 3)));
        }
    }
}

// ---------------------------------------------------------------- synthetic data, checksum

__device__ __forceinline__ uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void fill_kernel(uint64_t seed, uint64_t first, uint64_t *dst, uint64_t n)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = splitmix64(seed + first + i);
}

__global__ void xor_fold_kernel(const uint64_t *src, uint64_t n, unsigned long long *result)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        acc ^= src[i];
    for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicXor(result, (unsigned long long)acc);
}

// ---------------------------------------------------------------- launch plumbing

static int sm_count()
{
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

template <typename K>
static cudaError_t opt_in_smem(K kernel)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmem);
}

// grid = min(#SM, work units / warps per CTA), at least 1
static unsigned grid_for(uint64_t warp_units)
{
    const uint64_t need = (warp_units + kWarpsPerCta - 1) / kWarpsPerCta;
    const uint64_t sms = (uint64_t)sm_count();
    return (unsigned)(need < 1 ? 1 : need < sms ? need : sms);
}

template <int NR, int kCtrThreads, bool BS, int ILP = 1>
static cudaError_t launch_ctr_nt(const CtrArgs &a, cudaStream_t st)
{
    constexpr int kCtrWarps = kCtrThreads / 32;
    cudaError_t e = opt_in_smem(ctr_kernel<NR, kCtrThreads, BS, ILP>);
    if (e != cudaSuccess) return e;
    const uint64_t ngroups = (((uint32_t)a.v0 & 255u) + a.tt_blocks + 255) >> 8;
    // a pair of warps per group; at least 4 groups per pair before another CTA is worth its table fill
    uint64_t ctas = (ngroups + 4 * (kCtrWarps / 2) - 1) / (4 * (kCtrWarps / 2));
    if (BS && ctas < (a.bs_passes + 3) / 4) ctas = (a.bs_passes + 3) / 4;    // one pass per co-runner warp
    const uint64_t sms = (uint64_t)sm_count();
    ctr_kernel<NR, kCtrThreads, BS, ILP><<<(unsigned)(ctas < 1 ? 1 : ctas < sms ? ctas : sms), kCtrThreads + (BS ? kBsThreads : 0), kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

// CTR geometry: share of the blocks given to the bitsliced warps in 1/1024 (tuned on B200,
// profiles/; 0 turns the co-runner off), threads of the table-driven warps, and the call size
// below which one kernel flavour is simpler and as fast.  Environment variables set the initial
// values, uaes_ctr_tuning() changes them at run time (tests sweep them).
static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

static int g_ctr_threads = 0, g_ctr_share = -1;
static long long g_ctr_bs_min = 1ll << 23;    // 128 MiB: below it the co-runner's whole passes (1024 blocks x 592 warps)
                                               // quantise badly: 606 vs 659 GiB/s at 32 MiB, 716 vs 782 at 64 MiB, 879 vs 863 at 128 MiB

// Measured on B200, AES-128, 16 GiB (profiles/r1_ctr_hybrid_sweep.txt): table-driven warps alone
// 942-946 GiB/s; with the co-runner 1000 / 1006 / 1004 / 980 GiB/s at 170 / 185 / 200 / 215 per
// 1024, falling off quickly once the bitsliced warps become the tail.  Geometry codes:
//   384  = 384 table-driven threads, one block per thread in flight (+128 co-runner threads)
//   385  = the same with two blocks per thread in flight (default; the co-runner's instructions
//          lengthen every lookup round trip, the second block hides it: 968 -> 1006 GiB/s)
//   512 / 768 / 1024 = table-driven warps only
//   386  = the work-queue kernel (ctr_queue_kernel): 384 table-driven threads, two rows in flight,
//          + 128 co-runner threads; no static split, bs_permille only switches the co-runner on / off
//   388  = the work-queue kernel with NARROW bitsliced warps (ctr_queue8_kernel, default): 512 table-driven threads
//          with one row in flight (64 registers) + 256 co-runner threads of 8 blocks each (112 registers)
#ifndef UAES_CTR_DEFAULT_GEOMETRY
#define UAES_CTR_DEFAULT_GEOMETRY 388
#endif
constexpr int kCtrDefaultGeometry = UAES_CTR_DEFAULT_GEOMETRY, kCtrDefaultShare = 195;

// ---- queue words: one 32-byte slot per launch, from a per-device ring (a slot comes round again
// after kQRing launches on that device; far more than can be in flight)
constexpr int kQRing = 4096;
static unsigned long long *g_qring[64];
static std::atomic<unsigned> g_qnext[64];
static std::atomic_flag g_qlock = ATOMIC_FLAG_INIT;
static thread_local unsigned long long *tls_last_q = nullptr;
static thread_local unsigned long long tls_last_unit = 0;

static cudaError_t q_slot(cudaStream_t st, unsigned long long **out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!g_qring[dev]) {
        while (g_qlock.test_and_set(std::memory_order_acquire)) { }
        if (!g_qring[dev]) {
            void *p = nullptr;
            e = cudaMalloc(&p, (size_t)kQRing * 32);
            if (e == cudaSuccess) g_qring[dev] = (unsigned long long *)p;
        }
        g_qlock.clear(std::memory_order_release);
        if (e != cudaSuccess) return e;
    }
    unsigned long long *q = g_qring[dev] + 4 * (size_t)(g_qnext[dev].fetch_add(1) % kQRing);
    e = cudaMemsetAsync(q, 0, 32, st);
    *out = q;
    tls_last_q = q;
    return e;
}

static void ctr_tuning_init()
{
    if (g_ctr_threads) return;
    g_ctr_threads = env_int("UAES_CTR_THREADS", kCtrDefaultGeometry);
    g_ctr_share = env_int("UAES_CTR_BS_PERMILLE", kCtrDefaultShare);
}

template <int NR>
static cudaError_t launch_ctr_nr(CtrArgs &a, cudaStream_t st)
{
    ctr_tuning_init();
    const int threads = g_ctr_threads, share = g_ctr_share;
    a.tt_blocks = a.nblocks; a.bs_u0 = 0; a.bs_passes = 0;
    a.q = nullptr; a.q_u0 = 0; a.q_units = 0; a.q_bs_on = 0; a.q_zero = 0; a.q_shift = kQUnitShift;
    if (threads == 386) {
        // work queue: units of 2^q_shift counters, aligned in counter space; both kinds of warps clip to
        // [v0, v0 + nblocks).  Short calls keep the co-runner out (a bitsliced unit takes longer than a
        // table-driven one, which shows when there are fewer units than warps).
        // units of 2048 blocks; 1024 when a warp gets fewer than ~64 of them (below 4 GiB): the last unit of the
        // slowest warp is the tail of the launch (1 GiB: 979 -> see profiles/r2_sweep_ctr_unit_small.txt)
        a.q_shift = (uint32_t)env_int("UAES_CTR_UNIT_SHIFT", a.nblocks >= (1ull << 28) ? kQUnitShift : kQUnitShift - 1);
        if (a.q_shift < 10 || a.q_shift > 16) a.q_shift = kQUnitShift;
        const uint64_t unit = 1ull << a.q_shift;
        a.q_u0 = a.v0 & ~(unit - 1);
        a.q_units = (a.v0 - a.q_u0 + a.nblocks + unit - 1) >> a.q_shift;
        tls_last_unit = unit;
        a.q_bs_on = share > 0 && (long long)a.nblocks >= g_ctr_bs_min;
        if (a.q_bs_on) bs_make_key_planes(a.ks.w, NR, &a.bs);
        cudaError_t e = opt_in_smem(ctr_queue_kernel<NR, 384, 2>);
        if (e != cudaSuccess) return e;
        if ((e = q_slot(st, &a.q)) != cudaSuccess) return e;
        // one CTA per SM; fewer when there is not even one unit per table-driven warp (latency of short calls:
        // a unit takes a warp ~55 us, a CTA's table fill ~10 us)
        const uint64_t sms = (uint64_t)sm_count(), want = (a.q_units + 12 - 1) / 12;
        ctr_queue_kernel<NR, 384, 2><<<(unsigned)(want < 1 ? 1 : want < sms ? want : sms), 384 + kBsThreads, kDynSmem, st>>>(a);
        ++g_launches;
        return cudaGetLastError();
    }
    if (threads == 388) {
        // the work queue with narrow bitsliced warps (ctr_queue8_kernel)
        a.q_shift = (uint32_t)env_int("UAES_CTR_UNIT_SHIFT", a.nblocks >= (1ull << 28) ? kQUnitShift : kQUnitShift - 1);
        if (a.q_shift < 10 || a.q_shift > 16) a.q_shift = kQUnitShift;
        const uint64_t unit = 1ull << a.q_shift;
        a.q_u0 = a.v0 & ~(unit - 1);
        a.q_units = (a.v0 - a.q_u0 + a.nblocks + unit - 1) >> a.q_shift;
        tls_last_unit = unit;
        a.q_bs_on = share > 0 && (long long)a.nblocks >= g_ctr_bs_min;
        using G = Q8Geom<NR>;
        auto kernel = ctr_queue8_kernel<NR, G::TT, G::BS, G::ILP, G::TTREGS>;
        cudaError_t e = opt_in_smem(kernel);
        if (e != cudaSuccess) return e;
        if ((e = q_slot(st, &a.q)) != cudaSuccess) return e;
        static thread_local CtrArgs8 a8;
        static_cast<CtrArgsBase &>(a8) = a;
        if (a.q_bs_on) bs8_make_key_planes(a.ks.w, NR, &a8.bs8);
        const uint64_t sms = (uint64_t)sm_count(), ttw = G::TT / 32, want = (a.q_units + ttw - 1) / ttw;
        kernel<<<(unsigned)(want < 1 ? 1 : want < sms ? want : sms), G::TT + G::BS, kDynSmem, st>>>(a8);
        ++g_launches;
        return cudaGetLastError();
    }
    if (share > 0 && (threads == 384 || threads == 385) && (long long)a.nblocks >= g_ctr_bs_min) {
        // bitsliced range = [S, v0 + nblocks) with S a multiple of 1024 in counter space
        const uint64_t want = a.nblocks / 1024 * (uint64_t)share;           // blocks for the co-runner
        const uint64_t end = a.v0 + a.nblocks;
        uint64_t s0 = (end - want) & ~1023ull;
        if (s0 < a.v0) s0 = (a.v0 + 1023) & ~1023ull;
        if (s0 < end) {
            a.tt_blocks = s0 - a.v0; a.bs_u0 = s0; a.bs_passes = (end - s0 + 1023) >> 10;
            bs_make_key_planes(a.ks.w, NR, &a.bs);
            return threads == 384 ? launch_ctr_nt<NR, 384, true, 1>(a, st) : launch_ctr_nt<NR, 384, true, 2>(a, st);
        }
    }
    switch (threads) {
    case 384:  return launch_ctr_nt<NR, 384, false, 1>(a, st);
    case 385:  return launch_ctr_nt<NR, 384, false, 2>(a, st);
    case 1024: return launch_ctr_nt<NR, 1024, false>(a, st);
    case 512:  return launch_ctr_nt<NR, 512, false>(a, st);
    default:   return launch_ctr_nt<NR, 768, false>(a, st);
    }
}

// ---- ECB decryption with the co-runner: the shape of xts_sectors_hybrid_kernel's decrypt direction -- 16 table-driven
// warps with one row in flight at 64 registers (Td tables) + one warpgroup of bitsliced warps running the equivalent
// inverse cipher (bs_decrypt_planes) on tiles of 1024 blocks at 224.  a.e.ks = the inverse schedule (uaes_host.c).
#ifndef UAES_ECBDEC_TT
#define UAES_ECBDEC_TT 512
#endif
#ifndef UAES_ECBDEC_ILP
#define UAES_ECBDEC_ILP 1
#endif
#ifndef UAES_ECBDEC_TT_REGS
#define UAES_ECBDEC_TT_REGS 64
#endif
constexpr int kEcbDecTtThreads = UAES_ECBDEC_TT;

// CBC = true turns it into CBC decryption (micro_aes.c:746-782): P_k = D(C_k) ^ C_(k-1), C_(-1) = IV -- the neighbour's
// ciphertext block is a second, cache-resident load (the CS3 pair at the end is chain_dec_kernel's, uaes_chain.cuh).
template <int NR, bool CBC>
__global__ void __launch_bounds__(kEcbDecTtThreads + kBsThreads, 1) ecb_dec_hybrid_kernel(const __grid_constant__ EcbHybridArgs a)
{
    const uint4 iv = make_uint4(a.iv[0], a.iv[1], a.iv[2], a.iv[3]);
    auto prev = [&](uint64_t k) -> uint4 { return !CBC ? make_uint4(0, 0, 0, 0) : k ? a.e.in[k - 1] : iv; };
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_tables<false>(dyn);
    const uint32_t *dk = a.e.ks.w;
    const uint32_t lane = threadIdx.x & 31;
    constexpr int kTtWarps = kEcbDecTtThreads / 32;
    constexpr int kLaunchRegs = (65536 / (kEcbDecTtThreads + kBsThreads)) / 8 * 8;
    constexpr int kTtRegs = UAES_ECBDEC_TT_REGS, kBsRegs = (kLaunchRegs + (kLaunchRegs - kTtRegs) * kEcbDecTtThreads / kBsThreads) / 8 * 8;

    const uint64_t nt_all = (a.e.nblocks + 1023) / 1024;         // work-queue mode: tiles of the whole call
    if (threadIdx.x >= kEcbDecTtThreads) {
        reg_inc<kBsRegs>();
        auto do_tile = [&](uint64_t first_block) {
            const uint64_t kb = first_block + lane;
            uint32_t s[128];
#pragma unroll
            for (int tb = 0; tb < 32; tb += kBsLoadBatch) {
                uint4 v[kBsLoadBatch];
#pragma unroll
                for (int i = 0; i < kBsLoadBatch; ++i) v[i] = kb + 32 * (tb + i) < a.e.nblocks ? ld_stream(a.e.in + kb + 32 * (tb + i)) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int i = 0; i < kBsLoadBatch; ++i) { s[tb + i] = v[i].x; s[32 + tb + i] = v[i].y; s[64 + tb + i] = v[i].z; s[96 + tb + i] = v[i].w; }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
            bs_decrypt_planes<NR>(s, a.bs);
#pragma unroll
            for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
#pragma unroll
            for (int tb = 0; tb < 32; tb += 4) {
                uint4 x[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = kb + 32 * (tb + i) < a.e.nblocks ? prev(kb + 32 * (tb + i)) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int t = tb + i;
                    if (kb + 32 * t < a.e.nblocks)
                        st_stream(a.e.out + kb + 32 * t, make_uint4(s[t] ^ x[i].x, s[32 + t] ^ x[i].y, s[64 + t] ^ x[i].z, s[96 + t] ^ x[i].w));
                }
            }
        };
        if (a.q) {                                   // work queue: tiles from the BACK, the next one claimed a tile ahead
            uint64_t u = q_back(q_post(a.q, 1ull << 32, a.q_zero), nt_all);
            while (u != kQNone) {
                const unsigned long long posted = q_post(a.q, 1ull << 32, a.q_zero);
                do_tile(u * 1024);
                u = q_back(posted, nt_all);
            }
            return;
        }
        const uint64_t ntiles = (a.e.nblocks - a.tt_blocks + 1023) / 1024;
        const uint64_t gw = (uint64_t)blockIdx.x * (kBsThreads / 32) + ((threadIdx.x - kEcbDecTtThreads) >> 5);
        const uint64_t nw = (uint64_t)gridDim.x * (kBsThreads / 32);
        const uint64_t per = (ntiles + nw - 1) / nw;
        const uint64_t p0 = gw * per < ntiles ? gw * per : ntiles;
        const uint64_t p1 = p0 + per < ntiles ? p0 + per : ntiles;
        for (uint64_t tile = p0; tile < p1; ++tile) do_tile(a.tt_blocks + tile * 1024);
        return;
    }
    reg_dec<kTtRegs>();

    if (a.q) {
        // work queue: tiles from the FRONT; the next tile is claimed a tile ahead, its answer read half way through
        // this one, its first row requested with this one's last
        uint64_t tile = q_front(q_post(a.q, 1ull, a.q_zero), nt_all);
        uint4 cur = make_uint4(0, 0, 0, 0);
        if (tile != kQNone && tile * 1024 + lane < a.e.nblocks) cur = ld_stream(a.e.in + tile * 1024 + lane);
        while (tile != kQNone) {
            const unsigned long long posted = q_post(a.q, 1ull, a.q_zero);
            uint64_t next = kQNone;
            const uint64_t kb = tile * 1024 + lane;
#pragma unroll 1
            for (int r = 0; r < 32; ++r) {
                const uint64_t k = kb + 32 * r;
                if (r == 16) next = q_front(posted, nt_all);
                const uint64_t kn = r + 1 < 32 ? k + 32 : next * 1024 + lane;
                const bool okn = (r + 1 < 32 || next != kQNone) && kn < a.e.nblocks;
                const uint4 nxt = okn ? ld_stream(a.e.in + kn) : make_uint4(0, 0, 0, 0);
                const uint4 x[1] = {k < a.e.nblocks ? prev(k) : make_uint4(0, 0, 0, 0)};
                uint32_t st[1][4] = {{cur.x, cur.y, cur.z, cur.w}};
                dec_block_n<NR, 1>(lb, st, dk, x);
                if (k < a.e.nblocks) st_stream(a.e.out + k, make_uint4(st[0][0], st[0][1], st[0][2], st[0][3]));
                cur = nxt;
            }
            tile = next;
        }
        if (!CBC && a.e.tail && blockIdx.x == 0 && threadIdx.x == 0) {
            const uint8_t *x = (const uint8_t *)(a.e.in + a.e.nblocks);
            uint8_t *y = (uint8_t *)(a.e.out + a.e.nblocks);
            for (uint32_t i = 0; i < a.e.tail; ++i) y[i] = x[i];
        }
        return;
    }

    constexpr int ILP = UAES_ECBDEC_ILP;
    const uint64_t nsteps = a.tt_blocks / (32 * ILP);            // tt_blocks is a multiple of 1024
    const uint64_t gw = (uint64_t)blockIdx.x * kTtWarps + (threadIdx.x >> 5);
    const uint64_t nw = (uint64_t)gridDim.x * kTtWarps;
    const uint64_t per = (nsteps + nw - 1) / nw;
    const uint64_t q0 = gw * per < nsteps ? gw * per : nsteps;
    const uint64_t q1 = q0 + per < nsteps ? q0 + per : nsteps;
    uint4 cur[ILP], nxt[ILP];
    if (q0 < q1) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) cur[i] = ld_stream(a.e.in + q0 * (32 * ILP) + 32 * i + lane);
    }
    for (uint64_t q = q0; q < q1; ++q) {
        const uint64_t k = q * (32 * ILP) + lane;
        uint4 zero[ILP];
        uint32_t st[ILP][4];
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (q + 1 < q1) nxt[i] = ld_stream(a.e.in + k + 32 * ILP + 32 * i);
            zero[i] = prev(k + 32 * i);                          // the neighbour lane loads it too: cache hit
            st[i][0] = cur[i].x; st[i][1] = cur[i].y; st[i][2] = cur[i].z; st[i][3] = cur[i].w;
        }
        dec_block_n<NR, ILP>(lb, st, dk, zero);
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            st_stream(a.e.out + k + 32 * i, make_uint4(st[i][0], st[i][1], st[i][2], st[i][3]));
            cur[i] = nxt[i];
        }
    }
    if (!CBC && a.e.tail && blockIdx.x == 0 && threadIdx.x == 0) {   // the memcpy of micro_aes.c:667 leaves the ragged bytes as they are
        const uint8_t *x = (const uint8_t *)(a.e.in + a.e.nblocks);
        uint8_t *y = (uint8_t *)(a.e.out + a.e.nblocks);
        for (uint32_t i = 0; i < a.e.tail; ++i) y[i] = x[i];
    }
}

#ifndef UAES_ECB_DEC_DEFAULT_SHARE
#define UAES_ECB_DEC_DEFAULT_SHARE 150           // static split: 804 / 903 / 924 / 816 GiB/s at 0 / 130 / 160 / 190; only > 0 matters with the work queue (default)
#endif
constexpr int kEcbDecDefaultShare = UAES_ECB_DEC_DEFAULT_SHARE;
constexpr int kCfbDefaultShare = 170;   // CFB decryption in the 16-warp geometry: 870 / 892 / 883 / 842 GiB/s at 150 / 175 / 195 / 215 (profiles/r2_sweep_cfb_ilp.txt)
constexpr int kEcbDefaultShare = 195;   // 791 / 830 / 849 / 864 / 814 GiB/s at 0 / 100 / 140 / 180 / 220 (AES-128, profiles/r1_ecb_hybrid_sweep.txt)

template <int NR, bool CFB = false>
static cudaError_t launch_ecb_hybrid_nr(const EcbArgs &e0, uint64_t bs_blocks, cudaStream_t st, const uint32_t *iv = nullptr)
{
    cudaError_t e = opt_in_smem(ecb_hybrid_kernel<NR, CFB>);
    if (e != cudaSuccess) return e;
    static thread_local EcbHybridArgs a;                 // 8 KB of planes: off the stack, one per calling thread
    a.e = e0;
    for (int c = 0; c < 4; ++c) a.iv[c] = iv ? iv[c] : 0;
    a.tt_blocks = (e0.nblocks - bs_blocks) & ~1023ull;
    a.q = nullptr; a.q_zero = 0;
    bs_make_key_planes_full(e0.ks.w, NR, &a.bs);
    const uint64_t need = (e0.nblocks + 32 * 16 - 1) / (32 * 16), sms = (uint64_t)sm_count();
    ecb_hybrid_kernel<NR, CFB><<<(unsigned)(need < sms ? need : sms), EcbGeom<CFB>::TT + kBsThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

// ECB (CBC = false) or CBC (true) decryption of enough data with the co-runner; done = false: not applicable, nothing launched
template <int NR, bool CBC>
static cudaError_t launch_ecb_dec_hybrid_nr(const EcbArgs &a, const uint32_t *iv, cudaStream_t st, bool &done)
{
    done = false;
    ctr_tuning_init();
    const int share = g_ctr_share != kCtrDefaultShare ? g_ctr_share : env_int("UAES_ECB_DEC_BS_PERMILLE", kEcbDecDefaultShare);
    if (!(g_ctr_share > 0 && share > 0 && (long long)a.nblocks >= g_ctr_bs_min && a.nblocks >= 2048)) return cudaSuccess;
    const uint64_t bs_blocks = a.nblocks / 1024 * (uint64_t)share;
    if (!bs_blocks) return cudaSuccess;
    done = true;
    cudaError_t e = opt_in_smem(ecb_dec_hybrid_kernel<NR, CBC>);
    if (e != cudaSuccess) return e;
    static thread_local EcbHybridArgs h;
    h.e = a;
    for (int c = 0; c < 4; ++c) h.iv[c] = iv ? iv[c] : 0;
    h.tt_blocks = (a.nblocks - bs_blocks) & ~1023ull;
    h.q = nullptr; h.q_zero = 0;
    if (env_int("UAES_ECB_DEC_QUEUE", 1)) {              // dynamic split (the static share is ignored)
        if ((e = q_slot(st, &h.q)) != cudaSuccess) return e;
    }
    bs_make_key_planes_full(a.ks.w, NR, &h.bs);
    const uint64_t need = (a.nblocks + 32 * 16 - 1) / (32 * 16), sms = (uint64_t)sm_count();
    ecb_dec_hybrid_kernel<NR, CBC><<<(unsigned)(need < sms ? need : sms), kEcbDecTtThreads + kBsThreads, kDynSmem, st>>>(h);
    ++g_launches;
    return cudaGetLastError();
}

template <int NR, bool ENC>
static cudaError_t launch_ecb_nr(const EcbArgs &a, cudaStream_t st)
{
    if (ENC) {                                           // encryption of enough data: with the co-runner
        ctr_tuning_init();
        const int share = g_ctr_share != kCtrDefaultShare ? g_ctr_share : env_int("UAES_ECB_BS_PERMILLE", kEcbDefaultShare);
        if (g_ctr_share > 0 && share > 0 && (long long)a.nblocks >= g_ctr_bs_min && a.nblocks >= 2048) {
            const uint64_t bs_blocks = a.nblocks / 1024 * (uint64_t)share;
            if (bs_blocks) return launch_ecb_hybrid_nr<NR>(a, bs_blocks, st);
        }
    }
    if (!ENC) {                                          // decryption of enough data: with the inverse-cipher co-runner
        bool done = false;
        const cudaError_t e = launch_ecb_dec_hybrid_nr<NR, false>(a, nullptr, st, done);
        if (done) return e;
    }
    cudaError_t e = opt_in_smem(ecb_kernel<NR, ENC>);
    if (e != cudaSuccess) return e;
    ecb_kernel<NR, ENC><<<grid_for((a.nblocks + 31) / 32), kThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace uaes

using namespace uaes;

#include "uaes_xts.cuh"
#include "uaes_gcm.cuh"
#include "uaes_chain.cuh"
#include "uaes_ocb.cuh"
#include "uaes_batch.cuh"

extern "C" {

u64 uaes_launch_count(void) { return g_launches; }

/* units the two kinds of warps of the calling thread's most recent work-queue CTR launch ended up
 * with (synchronises the device); unit_blocks = blocks per unit */
int uaes_launch_ctr_queue_stats(u64 *tt_units, u64 *bs_units, u64 *unit_blocks)
{
    unsigned long long h[4] = {0, 0, 0, 0};
    if (!tls_last_q) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaMemcpy(h, tls_last_q, 32, cudaMemcpyDeviceToHost);
    *tt_units = h[1]; *bs_units = h[2]; *unit_blocks = tls_last_unit;
    return (int)e;
}

void uaes_launch_ctr_tuning(int tt_threads, int bs_permille, long long bs_min_blocks)
{
    ctr_tuning_init();
    if (tt_threads > 0) g_ctr_threads = tt_threads;
    if (bs_permille >= 0) g_ctr_share = bs_permille > 1024 ? 1024 : bs_permille;
    if (bs_min_blocks >= 0) g_ctr_bs_min = bs_min_blocks;
}

int uaes_launch_ctr(const uaes_keysched *ks, const uaes_ctrblock *cb, const void *in, void *out,
                    u64 len, void *stream)
{
    if (len == 0) return 0;
    CtrArgs a;
    a.ks = *ks;
    a.w0 = cb->w0; a.w1 = cb->w1; a.b8 = cb->b8; a.v0 = cb->v0 & kMask56;
    a.in = (const uint4 *)in; a.out = (uint4 *)out;
    a.nblocks = len / 16; a.tail = (uint32_t)(len % 16);
    cudaStream_t st = (cudaStream_t)stream;
    switch (ks->rounds) {
    case 10: return (int)launch_ctr_nr<10>(a, st);
    case 12: return (int)launch_ctr_nr<12>(a, st);
    case 14: return (int)launch_ctr_nr<14>(a, st);
    }
    return (int)cudaErrorInvalidValue;
}

int uaes_launch_ecb(const uaes_keysched *ks, int encrypt, int pad, const void *in, void *out, u64 len,
                    void *stream)
{
    if (len == 0 && !(encrypt && pad)) return 0;
    EcbArgs a;
    a.ks = *ks;
    a.in = (const uint4 *)in; a.out = (uint4 *)out;
    a.nblocks = len / 16; a.tail = (uint32_t)(len % 16); a.pad = encrypt ? (uint32_t)pad : 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (ks->rounds * 2 + (encrypt ? 1 : 0)) {
    case 21: return (int)launch_ecb_nr<10, true>(a, st);
    case 20: return (int)launch_ecb_nr<10, false>(a, st);
    case 25: return (int)launch_ecb_nr<12, true>(a, st);
    case 24: return (int)launch_ecb_nr<12, false>(a, st);
    case 29: return (int)launch_ecb_nr<14, true>(a, st);
    case 28: return (int)launch_ecb_nr<14, false>(a, st);
    }
    return (int)cudaErrorInvalidValue;
}

int uaes_launch_fill(u64 seed, u64 first_word, void *dst, u64 nwords, void *stream)
{
    if (nwords == 0) return 0;
    fill_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(seed, first_word, (uint64_t *)dst, nwords);
    ++g_launches;
    return (int)cudaGetLastError();
}

int uaes_launch_xor_fold(const void *src, u64 nwords, void *result_dev, void *stream)
{
    cudaError_t e = cudaMemsetAsync(result_dev, 0, 8, (cudaStream_t)stream);
    if (e != cudaSuccess || nwords == 0) return (int)e;
    xor_fold_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>((const uint64_t *)src, nwords,
                                                                       (unsigned long long *)result_dev);
    ++g_launches;
    return (int)cudaGetLastError();
}

}  // extern "C"
