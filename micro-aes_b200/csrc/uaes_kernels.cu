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

#include "uaes_core.cuh"
#include "uaes_gf128.cuh"

namespace uaes {

constexpr int kThreads = 1024;
constexpr int kWarpsPerCta = kThreads / 32;
constexpr uint32_t kDynSmem = 227 * 1024;          // everything an SM has; tables are aligned inside

static unsigned long long g_launches = 0;

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
    const uint32_t base = align_table_base(dyn);
    if (base + kEncTableBytes > smem_u32(dyn) + dyn_smem_size()) __trap();
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

struct CtrArgs {
    uaes_keysched ks;
    uint32_t w0, w1, b8;
    uint64_t v0;                 // counter of block 0
    const uint4 *in;
    uint4 *out;
    uint64_t nblocks;            // full blocks
    uint32_t tail;               // len % 16
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
template <int NR, int kCtrThreads>
__global__ void __launch_bounds__(kCtrThreads, 1) ctr_kernel(const __grid_constant__ CtrArgs a)
{
    constexpr int kCtrWarps = kCtrThreads / 32;
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_tables<true>(dyn);
    const uint32_t *rk = a.ks.w;
    const uint32_t lane = threadIdx.x & 31;

    const uint32_t lowoff = (uint32_t)a.v0 & 255u;
    const uint64_t g0 = a.v0 >> 8;
    const uint64_t ngroups = (lowoff + a.nblocks + 255) >> 8;
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
        if (grp < j1 && k >= 0 && (uint64_t)k < a.nblocks) return ld_stream(a.in + k);
        return make_uint4(0, 0, 0, 0);
    };

    uint64_t tag40 = ~0ull, tag16 = ~0ull;
    uint32_t U[4][4], Cp1 = 0, E0 = 0, E1 = 0, E2 = 0, E3 = 0;
    const uint32_t s0 = a.w0 ^ rk[0], s1 = a.w1 ^ rk[1];

    uint4 cur = fetch(j0, 0);
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
        for (int it = 0; it < 4; ++it) {
            const uint4 nxt = it < 3 ? fetch(j, it + 1) : fetch(j + 1, 0);
            const int64_t k = kof(j, it);
            uint32_t t0 = D0 ^ U[it][0], t1 = D1 ^ U[it][1], t2 = D2 ^ U[it][2], t3 = D3 ^ U[it][3];
            enc_finish<NR, 3>(lb, t0, t1, t2, t3, rk, cur.x, cur.y, cur.z, cur.w);
            if (k >= 0 && (uint64_t)k < a.nblocks) st_stream(a.out + k, make_uint4(t0, t1, t2, t3));
            cur = nxt;
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

// ---------------------------------------------------------------- ECB (micro_aes.c:636-680)

struct EcbArgs {
    uaes_keysched ks;
    const uint4 *in;
    uint4 *out;
    uint64_t nblocks;
    uint32_t tail;               // encrypt: zero-padded extra block; decrypt: bytes copied through
};

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

    if (a.tail && blockIdx.x == 0 && threadIdx.x == 0) {
        const uint8_t *x = (const uint8_t *)(a.in + a.nblocks);
        uint8_t *y = (uint8_t *)(a.out + a.nblocks);
        if (ENC) {                                    // padBlock, micro_aes.c:610-621 (zero padding)
            uint32_t s[4] = {0, 0, 0, 0};
            for (uint32_t i = 0; i < a.tail; ++i) s[i >> 2] |= (uint32_t)x[i] << (8 * (i & 3));
            enc_block<NR>(lb, s[0], s[1], s[2], s[3], rk);
            for (uint32_t i = 0; i < 16; ++i) y[i] = (uint8_t)(s[i >> 2] >> (8 * (i & 3)));
        } else {                                      // the memcpy of micro_aes.c:667 leaves them as is
            for (uint32_t i = 0; i < a.tail; ++i) y[i] = x[i];
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

template <int NR, int kCtrThreads>
static cudaError_t launch_ctr_nt(const CtrArgs &a, cudaStream_t st)
{
    constexpr int kCtrWarps = kCtrThreads / 32;
    cudaError_t e = opt_in_smem(ctr_kernel<NR, kCtrThreads>);
    if (e != cudaSuccess) return e;
    const uint64_t ngroups = (((uint32_t)a.v0 & 255u) + a.nblocks + 255) >> 8;
    // a pair of warps per group; at least 4 groups per pair before another CTA is worth its table fill
    const uint64_t ctas = (ngroups + 4 * (kCtrWarps / 2) - 1) / (4 * (kCtrWarps / 2));
    const uint64_t sms = (uint64_t)sm_count();
    ctr_kernel<NR, kCtrThreads><<<(unsigned)(ctas < 1 ? 1 : ctas < sms ? ctas : sms), kCtrThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

// CTA size of the CTR kernel: registers per thread trade against warps per SM (768 threads = 80
// registers, no spills).  UAES_CTR_THREADS overrides it for tuning runs.
template <int NR>
static cudaError_t launch_ctr_nr(const CtrArgs &a, cudaStream_t st)
{
    static int threads = 0;
    if (!threads) {
        const char *e = getenv("UAES_CTR_THREADS");
        threads = e ? atoi(e) : 768;
    }
    switch (threads) {
    case 1024: return launch_ctr_nt<NR, 1024>(a, st);
    case 896:  return launch_ctr_nt<NR, 896>(a, st);
    case 640:  return launch_ctr_nt<NR, 640>(a, st);
    case 512:  return launch_ctr_nt<NR, 512>(a, st);
    default:   return launch_ctr_nt<NR, 768>(a, st);
    }
}

template <int NR, bool ENC>
static cudaError_t launch_ecb_nr(const EcbArgs &a, cudaStream_t st)
{
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

extern "C" {

u64 uaes_launch_count(void) { return g_launches; }

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

int uaes_launch_ecb(const uaes_keysched *ks, int encrypt, const void *in, void *out, u64 len,
                    void *stream)
{
    if (len == 0) return 0;
    EcbArgs a;
    a.ks = *ks;
    a.in = (const uint4 *)in; a.out = (uint4 *)out;
    a.nblocks = len / 16; a.tail = (uint32_t)(len % 16);
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
