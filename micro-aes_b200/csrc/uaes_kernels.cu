// uaes_kernels.cu -- sm_100a kernels for the AES bulk path and their C launchers.
//
// Kernel shape shared by every mode ("Kernels"):
//   * persistent grid: one 1024-thread CTA per SM (148 on B200), 227 KB of dynamic shared memory
//     holding the lane-replicated T-tables (uaes_tables.cuh);
//   * one 16-byte block per thread per step, a warp covers 32 consecutive blocks = 512
//     contiguous bytes, moved with one 128-bit load and one 128-bit store per thread;
//   * the next step's input is requested before the current step's rounds start, so 32 warps
//     keep 16 KB of loads in flight per SM;
//   * round keys are kernel arguments (constant bank), no per-launch symbol copies.
#include <cuda_runtime.h>
#include <stdint.h>

#include "uaes_core.cuh"
#include "uaes_gf128.cuh"

namespace uaes {

constexpr int kThreads = 1024;
constexpr int kWarpsPerCta = kThreads / 32;
constexpr uint32_t kDynSmem = 227 * 1024;          // everything an SM has; tables are aligned inside

static unsigned long long g_launches = 0;

// ---------------------------------------------------------------- small device helpers

__device__ __forceinline__ uint4 ld_stream(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ void st_stream(uint4 *p, uint4 v)
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};"
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
// Inside a group only byte 15 changes, so AddRoundKey(0), 15/16 of round 1 and 12/16 of round 2
// are the same for all 256 blocks: the warp computes them once per group (27 lookups) and every
// block then needs 1 + 4 lookups for rounds 1-2 instead of 32.  Rounds 3..NR are the plain
// 16-lookup rounds.  Lane l takes counters with low byte = 32*it + l, it = 0..7, i.e. 8 coalesced
// 512-byte rows per group.
template <int NR>
__global__ void __launch_bounds__(kThreads, 1) ctr_kernel(const __grid_constant__ CtrArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_tables<true>(dyn);
    const uint32_t *rk = a.ks.w;
    const uint32_t lane = threadIdx.x & 31;

    const uint32_t lowoff = (uint32_t)a.v0 & 255u;
    const uint64_t g0 = a.v0 >> 8;
    const uint64_t ngroups = (lowoff + a.nblocks + 255) >> 8;
    const uint64_t nwarps = (uint64_t)gridDim.x * kWarpsPerCta;
    uint64_t j = (uint64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);

    // block index handled by this lane in row `it` of group j; valid iff 0 <= k < nblocks
    auto kof = [&](uint64_t grp, int it) -> int64_t {
        return (int64_t)(grp << 8) + it * 32 + (int64_t)lane - (int64_t)lowoff;
    };
    auto fetch = [&](uint64_t grp, int it) -> uint4 {
        const int64_t k = kof(grp, it);
        if (grp < ngroups && k >= 0 && (uint64_t)k < a.nblocks) return ld_stream(a.in + k);
        return make_uint4(0, 0, 0, 0);
    };

    uint4 cur = fetch(j, 0);
    for (; j < ngroups; j += nwarps) {
        // ---- per-group constants
        uint32_t w2, w3;
        ctr_words(a.b8, ((g0 + j) << 8) & kMask56, w2, w3);
        const uint32_t s0 = a.w0 ^ rk[0], s1 = a.w1 ^ rk[1], s2 = w2 ^ rk[2], s3 = w3 ^ rk[3];
        const uint32_t K0 = lut<0, kOffT0>(lb, s0) ^ lut<1, kOffT1>(lb, s1) ^ lut<2, kOffT2>(lb, s2) ^ rk[4];
        const uint32_t C1 = lut<0, kOffT0>(lb, s1) ^ lut<1, kOffT1>(lb, s2) ^ lut<2, kOffT2>(lb, s3) ^ lut<3, kOffT3>(lb, s0) ^ rk[5];
        const uint32_t C2 = lut<0, kOffT0>(lb, s2) ^ lut<1, kOffT1>(lb, s3) ^ lut<2, kOffT2>(lb, s0) ^ lut<3, kOffT3>(lb, s1) ^ rk[6];
        const uint32_t C3 = lut<0, kOffT0>(lb, s3) ^ lut<1, kOffT1>(lb, s0) ^ lut<2, kOffT2>(lb, s1) ^ lut<3, kOffT3>(lb, s2) ^ rk[7];
        const uint32_t D0 = lut<1, kOffT1>(lb, C1) ^ lut<2, kOffT2>(lb, C2) ^ lut<3, kOffT3>(lb, C3) ^ rk[8];
        const uint32_t D1 = lut<0, kOffT0>(lb, C1) ^ lut<1, kOffT1>(lb, C2) ^ lut<2, kOffT2>(lb, C3) ^ rk[9];
        const uint32_t D2 = lut<0, kOffT0>(lb, C2) ^ lut<1, kOffT1>(lb, C3) ^ lut<3, kOffT3>(lb, C1) ^ rk[10];
        const uint32_t D3 = lut<0, kOffT0>(lb, C3) ^ lut<2, kOffT2>(lb, C1) ^ lut<3, kOffT3>(lb, C2) ^ rk[11];

#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
            const uint4 nxt = it < 7 ? fetch(j, it + 1) : fetch(j + nwarps, 0);
            const int64_t k = kof(j, it);
            // round 1, column 0: the only column that sees byte 15 of the counter
            const uint32_t c0 = K0 ^ lut<3, kOffT3>(lb, s3 ^ ((uint32_t)(it * 32 + lane) << 24));
            // round 2: one varying byte per column
            uint32_t t0 = D0 ^ lut<0, kOffT0>(lb, c0);
            uint32_t t1 = D1 ^ lut<3, kOffT3>(lb, c0);
            uint32_t t2 = D2 ^ lut<2, kOffT2>(lb, c0);
            uint32_t t3 = D3 ^ lut<1, kOffT1>(lb, c0);
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

template <int NR>
static cudaError_t launch_ctr_nr(const CtrArgs &a, cudaStream_t st)
{
    cudaError_t e = opt_in_smem(ctr_kernel<NR>);
    if (e != cudaSuccess) return e;
    const uint64_t ngroups = (((uint32_t)a.v0 & 255u) + a.nblocks + 255) >> 8;
    ctr_kernel<NR><<<grid_for(ngroups), kThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
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
