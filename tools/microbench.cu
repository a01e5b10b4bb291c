// microbench.cu -- instruction micro-roofs that bound the AES kernels on B200:
//   1. conflict-free shared-memory gathers (LDS.32, lane-private banks)      lookups / clk / SM
//   2. texture fetches (tex1Dfetch, 1 KB table, L1 resident)                 lookups / clk / SM
//   3. both together: do the LSU and TEX data paths add up?
//   4. LOP3 and PRMT issue rates                                             lane-ops / clk / SM
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__constant__ uint32_t c_tab[1024];

// LDS gathers next to indexed constant-bank loads (LDC with a per-lane index): does the constant
// cache offer a second lookup path?
template <int NLDS, int NLDC>
__global__ void __launch_bounds__(1024, 1) ldc_kernel(uint32_t *out, int iters, uint32_t smem_bytes)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t base0 = (uint32_t)__cvta_generic_to_shared(dyn);
    const uint32_t base = (base0 + 65535u) & ~65535u;
    if (NLDS) {
        if (base + 65536u > base0 + smem_bytes) __trap();
        for (uint32_t w = threadIdx.x; w < 16384; w += blockDim.x)
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + w * 4), "r"(w * 2654435761u) : "memory");
    }
    __syncthreads();
    uint32_t lb = base + (threadIdx.x & 31) * 4;
    uint32_t s0 = threadIdx.x * 0x9E3779B9u, s1 = s0 ^ 0x12345678u, s2 = s0 + 77, s3 = ~s0;
    for (int it = 0; it < iters; ++it) {
        uint32_t a = 0, b = 0, c = 0, d = 0;
#pragma unroll
        for (int k = 0; k < NLDS; ++k) {
            const uint32_t w = (k & 3) == 0 ? s0 : (k & 3) == 1 ? s1 : (k & 3) == 2 ? s2 : s3;
            const uint32_t ad = __byte_perm(w, lb, 0x7604 | (((k >> 2) & 3) << 4));
            uint32_t r;
            asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(r) : "r"(ad), "n"(0));
            if ((k & 3) == 0) a ^= r; else if ((k & 3) == 1) b ^= r; else if ((k & 3) == 2) c ^= r; else d ^= r;
        }
#pragma unroll
        for (int k = 0; k < NLDC; ++k) {
            const uint32_t w = (k & 3) == 0 ? s1 : (k & 3) == 1 ? s2 : (k & 3) == 2 ? s3 : s0;
            const uint32_t r = c_tab[(w >> (8 * ((k >> 2) & 3))) & 0xff];
            if ((k & 3) == 0) a ^= r; else if ((k & 3) == 1) b ^= r; else if ((k & 3) == 2) c ^= r; else d ^= r;
        }
        s0 = a + it; s1 = b ^ s0; s2 = c + s1; s3 = d ^ s2;
    }
    if ((s0 ^ s1 ^ s2 ^ s3) == 0x1234567) out[1 + (threadIdx.x & 1)] = s0;
}

template <int NLDS, int NLDC>
static void run_ldc(uint32_t *out, int sms, double mhz, uint32_t smem);

template <int NLDS, int NTEX>
__global__ void __launch_bounds__(1024, 1) gather_kernel(cudaTextureObject_t tex, uint32_t *out, int iters, uint32_t smem_bytes)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t base0 = (uint32_t)__cvta_generic_to_shared(dyn);
    const uint32_t base = (base0 + 65535u) & ~65535u;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = base0;
    if (NLDS) {
        if (base + 65536u > base0 + smem_bytes) __trap();
        for (uint32_t w = threadIdx.x; w < 16384; w += blockDim.x)
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + w * 4), "r"(w * 2654435761u) : "memory");
    }
    __syncthreads();
    uint32_t lb = base + (threadIdx.x & 31) * 4;
    uint32_t s0 = threadIdx.x * 0x9E3779B9u, s1 = s0 ^ 0x12345678u, s2 = s0 + 77, s3 = ~s0;
    for (int it = 0; it < iters; ++it) {
        uint32_t a = 0, b = 0, c = 0, d = 0;
#pragma unroll
        for (int k = 0; k < NLDS; ++k) {
            const uint32_t w = (k & 3) == 0 ? s0 : (k & 3) == 1 ? s1 : (k & 3) == 2 ? s2 : s3;
            const uint32_t ad = __byte_perm(w, lb, 0x7604 | (((k >> 2) & 3) << 4));
            uint32_t r;
            asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(r) : "r"(ad), "n"(0));
            if ((k & 3) == 0) a ^= r; else if ((k & 3) == 1) b ^= r; else if ((k & 3) == 2) c ^= r; else d ^= r;
        }
#pragma unroll
        for (int k = 0; k < NTEX; ++k) {
            const uint32_t w = (k & 3) == 0 ? s1 : (k & 3) == 1 ? s2 : (k & 3) == 2 ? s3 : s0;
            const uint32_t r = tex1Dfetch<uint32_t>(tex, (int)((w >> (8 * ((k >> 2) & 3))) & 0xff));
            if ((k & 3) == 0) a ^= r; else if ((k & 3) == 1) b ^= r; else if ((k & 3) == 2) c ^= r; else d ^= r;
        }
        s0 = a + it; s1 = b ^ s0; s2 = c + s1; s3 = d ^ s2;
    }
    if ((s0 ^ s1 ^ s2 ^ s3) == 0x1234567) out[1 + (threadIdx.x & 1)] = s0;
}

template <int KIND>
__global__ void __launch_bounds__(1024, 1) alu_kernel(uint32_t *out, int iters)
{
    uint32_t a = threadIdx.x, b = a * 3 + 1, c = a ^ 0x55, d = a + 9, e = b ^ c, f = c + d, g = a * 7, h = ~a;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (KIND == 0) {        // LOP3: a ^ b ^ c style, 8 independent chains
                a = a ^ b ^ c; b = b ^ c ^ d; c = c ^ d ^ e; d = d ^ e ^ f; e = e ^ f ^ g; f = f ^ g ^ h; g = g ^ h ^ a; h = h ^ a ^ b;
            } else if (KIND == 1) { // PRMT
                a = __byte_perm(a, b, 0x7604); b = __byte_perm(b, c, 0x7614); c = __byte_perm(c, d, 0x7624); d = __byte_perm(d, e, 0x7634);
                e = __byte_perm(e, f, 0x5410); f = __byte_perm(f, g, 0x6521); g = __byte_perm(g, h, 0x7632); h = __byte_perm(h, a, 0x4703);
            } else if (KIND == 2) { // IMAD (fma pipe) next to LOP3 (alu pipe)
                a = a ^ b ^ c; b = b * 0x9E37u + c; c = c ^ d ^ e; d = d * 0x79B9u + e; e = e ^ f ^ g; f = f * 0x85EBu + g; g = g ^ h ^ a; h = h * 0xC2B2u + a;
            } else if (KIND == 3) { // IDP.4A alone: which pipe, what rate?
                a = __dp4a(a, 0x00008000u, b); b = __dp4a(b, 0x00800000u, c); c = __dp4a(c, 0x80000000u, d); d = __dp4a(d, 0x00000080u, e);
                e = __dp4a(e, 0x00008000u, f); f = __dp4a(f, 0x00800000u, g); g = __dp4a(g, 0x80000000u, h); h = __dp4a(h, 0x00000080u, a);
            } else if (KIND == 4) { // IDP.4A next to LOP3
                a = a ^ b ^ c; b = __dp4a(b, 0x00800000u, c); c = c ^ d ^ e; d = __dp4a(d, 0x00000080u, e); e = e ^ f ^ g; f = __dp4a(f, 0x00800000u, g); g = g ^ h ^ a; h = __dp4a(h, 0x00000080u, a);
            } else {                // IDP.4A next to IMAD: same pipe?
                a = a * 0x9E37u + b; b = __dp4a(b, 0x00800000u, c); c = c * 0x79B9u + d; d = __dp4a(d, 0x00000080u, e); e = e * 0x85EBu + f; f = __dp4a(f, 0x00800000u, g); g = g * 0xC2B2u + h; h = __dp4a(h, 0x00000080u, a);
            }
        }
    }
    if ((a ^ b ^ c ^ d ^ e ^ f ^ g ^ h) == 0x1234567) out[0] = a;
}

// One AES-like step per iteration: 16 lookups + 8 three-input XORs, the lookup address built either
// by PRMT (ALU pipe, 64 KiB-aligned x*256 layout) or by IDP.4A (FMA pipe:  b*128 + lane*4  in one
// dot product with a one-hot selector), plus NLOP extra LOP3 per step standing in for a bitsliced
// co-runner: how much ALU is left while the lookup pipe stays saturated?
template <int ADDR, int NLOP>
__global__ void __launch_bounds__(1024, 1) mix_kernel(uint32_t *out, int iters, uint32_t smem_bytes)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t base0 = (uint32_t)__cvta_generic_to_shared(dyn);
    const uint32_t base = (base0 + 65535u) & ~65535u;
    if (base + 65536u > base0 + smem_bytes) __trap();
    for (uint32_t w = threadIdx.x; w < 16384; w += blockDim.x)
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + w * 4), "r"(w * 2654435761u) : "memory");
    __syncthreads();
    const uint32_t lb = base + (threadIdx.x & 31) * 4;
    uint32_t s0 = threadIdx.x * 0x9E3779B9u, s1 = s0 ^ 0x12345678u, s2 = s0 + 77, s3 = ~s0;
    uint32_t e0 = s0 * 3, e1 = s1 * 5, e2 = s2 * 7, e3 = s3 * 11, e4 = s0 * 13, e5 = s1 * 17, e6 = s2 * 19, e7 = s3 * 23;
    for (int it = 0; it < iters; ++it) {
        uint32_t r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t w = (k & 3) == 0 ? s0 : (k & 3) == 1 ? s1 : (k & 3) == 2 ? s2 : s3;
            uint32_t ad;
            if (ADDR == 0) ad = __byte_perm(w, lb, 0x7604 | ((k >> 2) << 4));
            else ad = __dp4a(w, 0x80u << (8 * (k >> 2)), lb);
            asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(r[k]) : "r"(ad), "n"(0));
        }
#pragma unroll
        for (int k = 0; k < NLOP / 8; ++k) {
            e0 = e0 ^ e1 ^ e2; e1 = e1 ^ e2 ^ e3; e2 = e2 ^ e3 ^ e4; e3 = e3 ^ e4 ^ e5;
            e4 = e4 ^ e5 ^ e6; e5 = e5 ^ e6 ^ e7; e6 = e6 ^ e7 ^ e0; e7 = e7 ^ e0 ^ e1;
        }
        s0 = r[0] ^ r[5] ^ r[10] ^ r[15] ^ it; s1 = r[1] ^ r[6] ^ r[11] ^ r[12] ^ 0x55;
        s2 = r[2] ^ r[7] ^ r[8] ^ r[13] ^ 0x77; s3 = r[3] ^ r[4] ^ r[9] ^ r[14] ^ 0x99;
    }
    if ((s0 ^ s1 ^ s2 ^ s3 ^ e0 ^ e1 ^ e2 ^ e3 ^ e4 ^ e5 ^ e6 ^ e7) == 0x1234567) out[1 + (threadIdx.x & 1)] = s0;
}

template <typename F>
static double time_ms(F launch)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
}

template <int NLDS, int NTEX>
static void run_gather(cudaTextureObject_t tex, uint32_t *out, int sms, double mhz, uint32_t smem)
{
    const int iters = 4000;
    CK(cudaFuncSetAttribute(gather_kernel<NLDS, NTEX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double ms = time_ms([&] { gather_kernel<NLDS, NTEX><<<sms, 1024, smem>>>(tex, out, iters, smem); });
    CK(cudaGetLastError());
    const double clk = ms * 1e-3 * mhz * 1e6;
    const double lookups = (double)iters * 1024 * (NLDS + NTEX);
    printf("gather  LDS=%2d TEX=%2d smem=%3u KB: %8.3f ms  %6.2f lookups/clk/SM  (%.1f clk per %d-lookup step per warp-row)\n",
           NLDS, NTEX, smem >> 10, ms, lookups / clk, clk / iters / 32.0, NLDS + NTEX);
}

template <int ADDR, int NLOP>
static void run_mix(uint32_t *out, int sms, double mhz)
{
    const int iters = 4000;
    const uint32_t smem = 130 * 1024;
    CK(cudaFuncSetAttribute(mix_kernel<ADDR, NLOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double ms = time_ms([&] { mix_kernel<ADDR, NLOP><<<sms, 1024, smem>>>(out, iters, smem); });
    CK(cudaGetLastError());
    const double clk = ms * 1e-3 * mhz * 1e6;
    printf("mix     addr=%s +%2d LOP3/step: %8.3f ms  %6.2f lookups/clk/SM  %6.2f extra LOP3/clk/SM  (%.1f clk per step per warp-row)\n",
           ADDR ? "IDP4A" : "PRMT ", NLOP, ms, (double)iters * 1024 * 16 / clk, (double)iters * 1024 * NLOP / clk, clk / iters / 32.0);
}

template <int NLDS, int NLDC>
static void run_ldc(uint32_t *out, int sms, double mhz, uint32_t smem)
{
    const int iters = 2000;
    CK(cudaFuncSetAttribute(ldc_kernel<NLDS, NLDC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double ms = time_ms([&] { ldc_kernel<NLDS, NLDC><<<sms, 1024, smem>>>(out, iters, smem); });
    CK(cudaGetLastError());
    const double clk = ms * 1e-3 * mhz * 1e6;
    printf("gather  LDS=%2d LDC=%2d smem=%3u KB: %8.3f ms  %6.2f lookups/clk/SM  (%.1f clk per %d-lookup step per warp-row)\n",
           NLDS, NLDC, smem >> 10, ms, (double)iters * 1024 * (NLDS + NLDC) / clk, clk / iters / 32.0, NLDS + NLDC);
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, nominal max %.0f MHz (rates below assume the clock sits at max; see nvidia-smi)\n", p.name, p.multiProcessorCount, mhz);
    uint32_t *out, *tab;
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&tab, 1024));
    uint32_t h[256];
    for (int i = 0; i < 256; ++i) h[i] = i * 2654435761u;
    CK(cudaMemcpy(tab, h, 1024, cudaMemcpyHostToDevice));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = tab;
    rd.res.linear.desc = cudaCreateChannelDesc<uint32_t>();
    rd.res.linear.sizeInBytes = 1024;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    const int sms = p.multiProcessorCount;

    run_gather<16, 0>(tex, out, sms, mhz, 227 * 1024);
    uint32_t base0;
    CK(cudaMemcpy(&base0, out, 4, cudaMemcpyDeviceToHost));
    printf("dynamic shared window starts at shared address 0x%x\n", base0);
    run_gather<16, 0>(tex, out, sms, mhz, 130 * 1024);
    run_gather<0, 16>(tex, out, sms, mhz, 1024);
    run_gather<0, 4>(tex, out, sms, mhz, 1024);
    run_gather<16, 2>(tex, out, sms, mhz, 130 * 1024);
    run_gather<16, 4>(tex, out, sms, mhz, 130 * 1024);
    run_gather<16, 8>(tex, out, sms, mhz, 130 * 1024);
    run_gather<12, 4>(tex, out, sms, mhz, 130 * 1024);
    run_gather<16, 4>(tex, out, sms, mhz, 227 * 1024);

    {
        uint32_t hc[1024];
        for (int i = 0; i < 1024; ++i) hc[i] = i * 2246822519u;
        CK(cudaMemcpyToSymbol(c_tab, hc, sizeof hc));
    }
    run_ldc<0, 16>(out, sms, mhz, 1024);
    run_ldc<0, 4>(out, sms, mhz, 1024);
    run_ldc<16, 1>(out, sms, mhz, 227 * 1024);
    run_ldc<16, 2>(out, sms, mhz, 227 * 1024);
    run_ldc<16, 4>(out, sms, mhz, 227 * 1024);
    run_ldc<14, 2>(out, sms, mhz, 227 * 1024);

    const int iters = 20000;
    const char *names[6] = {"LOP3", "PRMT", "LOP3+IMAD", "IDP4A", "LOP3+IDP4A", "IMAD+IDP4A"};
    for (int kind = 0; kind < 6; ++kind) {
        double ms;
        if (kind == 0) ms = time_ms([&] { alu_kernel<0><<<sms, 1024>>>(out, iters); });
        else if (kind == 1) ms = time_ms([&] { alu_kernel<1><<<sms, 1024>>>(out, iters); });
        else if (kind == 2) ms = time_ms([&] { alu_kernel<2><<<sms, 1024>>>(out, iters); });
        else if (kind == 3) ms = time_ms([&] { alu_kernel<3><<<sms, 1024>>>(out, iters); });
        else if (kind == 4) ms = time_ms([&] { alu_kernel<4><<<sms, 1024>>>(out, iters); });
        else ms = time_ms([&] { alu_kernel<5><<<sms, 1024>>>(out, iters); });
        const double clk = ms * 1e-3 * mhz * 1e6;
        printf("alu     %-10s: %8.3f ms  %6.1f lane-ops/clk/SM\n", names[kind], ms, (double)iters * 16 * 8 * 1024 / clk);
    }
    run_mix<0, 0>(out, sms, mhz);  run_mix<1, 0>(out, sms, mhz);
    run_mix<0, 8>(out, sms, mhz);  run_mix<1, 8>(out, sms, mhz);
    run_mix<0, 16>(out, sms, mhz); run_mix<1, 16>(out, sms, mhz);
    run_mix<0, 24>(out, sms, mhz); run_mix<1, 24>(out, sms, mhz);
    run_mix<0, 32>(out, sms, mhz); run_mix<1, 32>(out, sms, mhz);
    run_mix<1, 48>(out, sms, mhz); run_mix<1, 64>(out, sms, mhz);
    return 0;
}
