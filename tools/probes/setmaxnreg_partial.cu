// Does setmaxnreg work for a PARTIAL warpgroup at the end of a CTA (576 threads = 4 full warpgroups
// + 2 warps)?  PTX asks that all threads of a warpgroup execute the same setmaxnreg; the GCM kernel
// with a 2-warp co-runner would rely on the hardware treating the instruction per warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

__global__ void __launch_bounds__(576, 1) k(unsigned *out, int iters)
{
    if (threadIdx.x >= 512) {
        reg_inc<224>();
        unsigned s[160];
#pragma unroll
        for (int i = 0; i < 160; ++i) s[i] = threadIdx.x * 2654435761u + i;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 160; ++i) s[i] = s[i] * 1664525u + s[(i + 37) % 160];
        unsigned acc = 0;
#pragma unroll
        for (int i = 0; i < 160; ++i) acc ^= s[i];
        out[blockIdx.x * 576 + threadIdx.x] = acc;
        return;
    }
    reg_dec<96>();
    unsigned s[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) s[i] = threadIdx.x * 2654435761u + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 48; ++i) s[i] = s[i] * 1664525u + s[(i + 11) % 48];
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < 48; ++i) acc ^= s[i];
    out[blockIdx.x * 576 + threadIdx.x] = acc;
}

int main()
{
    unsigned *d, *h = new unsigned[148 * 576];
    cudaMalloc(&d, 148 * 576 * 4);
    cudaMemset(d, 0, 148 * 576 * 4);
    k<<<148, 576>>>(d, 1000);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch+sync: %s\n", cudaGetErrorString(e));
    cudaMemcpy(h, d, 148 * 576 * 4, cudaMemcpyDeviceToHost);
    // reference on the host for a few threads
    int bad = 0;
    for (int t : {0, 511, 512, 575}) {
        unsigned acc = 0;
        if (t >= 512) {
            unsigned s[160];
            for (int i = 0; i < 160; ++i) s[i] = t * 2654435761u + i;
            for (int it = 0; it < 1000; ++it) for (int i = 0; i < 160; ++i) s[i] = s[i] * 1664525u + s[(i + 37) % 160];
            for (int i = 0; i < 160; ++i) acc ^= s[i];
        } else {
            unsigned s[48];
            for (int i = 0; i < 48; ++i) s[i] = t * 2654435761u + i;
            for (int it = 0; it < 1000; ++it) for (int i = 0; i < 48; ++i) s[i] = s[i] * 1664525u + s[(i + 11) % 48];
            for (int i = 0; i < 48; ++i) acc ^= s[i];
        }
        for (int b : {0, 147}) if (h[b * 576 + t] != acc) { ++bad; printf("mismatch block %d thread %d\n", b, t); }
    }
    printf("partial-warpgroup setmaxnreg: %s\n", bad ? "WRONG RESULTS" : "ok");
    return bad;
}
