// ghash_host.cu -- TEST INFRASTRUCTURE: runs the host compilation of the GCM kernel's
// multiply-by-constant (ghash_mul_table / ghash_table_entry, micro-aes_b200/csrc/uaes_gf128.cuh)
// on the CPU, so that tests/test_ghash_host.py can compare it with the oracle's mulGF128
// (micro_aes.c:476-493) without a GPU.  The device instance differs only in where the table entry
// comes from (an LDS.128 from the replicated shared-memory table instead of this array).
// Build: nvcc -O1 -shared -Xcompiler -fPIC -I micro-aes_b200/csrc -o tests/host_harness/libghash_host.so ...
#include <stdint.h>
#include <string.h>
#include "uaes_gf128.cuh"

using namespace uaes;

static uint32_t be32(const uint8_t *p)
{
    return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
}

static void put_be32(uint8_t *p, uint32_t v)
{
    p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}

struct HostFetch {
    const uint4 *M;
    __host__ __device__ uint4 operator()(uint32_t word, int byte) const { return M[(word >> (8 * byte)) & 255u]; }
};

// out_i = y_i * C for n blocks y_i, everything as 16 bytes in GCM order; table_out (optional) receives
// the 256 entries as 16 bytes each in GCM order
extern "C" int ghash_host_mul(const uint8_t *C, const uint8_t *y, int n, uint8_t *out, uint8_t *table_out)
{
    Gf c;
    c.hi = (uint64_t)be32(C) << 32 | be32(C + 4);
    c.lo = (uint64_t)be32(C + 8) << 32 | be32(C + 12);
    static uint4 M[256];
    for (uint32_t b = 0; b < 256; ++b) {
        M[b] = ghash_table_entry(c, b);
        if (table_out) {
            put_be32(table_out + 16 * b, M[b].x); put_be32(table_out + 16 * b + 4, M[b].y);
            put_be32(table_out + 16 * b + 8, M[b].z); put_be32(table_out + 16 * b + 12, M[b].w);
        }
    }
    const HostFetch fetch{M};
    for (int i = 0; i < n; ++i) {
        uint32_t y0 = be32(y + 16 * i), y1 = be32(y + 16 * i + 4), y2 = be32(y + 16 * i + 8), y3 = be32(y + 16 * i + 12);
        ghash_mul_table(fetch, y0, y1, y2, y3);
        put_be32(out + 16 * i, y0); put_be32(out + 16 * i + 4, y1);
        put_be32(out + 16 * i + 8, y2); put_be32(out + 16 * i + 12, y3);
    }
    return 0;
}
