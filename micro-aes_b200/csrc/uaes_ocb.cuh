// uaes_ocb.cuh -- AES-OCB (RFC 7253), included by uaes_kernels.cu.
//
// Restates OCB_cipher (micro_aes.c:1693-1767): Y_i = D_i ^ Cipher(D_i ^ X_i) with the offsets
// D_i = Offset_0 ^ XOR{ L_k : bit k of the Gray code i ^ (i>>1) } -- what the reference's getDelta
// (micro_aes.c:1662-1680, "how to parallelize it by independent calculation of the offset blocks")
// computes by repeated doubling.  Here the L_k = 2^(k+1) * L_$ come from a 64-entry table built
// once per call, a lane jumps to its first offset through the set bits of the Gray code and then
// steps 32 blocks at a time with   D_(i+32) = D_i ^ L_4 ^ L_(5 + ntz((i>>5)+1)).
// Checksum (XOR of all plaintext blocks) is a shuffle reduction plus one atomic XOR per warp; the
// tag = E(checksum ^ D_last ^ L_$) ^ PMAC(associated data) is finished by a one-CTA kernel.
#pragma once

namespace uaes {

struct OcbWork {
    uint4 Lstar, Ldollar, off0;
    uint4 L[64];                 // L[k] = 2^(k+1) * L_$   (doubleBblock, micro_aes.c:433-443)
    uint32_t checksum[4];
};

// 2 * v in GF(2^128), big-endian convention, on the four memory-order words of a block
__device__ inline uint4 ocb_double(uint4 v)
{
    Gf g = gf_from_words(v.x, v.y, v.z, v.w);
    const uint64_t msb = g.hi >> 63;
    g.hi = g.hi << 1 | g.lo >> 63;
    g.lo = (g.lo << 1) ^ (msb ? 0x87ull : 0);
    uint4 r;
    gf_to_words(g, r.x, r.y, r.z, r.w);
    return r;
}

__device__ __forceinline__ void xor4(uint4 &a, const uint4 &b) { a.x ^= b.x; a.y ^= b.y; a.z ^= b.z; a.w ^= b.w; }

// XOR of the L_k selected by the Gray code of index (0 for index 0); L is any address space
__device__ inline uint4 ocb_gray_sum(const uint4 *L, uint64_t index)
{
    uint4 d = make_uint4(0, 0, 0, 0);
    uint64_t gray = index ^ (index >> 1);
    for (int k = 0; gray; ++k, gray >>= 1)
        if (gray & 1) xor4(d, L[k]);
    return d;
}

struct OcbSetupArgs {
    uaes_keysched ks;            // encryption schedule
    uint32_t nonce[3];
    uint32_t taglen;             // OCB_TAG_LEN: enters the nonce block (micro_aes.c:1707)
    OcbWork *work;
};

// getSubkeys + K_top / Stretch / Offset_0 (micro_aes.c:1704-1722), one thread
__global__ void ocb_setup_kernel(const __grid_constant__ OcbSetupArgs a)
{
    if (threadIdx.x) return;
    uint32_t z[4] = {0, 0, 0, 0};
    small_encrypt(a.ks.w, a.ks.rounds, z);
    const uint4 Lstar = make_uint4(z[0], z[1], z[2], z[3]);
    uint4 L = ocb_double(Lstar);
    a.work->Lstar = Lstar;
    a.work->Ldollar = L;
    for (int k = 0; k < 64; ++k) { L = ocb_double(L); a.work->L[k] = L; }
    // nonce block: 00 00 00 01 || nonce with its last six bits cleared; bottom = those six bits
    const uint32_t bottom = (a.nonce[2] >> 24) & 63;
    // byte 0 carries the tag length: kt[0] |= OCB_TAG_LEN << 4 as a uint8_t (micro_aes.c:1707; 16 -> 0)
    uint32_t kt[4] = {0x01000000u | ((a.taglen << 4) & 0xffu), a.nonce[0], a.nonce[1], a.nonce[2] & 0xc0ffffffu};
    small_encrypt(a.ks.w, a.ks.rounds, kt);
    const Gf K = gf_from_words(kt[0], kt[1], kt[2], kt[3]);
    const uint64_t ext = K.hi ^ (K.hi << 8 | K.lo >> 56);     // Stretch = K_top || (K_top[0..8) ^ K_top[1..9))
    Gf o = K;
    if (bottom) {
        o.hi = K.hi << bottom | K.lo >> (64 - bottom);
        o.lo = K.lo << bottom | ext >> (64 - bottom);
    }
    uint4 off0;
    gf_to_words(o, off0.x, off0.y, off0.z, off0.w);
    a.work->off0 = off0;
    a.work->checksum[0] = a.work->checksum[1] = a.work->checksum[2] = a.work->checksum[3] = 0;
}

struct OcbBulkArgs {
    uaes_keysched ks;            // encryption schedule, or the inverse schedule when decrypting
    const uint4 *in;
    uint4 *out;
    uint64_t nblocks;            // full blocks
    OcbWork *work;
};

template <int NR, bool ENC>
__global__ void __launch_bounds__(kThreads, 1) ocb_bulk_kernel(const __grid_constant__ OcbBulkArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    // the L table lives in the last KiB of the dynamic window (the AES tables end >= 35 KiB earlier)
    uint4 *Ls = (uint4 *)(dyn + dyn_smem_size() - 1024);
    if (threadIdx.x < 64) Ls[threadIdx.x] = a.work->L[threadIdx.x];
    const uint32_t lb = setup_tables<ENC>(dyn);               // contains __syncthreads()
    const uint32_t *rk = a.ks.w;
    const uint32_t lane = threadIdx.x & 31;

    const uint64_t rows = (a.nblocks + 31) / 32;
    const uint64_t nwarps = (uint64_t)gridDim.x * kWarpsPerCta;
    const uint64_t rpw = (rows + nwarps - 1) / nwarps;
    const uint64_t warp = (uint64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const uint64_t r0 = warp * rpw, r1 = r0 + rpw < rows ? r0 + rpw : rows;
    uint4 sum = make_uint4(0, 0, 0, 0);

    if (r0 < r1) {
        uint64_t k = r0 * 32 + lane;                           // 0-based block, OCB index i = k + 1
        uint4 delta = a.work->off0;
        xor4(delta, ocb_gray_sum(Ls, k + 1));
        const uint4 L4 = Ls[4];
        uint4 cur = k < a.nblocks ? ld_stream(a.in + k) : make_uint4(0, 0, 0, 0);
        for (uint64_t r = r0; r < r1; ++r, k += 32) {
            const uint4 nxt = (r + 1 < r1 && k + 32 < a.nblocks) ? ld_stream(a.in + k + 32) : make_uint4(0, 0, 0, 0);
            uint32_t s0 = cur.x ^ delta.x, s1 = cur.y ^ delta.y, s2 = cur.z ^ delta.z, s3 = cur.w ^ delta.w;
            if (ENC) enc_block<NR>(lb, s0, s1, s2, s3, rk, delta.x, delta.y, delta.z, delta.w);
            else     dec_block<NR>(lb, s0, s1, s2, s3, rk, delta.x, delta.y, delta.z, delta.w);
            if (k < a.nblocks) {
                st_stream(a.out + k, make_uint4(s0, s1, s2, s3));
                if (ENC) xor4(sum, cur); else xor4(sum, make_uint4(s0, s1, s2, s3));   // plaintext checksum
            }
            // D_(i+32) = D_i ^ L_4 ^ L_(5 + ntz((i >> 5) + 1)),  i = k + 1
            const uint32_t m = 5 + (uint32_t)__ffsll((long long)(((k + 1) >> 5) + 1)) - 1;
            xor4(delta, L4);
            xor4(delta, Ls[m]);
            cur = nxt;
        }
    }
    for (int o = 16; o; o >>= 1) {
        sum.x ^= __shfl_xor_sync(0xffffffffu, sum.x, o); sum.y ^= __shfl_xor_sync(0xffffffffu, sum.y, o);
        sum.z ^= __shfl_xor_sync(0xffffffffu, sum.z, o); sum.w ^= __shfl_xor_sync(0xffffffffu, sum.w, o);
    }
    if (lane == 0 && (sum.x | sum.y | sum.z | sum.w)) {
        atomicXor(&a.work->checksum[0], sum.x); atomicXor(&a.work->checksum[1], sum.y);
        atomicXor(&a.work->checksum[2], sum.z); atomicXor(&a.work->checksum[3], sum.w);
    }
}

// ---- OCB encryption with the co-runner (the shape of ecb_hybrid_kernel): 12 table-driven warps with
// two rows in flight + one warpgroup of bitsliced warps (general form).  Offsets stay in the word
// layout on both sides: a lane jumps to its first Delta through the Gray code and steps 32 blocks at
// a time; the bitsliced warps walk the 32 rows of a tile twice (whitening in, whitening out) from the
// saved first Delta instead of keeping 32 of them.  The checksum is the XOR of the plaintext words as
// they are loaded.
struct OcbHybridArgs {
    OcbBulkArgs o;
    uint64_t tt_blocks;          // blocks [0, tt_blocks): table-driven warps; a multiple of 1024
    BsKeyPlanesFull bs;
};

// table-driven side: UAES_OCB_TT threads with UAES_OCB_ILP rows in flight at UAES_OCB_TT_REGS registers (profiles/r2_sweep_ocb_ilp.txt)
#ifndef UAES_OCB_TT
#define UAES_OCB_TT 384
#endif
#ifndef UAES_OCB_ILP
#define UAES_OCB_ILP 2
#endif
#ifndef UAES_OCB_TT_REGS
#define UAES_OCB_TT_REGS kHybridTtRegs
#endif
constexpr int kOcbTtThreads = UAES_OCB_TT;

// D_(i+32) = D_i ^ L_4 ^ L_(5 + ntz((i >> 5) + 1)),  i = k + 1 (1-based index of the block just done)
__device__ __forceinline__ void ocb_step32(uint4 &delta, const uint4 *Ls, const uint4 &L4, uint64_t k)
{
    const uint32_t m = 5 + (uint32_t)__ffsll((long long)(((k + 1) >> 5) + 1)) - 1;
    xor4(delta, L4);
    xor4(delta, Ls[m]);
}

template <int NR>
__global__ void __launch_bounds__(kOcbTtThreads + kBsThreads, 1) ocb_hybrid_kernel(const __grid_constant__ OcbHybridArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    uint4 *Ls = (uint4 *)(dyn + dyn_smem_size() - 1024);
    if (threadIdx.x < 64) Ls[threadIdx.x] = a.o.work->L[threadIdx.x];
    const uint32_t lb = setup_tables<true>(dyn);                // contains __syncthreads()
    const uint32_t *rk = a.o.ks.w;
    const uint32_t lane = threadIdx.x & 31;
    constexpr int kTtWarps = kOcbTtThreads / 32;
    constexpr int kLaunchRegs = (65536 / (kOcbTtThreads + kBsThreads)) / 8 * 8;
    constexpr int kTtRegs = UAES_OCB_TT_REGS, kBsRegs = (kLaunchRegs + (kLaunchRegs - kTtRegs) * kOcbTtThreads / kBsThreads) / 8 * 8;
    const uint64_t nblocks = a.o.nblocks;
    uint4 sum = make_uint4(0, 0, 0, 0);

    if (threadIdx.x >= kOcbTtThreads) {
        reg_inc<kBsRegs>();
        const uint64_t ntiles = (nblocks - a.tt_blocks + 1023) / 1024;
        const uint64_t gw = (uint64_t)blockIdx.x * (kBsThreads / 32) + ((threadIdx.x - kOcbTtThreads) >> 5);
        const uint64_t nw = (uint64_t)gridDim.x * (kBsThreads / 32);
        const uint64_t per = (ntiles + nw - 1) / nw;
        const uint64_t p0 = gw * per < ntiles ? gw * per : ntiles;
        const uint64_t p1 = p0 + per < ntiles ? p0 + per : ntiles;
        for (uint64_t tile = p0; tile < p1; ++tile) {
            const uint64_t kb = a.tt_blocks + tile * 1024 + lane;
            uint32_t s[128];
            {   // whitening in; nothing of this phase but the checksum stays live across the rounds
                const uint4 L4 = Ls[4];
                uint4 delta = a.o.work->off0;
                xor4(delta, ocb_gray_sum(Ls, kb + 1));
#pragma unroll
                for (int tb = 0; tb < 32; tb += 8) {
                    uint4 v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = kb + 32 * (tb + i) < nblocks ? ld_stream(a.o.in + kb + 32 * (tb + i)) : make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        xor4(sum, v[i]);                             // out-of-range rows are zero
                        s[tb + i] = v[i].x ^ delta.x; s[32 + tb + i] = v[i].y ^ delta.y;
                        s[64 + tb + i] = v[i].z ^ delta.z; s[96 + tb + i] = v[i].w ^ delta.w;
                        ocb_step32(delta, Ls, L4, kb + 32 * (tb + i));
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
            bs_encrypt_planes<NR, false>(s, a.bs);
#pragma unroll
            for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
            {   // whitening out: the offsets are walked again from the block index (opaque to the compiler,
                // so that the first walk's values are not kept in registers across the rounds)
                uint64_t kb2 = kb;
                uint32_t four = 4;
                asm volatile("" : "+l"(kb2), "+r"(four));
                const uint4 L4 = Ls[four];
                uint4 delta = a.o.work->off0;
                xor4(delta, ocb_gray_sum(Ls, kb2 + 1));
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    if (kb2 + 32 * t < nblocks)
                        st_stream(a.o.out + kb2 + 32 * t, make_uint4(s[t] ^ delta.x, s[32 + t] ^ delta.y, s[64 + t] ^ delta.z, s[96 + t] ^ delta.w));
                    ocb_step32(delta, Ls, L4, kb2 + 32 * t);
                }
            }
        }
    } else {
        reg_dec<kTtRegs>();
#if UAES_OCB_ILP == 1
        // one row in flight per thread (the 16-warp geometry of xts_sectors_hybrid_kernel / ecb_dec_hybrid_kernel)
        const uint64_t nrows = a.tt_blocks / 32;
        const uint64_t gw = (uint64_t)blockIdx.x * kTtWarps + (threadIdx.x >> 5);
        const uint64_t nw = (uint64_t)gridDim.x * kTtWarps;
        const uint64_t per = (nrows + nw - 1) / nw;
        const uint64_t q0 = gw * per < nrows ? gw * per : nrows;
        const uint64_t q1 = q0 + per < nrows ? q0 + per : nrows;
        if (q0 < q1) {
            const uint4 L4 = Ls[4];
            uint4 dl[1];
            dl[0] = a.o.work->off0;
            xor4(dl[0], ocb_gray_sum(Ls, q0 * 32 + lane + 1));
            uint4 cur = ld_stream(a.o.in + q0 * 32 + lane), nxt = make_uint4(0, 0, 0, 0);
            for (uint64_t q = q0; q < q1; ++q) {
                const uint64_t k = q * 32 + lane;
                if (q + 1 < q1) nxt = ld_stream(a.o.in + k + 32);
                xor4(sum, cur);
                uint32_t st[1][4] = {{cur.x ^ dl[0].x ^ rk[0], cur.y ^ dl[0].y ^ rk[1], cur.z ^ dl[0].z ^ rk[2], cur.w ^ dl[0].w ^ rk[3]}};
                enc_finish_n<NR, 1, 1>(lb, st, rk, dl);
                st_stream(a.o.out + k, make_uint4(st[0][0], st[0][1], st[0][2], st[0][3]));
                ocb_step32(dl[0], Ls, L4, k);
                cur = nxt;
            }
        }
#else
        const uint64_t npairs = a.tt_blocks / 64;
        const uint64_t gw = (uint64_t)blockIdx.x * kTtWarps + (threadIdx.x >> 5);
        const uint64_t nw = (uint64_t)gridDim.x * kTtWarps;
        const uint64_t per = (npairs + nw - 1) / nw;
        const uint64_t q0 = gw * per < npairs ? gw * per : npairs;
        const uint64_t q1 = q0 + per < npairs ? q0 + per : npairs;
        if (q0 < q1) {
            const uint4 L4 = Ls[4];
            uint4 cur[2], nxt[2], dl[2];
            dl[0] = a.o.work->off0;
            xor4(dl[0], ocb_gray_sum(Ls, q0 * 64 + lane + 1));
            cur[0] = ld_stream(a.o.in + q0 * 64 + lane); cur[1] = ld_stream(a.o.in + q0 * 64 + 32 + lane);
            for (uint64_t q = q0; q < q1; ++q) {
                const uint64_t k = q * 64 + lane;
                if (q + 1 < q1) { nxt[0] = ld_stream(a.o.in + k + 64); nxt[1] = ld_stream(a.o.in + k + 96); }
                dl[1] = dl[0];
                ocb_step32(dl[1], Ls, L4, k);
                uint32_t st[2][4];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    xor4(sum, cur[i]);
                    st[i][0] = cur[i].x ^ dl[i].x ^ rk[0]; st[i][1] = cur[i].y ^ dl[i].y ^ rk[1];
                    st[i][2] = cur[i].z ^ dl[i].z ^ rk[2]; st[i][3] = cur[i].w ^ dl[i].w ^ rk[3];
                }
                enc_finish_n<NR, 1, 2>(lb, st, rk, dl);
                st_stream(a.o.out + k, make_uint4(st[0][0], st[0][1], st[0][2], st[0][3]));
                st_stream(a.o.out + k + 32, make_uint4(st[1][0], st[1][1], st[1][2], st[1][3]));
                dl[0] = dl[1];
                ocb_step32(dl[0], Ls, L4, k + 32);
                cur[0] = nxt[0]; cur[1] = nxt[1];
            }
        }
#endif
    }
    for (int o = 16; o; o >>= 1) {
        sum.x ^= __shfl_xor_sync(0xffffffffu, sum.x, o); sum.y ^= __shfl_xor_sync(0xffffffffu, sum.y, o);
        sum.z ^= __shfl_xor_sync(0xffffffffu, sum.z, o); sum.w ^= __shfl_xor_sync(0xffffffffu, sum.w, o);
    }
    if (lane == 0 && (sum.x | sum.y | sum.z | sum.w)) {
        atomicXor(&a.o.work->checksum[0], sum.x); atomicXor(&a.o.work->checksum[1], sum.y);
        atomicXor(&a.o.work->checksum[2], sum.z); atomicXor(&a.o.work->checksum[3], sum.w);
    }
}

template <int NR>
static cudaError_t launch_ocb_hybrid_nr(const OcbBulkArgs &o, uint64_t bs_blocks, cudaStream_t st)
{
    cudaError_t e = opt_in_smem(ocb_hybrid_kernel<NR>);
    if (e != cudaSuccess) return e;
    static thread_local OcbHybridArgs a;                 // 8 KB of planes: off the stack, one per calling thread
    a.o = o;
    a.tt_blocks = (o.nblocks - bs_blocks) & ~1023ull;
    bs_make_key_planes_full(o.ks.w, NR, &a.bs);
    const uint64_t need = (o.nblocks + 32 * 16 - 1) / (32 * 16), sms = (uint64_t)sm_count();
    ocb_hybrid_kernel<NR><<<(unsigned)(need < sms ? need : sms), kOcbTtThreads + kBsThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

struct OcbFinishArgs {
    uaes_keysched ks;            // encryption schedule
    const uint8_t *in;
    uint8_t *out;
    uint64_t len;
    const uint8_t *aad;
    uint64_t aadlen;
    int encrypt;
    uint8_t *tag_out;
    uint32_t taglen;             // OCB_TAG_LEN
    OcbWork *work;
};

// ragged tail, PMAC over the associated data (block-parallel across the CTA), tag
template <int NR>
__global__ void __launch_bounds__(kThreads, 1) ocb_finish_kernel(const __grid_constant__ OcbFinishArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    uint4 *red = (uint4 *)(dyn + dyn_smem_size() - 1024);     // 32 x 16 B reduction scratch
    const uint32_t lb = setup_tables<true>(dyn);
    const uint32_t *rk = a.ks.w;
    const uint4 *L = a.work->L;
    const uint4 Lstar = a.work->Lstar;

    // PMAC: sum_i E(A_i ^ gray_sum(i)), i = 1..m, plus the padded partial block (micro_aes.c:1750-1765)
    const uint64_t m = a.aadlen / 16;
    const uint32_t ar = (uint32_t)(a.aadlen % 16);
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (uint64_t j = threadIdx.x; j < m; j += kThreads) {
        uint4 b = load_block_bytes(a.aad + 16 * j, 16);
        xor4(b, ocb_gray_sum(L, j + 1));
        enc_block<NR>(lb, b.x, b.y, b.z, b.w, rk);
        xor4(acc, b);
    }
    if (threadIdx.x == 0 && ar) {
        uint4 b = load_block_bytes(a.aad + 16 * m, ar);
        b.x ^= ar < 4 ? 0x80u << (8 * ar) : 0;                 // 10* padding right after the data
        b.y ^= ar >= 4 && ar < 8 ? 0x80u << (8 * (ar - 4)) : 0;
        b.z ^= ar >= 8 && ar < 12 ? 0x80u << (8 * (ar - 8)) : 0;
        b.w ^= ar >= 12 ? 0x80u << (8 * (ar - 12)) : 0;
        xor4(b, ocb_gray_sum(L, m));
        xor4(b, Lstar);
        enc_block<NR>(lb, b.x, b.y, b.z, b.w, rk);
        xor4(acc, b);
    }
    for (int o = 16; o; o >>= 1) {
        acc.x ^= __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y ^= __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z ^= __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w ^= __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x) return;
    uint4 pmac = make_uint4(0, 0, 0, 0);
    for (int w = 0; w < kWarpsPerCta; ++w) xor4(pmac, red[w]);

    // D_n, the ragged tail (micro_aes.c:1738-1743) and the checksum
    const uint64_t n = a.len / 16;
    const uint32_t r = (uint32_t)(a.len % 16);
    uint4 delta = a.work->off0;
    xor4(delta, ocb_gray_sum(L, n));
    uint4 sum = make_uint4(a.work->checksum[0], a.work->checksum[1], a.work->checksum[2], a.work->checksum[3]);
    if (r) {
        xor4(delta, Lstar);                                    // Offset_* = D_n ^ L_*
        uint4 pad = delta;
        enc_block<NR>(lb, pad.x, pad.y, pad.z, pad.w, rk);
        const uint4 x = load_block_bytes(a.in + 16 * n, r);
        const uint32_t keep[4] = {r >= 4 ? 0xffffffffu : (1u << (8 * r)) - 1,
                                  r >= 8 ? 0xffffffffu : r > 4 ? (1u << (8 * (r - 4))) - 1 : 0,
                                  r >= 12 ? 0xffffffffu : r > 8 ? (1u << (8 * (r - 8))) - 1 : 0,
                                  r > 12 ? (1u << (8 * (r - 12))) - 1 : 0};
        const uint32_t yw[4] = {(x.x ^ pad.x) & keep[0], (x.y ^ pad.y) & keep[1], (x.z ^ pad.z) & keep[2], (x.w ^ pad.w) & keep[3]};
        store_bytes(a.out + 16 * n, yw, r);
        uint4 p = a.encrypt ? x : make_uint4(yw[0], yw[1], yw[2], yw[3]);   // the PLAINtext tail
        p.x ^= r < 4 ? 0x80u << (8 * r) : 0;
        p.y ^= r >= 4 && r < 8 ? 0x80u << (8 * (r - 4)) : 0;
        p.z ^= r >= 8 && r < 12 ? 0x80u << (8 * (r - 8)) : 0;
        p.w ^= r >= 12 ? 0x80u << (8 * (r - 12)) : 0;
        xor4(sum, p);
    }
    // tag = E(checksum ^ D ^ L_$) ^ PMAC   (cMac(Ld, NULL, del, 16, tag), micro_aes.c:1746)
    xor4(sum, delta);
    xor4(sum, a.work->Ldollar);
    enc_block<NR>(lb, sum.x, sum.y, sum.z, sum.w, rk);
    xor4(sum, pmac);
    const uint32_t tw[4] = {sum.x, sum.y, sum.z, sum.w};
    store_bytes(a.tag_out, tw, a.taglen);                   // micro_aes.c:1783
}

template <int NR, bool ENC>
static cudaError_t launch_ocb_bulk_nr(const OcbBulkArgs &a, cudaStream_t st)
{
    if (ENC) {                                           // encryption of enough data: with the co-runner
        ctr_tuning_init();
        const int share = g_ctr_share != kCtrDefaultShare ? g_ctr_share : env_int("UAES_OCB_BS_PERMILLE", 185);
        if (g_ctr_share > 0 && share > 0 && (long long)a.nblocks >= g_ctr_bs_min && a.nblocks >= 2048)
            return launch_ocb_hybrid_nr<NR>(a, a.nblocks / 1024 * (uint64_t)share, st);
    }
    cudaError_t e = opt_in_smem(ocb_bulk_kernel<NR, ENC>);
    if (e != cudaSuccess) return e;
    ocb_bulk_kernel<NR, ENC><<<grid_for((a.nblocks + 255) / 256), kThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

template <int NR>
static cudaError_t launch_ocb_finish_nr(const OcbFinishArgs &a, cudaStream_t st)
{
    cudaError_t e = opt_in_smem(ocb_finish_kernel<NR>);
    if (e != cudaSuccess) return e;
    ocb_finish_kernel<NR><<<1, kThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace uaes

// enc = encryption schedule; bulk = enc (encrypt) or the inverse schedule (decrypt).  aad_dev and
// work are device memory (work >= uaes_ocb_work_bytes()); the tag goes to tag_out (device).
extern "C" size_t uaes_ocb_work_bytes(void) { return sizeof(uaes::OcbWork); }

extern "C" int uaes_launch_ocb(const uaes_keysched *enc, const uaes_keysched *bulk, int encrypt,
                               const unsigned char nonce[12], const void *aad_dev, u64 aadlen,
                               const void *in, void *out, u64 len, void *tag_out, unsigned taglen, void *work, void *stream)
{
    using namespace uaes;
    cudaStream_t st = (cudaStream_t)stream;
    OcbSetupArgs s;
    s.ks = *enc;
    for (int c = 0; c < 3; ++c)
        s.nonce[c] = (uint32_t)nonce[4 * c] | (uint32_t)nonce[4 * c + 1] << 8 | (uint32_t)nonce[4 * c + 2] << 16 | (uint32_t)nonce[4 * c + 3] << 24;
    s.taglen = taglen;
    s.work = (OcbWork *)work;
    ocb_setup_kernel<<<1, 32, 0, st>>>(s);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;

    if (len / 16) {
        OcbBulkArgs b;
        b.ks = *bulk;
        b.in = (const uint4 *)in; b.out = (uint4 *)out; b.nblocks = len / 16; b.work = (OcbWork *)work;
        switch (enc->rounds * 2 + (encrypt ? 1 : 0)) {
        case 21: e = launch_ocb_bulk_nr<10, true>(b, st); break;
        case 20: e = launch_ocb_bulk_nr<10, false>(b, st); break;
        case 25: e = launch_ocb_bulk_nr<12, true>(b, st); break;
        case 24: e = launch_ocb_bulk_nr<12, false>(b, st); break;
        case 29: e = launch_ocb_bulk_nr<14, true>(b, st); break;
        case 28: e = launch_ocb_bulk_nr<14, false>(b, st); break;
        default: e = cudaErrorInvalidValue;
        }
        if (e != cudaSuccess) return (int)e;
    }
    OcbFinishArgs f;
    f.ks = *enc;
    f.in = (const uint8_t *)in; f.out = (uint8_t *)out; f.len = len;
    f.aad = (const uint8_t *)aad_dev; f.aadlen = aadlen; f.encrypt = encrypt;
    f.tag_out = (uint8_t *)tag_out; f.taglen = taglen; f.work = (OcbWork *)work;
    switch (enc->rounds) {
    case 10: return (int)launch_ocb_finish_nr<10>(f, st);
    case 12: return (int)launch_ocb_finish_nr<12>(f, st);
    case 14: return (int)launch_ocb_finish_nr<14>(f, st);
    }
    return (int)cudaErrorInvalidValue;
}
