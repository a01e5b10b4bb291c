// uaes_chain.cuh -- the block-parallel DEcrypt directions of CBC and CFB (SURVEY.md 8f row 2;
// included by uaes_kernels.cu).
//
//   CBC decrypt (micro_aes.c:746-782):  P_k = D_K(C_k) ^ C_(k-1),  C_(-1) = IV, with the
//                                        reference's default CS3 ciphertext stealing for the
//                                        last two blocks (CTS = 1, micro_aes.h:55-57)
//   CFB decrypt (micro_aes.c:799-845):  P_k = E_K(C_(k-1)) ^ C_k,  ragged tail via mixThenXor
//
// Every output block depends on two INPUT blocks only, so both are as parallel as ECB; the
// neighbour's ciphertext block comes from a second (cache-resident) 128-bit load.  The encrypt
// directions are serial chains and are not provided.  in and out must not overlap (the
// reference chains through the input buffer as well, micro_aes.c:766).
#pragma once

namespace uaes {

struct ChainArgs {
    uaes_keysched ks;            // CBC: inverse schedule; CFB: encryption schedule
    uaes_keysched kse;           // CBC: encryption-order schedule for the one-thread CTS pair
    uint32_t iv[4];
    const uint4 *in;
    uint4 *out;
    uint64_t nblocks;            // blocks handled by the plain loop
    uint32_t tail;               // CBC: r of the CTS pair (1..16, 0 = none); CFB: len % 16
    uint32_t tail_only;          // the nblocks whole blocks were done by another kernel (ecb_dec_hybrid_kernel): only the tail
};

__device__ inline void store_bytes(uint8_t *y, const uint32_t w[4], uint32_t n)
{
    for (uint32_t i = 0; i < n; ++i) y[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
}

template <int NR, bool CBC>
__global__ void __launch_bounds__(kThreads, 1) chain_dec_kernel(const __grid_constant__ ChainArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    const uint32_t lb = setup_tables<!CBC>(dyn);             // CBC needs the inverse tables
    const uint32_t *rk = a.ks.w;
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    const uint4 iv = make_uint4(a.iv[0], a.iv[1], a.iv[2], a.iv[3]);

    for (uint64_t k = (uint64_t)blockIdx.x * kThreads + threadIdx.x; k < (a.tail_only ? 0 : a.nblocks); k += stride) {
        const uint4 cur = ld_stream(a.in + k);
        const uint4 prev = k ? a.in[k - 1] : iv;             // the neighbour lane loads it too: L1 hit
        uint32_t s0, s1, s2, s3;
        if (CBC) {
            s0 = cur.x; s1 = cur.y; s2 = cur.z; s3 = cur.w;
            dec_block<NR>(lb, s0, s1, s2, s3, rk, prev.x, prev.y, prev.z, prev.w);
        } else {
            s0 = prev.x; s1 = prev.y; s2 = prev.z; s3 = prev.w;
            enc_block<NR>(lb, s0, s1, s2, s3, rk, cur.x, cur.y, cur.z, cur.w);
        }
        st_stream(a.out + k, make_uint4(s0, s1, s2, s3));
    }

    if (a.tail && blockIdx.x == 0 && threadIdx.x == 0) {
        const uint64_t m = a.nblocks;
        const uint8_t *x = (const uint8_t *)(a.in + m);
        uint8_t *y = (uint8_t *)(a.out + m);
        const uint4 chain = m ? load_block_bytes((const uint8_t *)(a.in + m - 1), 16) : iv;
        if (CBC) {
            // CS3 pair {X, Z}: P2 = Z ^ Dec(X) (r bytes), P1 = chain ^ Dec(Z | tail of Dec(X))
            const uint32_t r = a.tail;
            const uint4 X = load_block_bytes(x, 16), Z = load_block_bytes(x + 16, r);
            uint32_t dx[4] = {X.x, X.y, X.z, X.w};
            small_decrypt(a.kse.w, a.kse.rounds, dx);
            const uint32_t zw[4] = {Z.x, Z.y, Z.z, Z.w};
            uint32_t p2[4], blk[4];
            for (int c = 0; c < 4; ++c) p2[c] = dx[c] ^ zw[c];
            for (int c = 0; c < 4; ++c) {                        // Z's r bytes, then Dec(X)'s remaining ones
                const uint32_t lo = 4 * c, keep = r >= lo + 4 ? 0xffffffffu : r > lo ? (1u << (8 * (r - lo))) - 1 : 0;
                blk[c] = (zw[c] & keep) | (dx[c] & ~keep);
            }
            small_decrypt(a.kse.w, a.kse.rounds, blk);
            blk[0] ^= chain.x; blk[1] ^= chain.y; blk[2] ^= chain.z; blk[3] ^= chain.w;
            store_bytes(y, blk, 16);
            store_bytes(y + 16, p2, r);
        } else {
            uint32_t s[4] = {chain.x, chain.y, chain.z, chain.w};
            small_encrypt(a.ks.w, a.ks.rounds, s);
            const uint4 C = load_block_bytes(x, a.tail);
            s[0] ^= C.x; s[1] ^= C.y; s[2] ^= C.z; s[3] ^= C.w;
            store_bytes(y, s, a.tail);
        }
    }
}

template <int NR, bool CBC>
static cudaError_t launch_chain_nr(const ChainArgs &a, cudaStream_t st)
{
    if (!CBC) {                                          // CFB decryption of enough data: ECB-shaped, with the co-runner
        ctr_tuning_init();
        const int share = g_ctr_share != kCtrDefaultShare ? g_ctr_share : env_int("UAES_ECB_BS_PERMILLE", kCfbDefaultShare);
        if (g_ctr_share > 0 && share > 0 && (long long)a.nblocks >= g_ctr_bs_min && a.nblocks >= 2048) {
            EcbArgs e0;
            e0.ks = a.ks; e0.in = a.in; e0.out = a.out; e0.nblocks = a.nblocks; e0.tail = a.tail; e0.pad = 0;
            return launch_ecb_hybrid_nr<NR, true>(e0, a.nblocks / 1024 * (uint64_t)share, st, a.iv);
        }
    }
    if (CBC) {                                           // CBC decryption of enough data: ECB-decrypt-shaped, with the inverse-cipher co-runner
        EcbArgs e0;
        e0.ks = a.ks; e0.in = a.in; e0.out = a.out; e0.nblocks = a.nblocks; e0.tail = 0; e0.pad = 0;
        bool done = false;
        const cudaError_t eh = launch_ecb_dec_hybrid_nr<NR, true>(e0, a.iv, st, done);
        if (done) {
            if (eh != cudaSuccess || !a.tail) return eh;
            ChainArgs t = a;                             // the CS3 pair: one thread of a one-CTA launch
            t.tail_only = 1;
            cudaError_t e = opt_in_smem(chain_dec_kernel<NR, CBC>);
            if (e != cudaSuccess) return e;
            chain_dec_kernel<NR, CBC><<<1, kThreads, kDynSmem, st>>>(t);
            ++g_launches;
            return cudaGetLastError();
        }
    }
    cudaError_t e = opt_in_smem(chain_dec_kernel<NR, CBC>);
    if (e != cudaSuccess) return e;
    chain_dec_kernel<NR, CBC><<<grid_for((a.nblocks + 31) / 32), kThreads, kDynSmem, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace uaes

// ks = inverse schedule (CBC) or encryption schedule (CFB); kse = encryption schedule (CBC tail).
// nblocks/tail are decided by the host, which applies the reference's CS3 rules.
extern "C" int uaes_launch_chain_dec(const uaes_keysched *ks, const uaes_keysched *kse, int cbc,
                                     const unsigned char iv[16], const void *in, void *out, u64 nblocks,
                                     unsigned tail, void *stream)
{
    using namespace uaes;
    if (nblocks == 0 && tail == 0) return 0;
    ChainArgs a;
    a.ks = *ks; a.kse = *kse;
    for (int c = 0; c < 4; ++c)
        a.iv[c] = (uint32_t)iv[4 * c] | (uint32_t)iv[4 * c + 1] << 8 | (uint32_t)iv[4 * c + 2] << 16 | (uint32_t)iv[4 * c + 3] << 24;
    a.in = (const uint4 *)in; a.out = (uint4 *)out; a.nblocks = nblocks; a.tail = tail; a.tail_only = 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (ks->rounds * 2 + (cbc ? 1 : 0)) {
    case 21: return (int)launch_chain_nr<10, true>(a, st);
    case 20: return (int)launch_chain_nr<10, false>(a, st);
    case 25: return (int)launch_chain_nr<12, true>(a, st);
    case 24: return (int)launch_chain_nr<12, false>(a, st);
    case 29: return (int)launch_chain_nr<14, true>(a, st);
    case 28: return (int)launch_chain_nr<14, false>(a, st);
    }
    return (int)cudaErrorInvalidValue;
}
