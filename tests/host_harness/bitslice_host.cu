// bitslice_host.cu -- TEST INFRASTRUCTURE: runs the host compilation of uaes_bitslice.cuh (the
// bitsliced co-runner of ctr_kernel) for one 1024-counter pass on the CPU, so that
// tests/test_bitslice_host.py can compare it with the oracle without a GPU.
// Build: nvcc -O1 -shared -Xcompiler -fPIC -I micro-aes_b200/csrc -o tests/host_harness/libbitslice_host.so ...
#include <stdint.h>
#include <string.h>
#include "uaes_tables.cuh"
#include "uaes_bitslice.cuh"
#include "uaes_bitslice8.cuh"

using namespace uaes;

static const WordTable kTe0 = make_te0();

static uint32_t te_host(int tbl, uint32_t x)
{
    const uint32_t v = kTe0.v[x & 255];
    return tbl ? (v << (8 * tbl)) | (v >> (32 - 8 * tbl)) : v;
}

// rk: expanded key words (little-endian words of the byte schedule), block0: the 16 counter-block
// bytes of the pass base (counter value multiple of 1024 in bytes 9..15), out: 1024 keystream blocks
extern "C" int bs_host_pass(const uint32_t *rk, int rounds, const uint8_t *block0, uint8_t *out)
{
    static BsKeyPlanes kp;
    bs_make_key_planes(rk, rounds, &kp);
    uint32_t w[4];
    memcpy(w, block0, 16);
    uint32_t uw[6], um[kBsUniformMasks];
    bs_uniform_words(te_host, w[0] ^ rk[0], w[1] ^ rk[1], w[2] ^ rk[2], w[3] ^ rk[3], rk, uw);
    for (int i = 0; i < kBsUniformMasks; ++i) um[i] = bs_mask(uw[i / 32], i % 32);
    const uint32_t c14 = block0[14];
    if ((block0[14] & 3) || block0[15]) return 1;
    for (uint32_t lane = 0; lane < 32; ++lane) {
        uint32_t s[128];
        bs_first_rounds(s, lane, c14, kp.k0, um);
        for (int r = 3; r < rounds; ++r) bs_round(s, kp.k[r - 3]);
        bs_last_round(s, kp.k[rounds - 3]);
        for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
        for (int t = 0; t < 32; ++t)
            for (int c = 0; c < 4; ++c) memcpy(out + 16 * (32 * t + lane) + 4 * c, &s[32 * c + t], 4);
    }
    return 0;
}

// 32 arbitrary blocks (512 bytes) through the general bitsliced path: words -> planes (four 32x32
// transposes), all rounds, planes -> words.  in/out: block t at byte 16*t.
template <int NR>
static void ecb32(const BsKeyPlanesFull &kp, const uint8_t *in, uint8_t *out)
{
    uint32_t s[128];
    for (int t = 0; t < 32; ++t)
        for (int c = 0; c < 4; ++c) memcpy(&s[32 * c + t], in + 16 * t + 4 * c, 4);
    for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
    bs_encrypt_planes<NR>(s, kp);
    for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
    for (int t = 0; t < 32; ++t)
        for (int c = 0; c < 4; ++c) memcpy(out + 16 * t + 4 * c, &s[32 * c + t], 4);
}

extern "C" int bs_host_ecb32(const uint32_t *rk, int rounds, const uint8_t *in, uint8_t *out)
{
    static BsKeyPlanesFull kp;
    bs_make_key_planes_full(rk, rounds, &kp);
    switch (rounds) {
    case 10: ecb32<10>(kp, in, out); return 0;
    case 12: ecb32<12>(kp, in, out); return 0;
    case 14: ecb32<14>(kp, in, out); return 0;
    }
    return 1;
}

// decryption: dk = equivalent-inverse-cipher schedule built here from the encryption schedule
static uint8_t gmul8(uint8_t a, uint8_t b)
{
    uint8_t r = 0;
    for (int i = 0; i < 8; ++i) { if (b & 1) r ^= a; a = (uint8_t)((a << 1) ^ ((a >> 7) * 0x1b)); b >>= 1; }
    return r;
}

static uint32_t inv_mix_word(uint32_t w)
{
    uint8_t a[4] = {(uint8_t)w, (uint8_t)(w >> 8), (uint8_t)(w >> 16), (uint8_t)(w >> 24)}, o[4];
    for (int r = 0; r < 4; ++r)
        o[r] = gmul8(a[r], 14) ^ gmul8(a[(r + 1) & 3], 11) ^ gmul8(a[(r + 2) & 3], 13) ^ gmul8(a[(r + 3) & 3], 9);
    return (uint32_t)o[0] | (uint32_t)o[1] << 8 | (uint32_t)o[2] << 16 | (uint32_t)o[3] << 24;
}

template <int NR>
static void ecb32_dec(const BsKeyPlanesFull &kp, const uint8_t *in, uint8_t *out)
{
    uint32_t s[128];
    for (int t = 0; t < 32; ++t)
        for (int c = 0; c < 4; ++c) memcpy(&s[32 * c + t], in + 16 * t + 4 * c, 4);
    for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
    bs_decrypt_planes<NR>(s, kp);
    for (int c = 0; c < 4; ++c) bs_transpose32(s + 32 * c);
    for (int t = 0; t < 32; ++t)
        for (int c = 0; c < 4; ++c) memcpy(out + 16 * t + 4 * c, &s[32 * c + t], 4);
}

extern "C" int bs_host_ecb32_decrypt(const uint32_t *rk, int rounds, const uint8_t *in, uint8_t *out)
{
    static BsKeyPlanesFull kp;
    uint32_t dk[60];
    for (int c = 0; c < 4; ++c) { dk[c] = rk[4 * rounds + c]; dk[4 * rounds + c] = rk[c]; }
    for (int r = 1; r < rounds; ++r)
        for (int c = 0; c < 4; ++c) dk[4 * r + c] = inv_mix_word(rk[4 * (rounds - r) + c]);
    bs_make_key_planes_full(dk, rounds, &kp);
    switch (rounds) {
    case 10: ecb32_dec<10>(kp, in, out); return 0;
    case 12: ecb32_dec<12>(kp, in, out); return 0;
    case 14: ecb32_dec<14>(kp, in, out); return 0;
    }
    return 1;
}

// The narrow form (uaes_bitslice8.cuh): one pass = one group = the 256 counters sharing bytes 0..14.
// block0: the group's counter block with byte 15 = 0; out: 256 keystream blocks.  Rounds 0-2 are
// factored exactly as the kernels factor them (state entering round 3 = D_j ^ U_j(byte 15)).
extern "C" int bs8_host_group(const uint32_t *rk, int rounds, const uint8_t *block0, uint8_t *out)
{
    static BsKeyPlanes8 kp;
    bs8_make_key_planes(rk, rounds, &kp);
    if (block0[15]) return 1;
    uint32_t w[4];
    memcpy(w, block0, 16);
    const uint32_t s0 = w[0] ^ rk[0], s1 = w[1] ^ rk[1], s2 = w[2] ^ rk[2], s3 = w[3] ^ rk[3];
    auto B = [](uint32_t x, int i) { return (x >> (8 * i)) & 255u; };
    const uint32_t K0 = te_host(0, B(s0, 0)) ^ te_host(1, B(s1, 1)) ^ te_host(2, B(s2, 2)) ^ rk[4];
    const uint32_t C1 = te_host(0, B(s1, 0)) ^ te_host(1, B(s2, 1)) ^ te_host(2, B(s3, 2)) ^ te_host(3, B(s0, 3)) ^ rk[5];
    const uint32_t C2 = te_host(0, B(s2, 0)) ^ te_host(1, B(s3, 1)) ^ te_host(2, B(s0, 2)) ^ te_host(3, B(s1, 3)) ^ rk[6];
    const uint32_t C3 = te_host(0, B(s3, 0)) ^ te_host(1, B(s0, 1)) ^ te_host(2, B(s1, 2)) ^ te_host(3, B(s2, 3)) ^ rk[7];
    const uint32_t D0 = te_host(1, B(C1, 1)) ^ te_host(2, B(C2, 2)) ^ te_host(3, B(C3, 3)) ^ rk[8];
    const uint32_t D1 = te_host(0, B(C1, 0)) ^ te_host(1, B(C2, 1)) ^ te_host(2, B(C3, 2)) ^ rk[9];
    const uint32_t D2 = te_host(3, B(C1, 3)) ^ te_host(0, B(C2, 0)) ^ te_host(1, B(C3, 1)) ^ rk[10];
    const uint32_t D3 = te_host(2, B(C1, 2)) ^ te_host(3, B(C2, 3)) ^ te_host(0, B(C3, 0)) ^ rk[11];
    uint32_t dm[32];
    for (int j = 0; j < 32; ++j) dm[j] = bs8_spread(D0, D1, D2, D3, j);
    for (uint32_t lane = 0; lane < 32; ++lane) {
        uint32_t up[32], s[32];
        for (uint32_t t = 0; t < 8; ++t) {
            const uint32_t c0 = K0 ^ te_host(3, B(s3, 3) ^ (32 * t + lane));
            up[8 * 0 + t] = te_host(0, B(c0, 0)); up[8 * 1 + t] = te_host(3, B(c0, 3));
            up[8 * 2 + t] = te_host(2, B(c0, 2)); up[8 * 3 + t] = te_host(1, B(c0, 1));
        }
        bs_transpose32(up);
        for (int j = 0; j < 32; ++j) s[j] = up[j] ^ dm[j];
        switch (rounds) {
        case 10: bs8_finish<10>(s, kp); break;
        case 12: bs8_finish<12>(s, kp); break;
        case 14: bs8_finish<14>(s, kp); break;
        default: return 2;
        }
        bs_transpose32(s);
        for (int t = 0; t < 8; ++t)
            for (int c = 0; c < 4; ++c) memcpy(out + 16 * (32 * t + lane) + 4 * c, &s[8 * c + t], 4);
    }
    return 0;
}

// 8 arbitrary blocks (128 bytes) through the narrow general form, encrypt (dec = 0) or decrypt (dec = 1)
template <int NR>
static void ecb8(const BsKeyPlanes8Full &kp, const uint8_t *in, uint8_t *out, int dec)
{
    uint32_t s[32];
    for (int t = 0; t < 8; ++t)
        for (int c = 0; c < 4; ++c) memcpy(&s[8 * c + t], in + 16 * t + 4 * c, 4);
    bs_transpose32(s);
    if (dec) bs8_decrypt_planes<NR>(s, kp); else bs8_encrypt_planes<NR>(s, kp);
    bs_transpose32(s);
    for (int t = 0; t < 8; ++t)
        for (int c = 0; c < 4; ++c) memcpy(out + 16 * t + 4 * c, &s[8 * c + t], 4);
}

extern "C" int bs8_host_ecb8(const uint32_t *rk, int rounds, const uint8_t *in, uint8_t *out, int dec)
{
    static BsKeyPlanes8Full kp;
    uint32_t dk[60];
    if (dec) {
        for (int c = 0; c < 4; ++c) { dk[c] = rk[4 * rounds + c]; dk[4 * rounds + c] = rk[c]; }
        for (int r = 1; r < rounds; ++r)
            for (int c = 0; c < 4; ++c) dk[4 * r + c] = inv_mix_word(rk[4 * (rounds - r) + c]);
    }
    bs8_make_key_planes_full(dec ? dk : rk, rounds, &kp);
    switch (rounds) {
    case 10: ecb8<10>(kp, in, out, dec); return 0;
    case 12: ecb8<12>(kp, in, out, dec); return 0;
    case 14: ecb8<14>(kp, in, out, dec); return 0;
    }
    return 1;
}

extern "C" int bs_host_sbox_lut3_count(void) { return kSboxLut3Count; }
