/*
 * uaes_host.c -- host side of libuaes_b200.so, plain C.
 *
 * What stays on the host is what the reference does ONCE per call: KeyExpansion
 * (micro_aes.c:144-178), building the initial counter block (micro_aes.c:962-971) and argument
 * checks with the reference's return codes.  Every per-block operation -- Cipher(), the CTR /
 * XTS / GCM chaining, GHASH, even E_K(0) and E_K(J0) -- runs in the CUDA kernels of
 * uaes_kernels.cu.  There is no CPU implementation of the data path in this library: without a
 * CUDA device every entry point fails with UAES_E_NO_DEVICE.
 *
 * Buffers are classified per call (cudaPointerGetAttributes):
 *   device / managed, 16-byte aligned  -> kernels run directly on them, zero copies;
 *   anything else (pageable or pinned host memory, misaligned device memory) -> staged through
 *   three device chunks on three streams so that H2D, kernel and D2H of consecutive chunks
 *   overlap (CTR, ECB, XTS sectors), or through one full-size device buffer (GCM, single XTS
 *   data unit, which cannot be cut).
 */
#include <cuda_runtime_api.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/uaes_b200.h"
#include "uaes_launch.h"

#if defined(__GNUC__)
#define UAES_TLS __thread
#else
#define UAES_TLS
#endif

#define MAX_SLOT    8
#define MAX_CHUNK   ((size_t)256 << 20)
#define MAX_DEV     64

/* staging geometry: g_nslot device chunks of g_chunk bytes, one stream each (defaults measured on
 * B200/PCIe gen5, see profiles/); UAES_STAGE_SLOTS / UAES_STAGE_CHUNK_MIB override for tuning */
static int    g_nslot = 3;
static size_t g_chunk = (size_t)64 << 20;
#define NSLOT       g_nslot
#define CHUNK_BYTES g_chunk

typedef unsigned char u8;

/* ------------------------------------------------------------------ error latch */

static UAES_TLS int  tls_err;
static UAES_TLS char tls_msg[160];
static UAES_TLS void *tls_stream;
static UAES_TLS int  tls_async;

static int fail(int code, const char *what, int cuda_err)
{
    tls_err = code;
    if (cuda_err)
        snprintf(tls_msg, sizeof tls_msg, "%s: %s", what, cudaGetErrorString((cudaError_t)cuda_err));
    else
        snprintf(tls_msg, sizeof tls_msg, "%s", what);
    return code;
}

int uaes_last_error(void) { return tls_err; }
const char *uaes_last_error_string(void) { return tls_err ? tls_msg : ""; }
void uaes_clear_error(void) { tls_err = 0; tls_msg[0] = 0; }
void uaes_set_stream(void *stream) { tls_stream = stream; }
void uaes_set_async(int enable) { tls_async = enable; }
uaes_u64 uaes_kernel_launches(void) { return uaes_launch_count(); }
void uaes_ctr_tuning(int tt_threads, int bs_permille, long long bs_min_blocks)
{
    uaes_launch_ctr_tuning(tt_threads, bs_permille, bs_min_blocks);
}

int uaes_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void *uaes_host_alloc(size_t bytes)
{
    void *p = NULL;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        fail(UAES_E_NO_MEMORY, "cudaHostAlloc", (int)cudaGetLastError());
        return NULL;
    }
    return p;
}

void uaes_host_free(void *p) { if (p) cudaFreeHost(p); }

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail(UAES_E_CUDA, #call, (int)e_); goto done; } } while (0)
#define LAUNCH(call) do { int e_ = (call); if (e_) { rc = fail(UAES_E_CUDA, #call, e_); goto done; } } while (0)

/* ------------------------------------------------------------------ key schedules */

static u8 g_sbox[256];
static pthread_once_t g_sbox_once = PTHREAD_ONCE_INIT;

static u8 xtime(u8 a) { return (u8)((a << 1) ^ ((a >> 7) * 0x1b)); }   /* micro_aes.c:115-118 */

static u8 gmul(u8 a, u8 b)
{
    u8 r = 0;
    while (b) { if (b & 1) r ^= a; a = xtime(a); b >>= 1; }
    return r;
}

/* S(a) = affine(a^-1), FIPS-197 5.1.1 (the values of micro_aes.c:41-51) */
static void build_sbox(void)
{
    int a, b;
    for (a = 0; a < 256; ++a) {
        u8 inv = 0, s;
        if (a) for (b = 1; b < 256; ++b) if (gmul((u8)a, (u8)b) == 1) { inv = (u8)b; break; }
        s = (u8)(inv ^ (u8)(inv << 1 | inv >> 7) ^ (u8)(inv << 2 | inv >> 6) ^
                 (u8)(inv << 3 | inv >> 5) ^ (u8)(inv << 4 | inv >> 4) ^ 0x63);
        g_sbox[a] = s;
    }
}

/* KeyExpansion (micro_aes.c:144-178) into little-endian column words */
static int expand_key(int keybits, const u8 *key, uaes_keysched *ks)
{
    int nk, total, i;
    u8 rk[240], rcon = 1;

    if (keybits != 128 && keybits != 192 && keybits != 256) return -1;
    pthread_once(&g_sbox_once, build_sbox);
    nk = keybits / 32;
    total = 4 * (nk + 7);
    memcpy(rk, key, (size_t)(4 * nk));
    for (i = nk; i < total; ++i) {
        u8 t0 = rk[4 * i - 4], t1 = rk[4 * i - 3], t2 = rk[4 * i - 2], t3 = rk[4 * i - 1];
        if (i % nk == 0) {
            const u8 first = t0;
            t0 = (u8)(g_sbox[t1] ^ rcon); t1 = g_sbox[t2]; t2 = g_sbox[t3]; t3 = g_sbox[first];
            rcon = xtime(rcon);
        } else if (nk == 8 && i % nk == 4) {                  /* micro_aes.c:165-172 */
            t0 = g_sbox[t0]; t1 = g_sbox[t1]; t2 = g_sbox[t2]; t3 = g_sbox[t3];
        }
        rk[4 * i + 0] = (u8)(rk[4 * (i - nk) + 0] ^ t0);
        rk[4 * i + 1] = (u8)(rk[4 * (i - nk) + 1] ^ t1);
        rk[4 * i + 2] = (u8)(rk[4 * (i - nk) + 2] ^ t2);
        rk[4 * i + 3] = (u8)(rk[4 * (i - nk) + 3] ^ t3);
    }
    memset(ks, 0, sizeof *ks);
    ks->rounds = nk + 6;
    for (i = 0; i < total; ++i)
        ks->w[i] = (u32)rk[4 * i] | (u32)rk[4 * i + 1] << 8 | (u32)rk[4 * i + 2] << 16 | (u32)rk[4 * i + 3] << 24;
    return 0;
}

/* InvMixColumns of one column word (micro_aes.c:301-312) */
static u32 inv_mix_word(u32 w)
{
    const u8 a0 = (u8)w, a1 = (u8)(w >> 8), a2 = (u8)(w >> 16), a3 = (u8)(w >> 24);
    const u8 b0 = (u8)(gmul(a0, 14) ^ gmul(a1, 11) ^ gmul(a2, 13) ^ gmul(a3, 9));
    const u8 b1 = (u8)(gmul(a0, 9) ^ gmul(a1, 14) ^ gmul(a2, 11) ^ gmul(a3, 13));
    const u8 b2 = (u8)(gmul(a0, 13) ^ gmul(a1, 9) ^ gmul(a2, 14) ^ gmul(a3, 11));
    const u8 b3 = (u8)(gmul(a0, 11) ^ gmul(a1, 13) ^ gmul(a2, 9) ^ gmul(a3, 14));
    return (u32)b0 | (u32)b1 << 8 | (u32)b2 << 16 | (u32)b3 << 24;
}

/* The reference decrypts with the encryption schedule read backwards (micro_aes.c:315-332).
 * The table-driven kernel uses the equivalent inverse cipher, which needs InvMixColumns applied
 * to round keys 1..rounds-1; both produce the same plaintext (FIPS-197 5.3.5). */
static void invert_schedule(const uaes_keysched *enc, uaes_keysched *dec)
{
    int r, c;
    const int nr = enc->rounds;
    memset(dec, 0, sizeof *dec);
    dec->rounds = nr;
    for (r = 0; r <= nr; ++r)
        for (c = 0; c < 4; ++c) {
            const u32 w = enc->w[4 * (nr - r) + c];
            dec->w[4 * r + c] = (r == 0 || r == nr) ? w : inv_mix_word(w);
        }
}

/* counter block of keystream block `first` for a 12-byte IV whose counter field starts at
 * `start` (1 for CTR: micro_aes.c:968-971) */
static void make_ctrblock(const u8 *iv, u64 start, u64 first, uaes_ctrblock *cb)
{
    u64 v = (u64)iv[9] << 48 | (u64)iv[10] << 40 | (u64)iv[11] << 32;
    v ^= start;                                   /* xorBEint into a zeroed field */
    cb->w0 = (u32)iv[0] | (u32)iv[1] << 8 | (u32)iv[2] << 16 | (u32)iv[3] << 24;
    cb->w1 = (u32)iv[4] | (u32)iv[5] << 8 | (u32)iv[6] << 16 | (u32)iv[7] << 24;
    cb->b8 = iv[8];
    cb->v0 = (v + first) & (((u64)1 << 56) - 1);  /* 56-bit carry, micro_aes.c:421-427 */
}

/* ------------------------------------------------------------------ per-device resources */

typedef struct {
    int ready;
    void *slot[MAX_SLOT];
    cudaStream_t st[MAX_SLOT];
    void *big;  size_t big_bytes;       /* grow-only full-size staging (GCM / XTS unit) */
    void *work; size_t work_bytes;      /* grow-only GCM scratch + AAD copy */
} devctx;

static devctx g_dev[MAX_DEV];
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

static int get_ctx(devctx **out)
{
    int dev = 0, i, n = 0;
    devctx *c;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(UAES_E_NO_DEVICE, "no CUDA device: libuaes_b200 has no CPU fallback", 0);
    }
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV)
        return fail(UAES_E_NO_DEVICE, "cudaGetDevice failed", (int)cudaGetLastError());
    c = &g_dev[dev];
    if (!c->ready) {
        const char *e;
        if ((e = getenv("UAES_STAGE_SLOTS")) != NULL && atoi(e) >= 1 && atoi(e) <= MAX_SLOT) g_nslot = atoi(e);
        if ((e = getenv("UAES_STAGE_CHUNK_MIB")) != NULL && atoi(e) >= 1 && (size_t)atoi(e) <= (MAX_CHUNK >> 20))
            g_chunk = (size_t)atoi(e) << 20;
        for (i = 0; i < MAX_SLOT; ++i)
            if (cudaStreamCreateWithFlags(&c->st[i], cudaStreamNonBlocking) != cudaSuccess)
                return fail(UAES_E_CUDA, "cudaStreamCreate", (int)cudaGetLastError());
        c->ready = 1;
    }
    *out = c;
    return 0;
}

static int need_slots(devctx *c)
{
    int i;
    for (i = 0; i < NSLOT; ++i)
        if (!c->slot[i] && cudaMalloc(&c->slot[i], CHUNK_BYTES + 16) != cudaSuccess)
            return fail(UAES_E_NO_MEMORY, "cudaMalloc(staging chunk)", (int)cudaGetLastError());
    return 0;
}

static int grow(void **p, size_t *have, size_t want, const char *what)
{
    if (*have >= want) return 0;
    if (*p) { cudaFree(*p); *p = NULL; *have = 0; }
    want += want / 8 + 4096;
    if (cudaMalloc(p, want) != cudaSuccess) return fail(UAES_E_NO_MEMORY, what, (int)cudaGetLastError());
    *have = want;
    return 0;
}

/* device (or managed) memory that the kernels may touch directly with 128-bit accesses */
static int is_direct(const void *p)
{
    struct cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) return 0;
    return ((size_t)p & 15) == 0;
}

static int finish_direct(void)
{
    if (!tls_async) {
        cudaError_t e = cudaStreamSynchronize((cudaStream_t)tls_stream);
        if (e != cudaSuccess) return fail(UAES_E_CUDA, "cudaStreamSynchronize", (int)e);
    }
    return 0;
}

/* ------------------------------------------------------------------ chunked staging pipeline */

typedef int (*chunk_fn)(void *user, u64 offset, void *dev, size_t bytes, size_t out_bytes, void *stream);

/* in/out may be any mix of host and device memory.  `unit` = granularity a chunk must respect.
 * out_extra = bytes the LAST chunk writes beyond its input size (ECB padding). */
static int run_chunked(devctx *c, const void *in, void *out, size_t len, size_t unit, size_t out_extra,
                       chunk_fn fn, void *user)
{
    int rc = 0, i;
    size_t off = 0, chunk = CHUNK_BYTES - CHUNK_BYTES % unit;
    unsigned n = 0;

    if ((rc = need_slots(c)) != 0) return rc;
    /* inputs produced on the caller's stream (mixed host/device calls) must be complete */
    CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
    while (off < len) {
        const size_t bytes = len - off < chunk ? len - off : chunk;
        const size_t obytes = bytes + (off + bytes == len ? out_extra : 0);
        const int s = (int)(n++ % NSLOT);
        CU(cudaMemcpyAsync(c->slot[s], (const u8 *)in + off, bytes, cudaMemcpyDefault, c->st[s]));
        rc = fn(user, off, c->slot[s], bytes, obytes, c->st[s]);
        if (rc) goto done;
        CU(cudaMemcpyAsync((u8 *)out + off, c->slot[s], obytes, cudaMemcpyDefault, c->st[s]));
        off += bytes;
    }
done:
    for (i = 0; i < NSLOT; ++i) {
        cudaError_t e = cudaStreamSynchronize(c->st[i]);
        if (e != cudaSuccess && !rc) rc = fail(UAES_E_CUDA, "cudaStreamSynchronize(staging)", (int)e);
    }
    return rc;
}

/* ------------------------------------------------------------------ CTR */

typedef struct { uaes_keysched ks; const u8 *iv; u64 first; } ctr_job;

static int ctr_chunk(void *user, u64 offset, void *dev, size_t bytes, size_t obytes, void *stream)
{
    ctr_job *j = (ctr_job *)user;
    uaes_ctrblock cb;
    int e;
    (void)obytes;
    make_ctrblock(j->iv, 1, j->first + offset / 16, &cb);
    e = uaes_launch_ctr(&j->ks, &cb, dev, dev, bytes, stream);
    return e ? fail(UAES_E_CUDA, "ctr kernel launch", e) : 0;
}

int uaes_ctr_crypt_range(int keybits, const uaes_u8 *key, const uaes_u8 *iv, uaes_u64 first_block,
                         const void *in, size_t len, void *out)
{
    devctx *c;
    ctr_job j;
    int rc;
    if (expand_key(keybits, key, &j.ks)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    if (len == 0) return 0;                         /* NULL data is fine when there is none */
    pthread_mutex_lock(&g_lock);
    if ((rc = get_ctx(&c)) != 0) goto done;
    j.iv = iv; j.first = first_block;
    if (is_direct(in) && is_direct(out)) {
        uaes_ctrblock cb;
        make_ctrblock(iv, 1, first_block, &cb);
        LAUNCH(uaes_launch_ctr(&j.ks, &cb, in, out, len, tls_stream));
        rc = finish_direct();
    } else {
        rc = run_chunked(c, in, out, len, 16, 0, ctr_chunk, &j);
    }
done:
    pthread_mutex_unlock(&g_lock);
    return rc;
}

int uaes_ctr_crypt(int keybits, const uaes_u8 *key, const uaes_u8 *iv, const void *in, size_t len, void *out)
{
    return uaes_ctr_crypt_range(keybits, key, iv, 0, in, len, out);
}

/* ------------------------------------------------------------------ ECB */

typedef struct { uaes_keysched ks; int encrypt; } ecb_job;

static int ecb_chunk(void *user, u64 offset, void *dev, size_t bytes, size_t obytes, void *stream)
{
    ecb_job *j = (ecb_job *)user;
    int e;
    (void)offset; (void)obytes;
    e = uaes_launch_ecb(&j->ks, j->encrypt, dev, dev, bytes, stream);
    return e ? fail(UAES_E_CUDA, "ecb kernel launch", e) : 0;
}

static int ecb_common(int keybits, const u8 *key, const void *in, size_t len, void *out, int encrypt)
{
    devctx *c;
    ecb_job j;
    uaes_keysched enc;
    int rc;
    if (expand_key(keybits, key, &enc)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    if (encrypt) j.ks = enc; else invert_schedule(&enc, &j.ks);
    j.encrypt = encrypt;
    if (len == 0) return 0;
    pthread_mutex_lock(&g_lock);
    if ((rc = get_ctx(&c)) != 0) goto done;
    if (is_direct(in) && is_direct(out)) {
        LAUNCH(uaes_launch_ecb(&j.ks, encrypt, in, out, len, tls_stream));
        rc = finish_direct();
    } else {
        /* encrypt pads the ragged tail to a whole block: the last chunk returns up to 15 more bytes */
        const size_t extra = (encrypt && len % 16) ? 16 - len % 16 : 0;
        rc = run_chunked(c, in, out, len, 16, extra, ecb_chunk, &j);
    }
done:
    pthread_mutex_unlock(&g_lock);
    return rc;
}

int uaes_ecb_encrypt(int keybits, const uaes_u8 *key, const void *in, size_t len, void *out)
{
    return ecb_common(keybits, key, in, len, out, 1);
}

int uaes_ecb_decrypt(int keybits, const uaes_u8 *key, const void *in, size_t len, void *out)
{
    const int rc = ecb_common(keybits, key, in, len, out, 0);
    if (rc) return rc;
    return len % 16 ? UAES_DECRYPTION_ERROR : UAES_OK;       /* micro_aes.c:679 */
}

/* ------------------------------------------------------------------ XTS */

typedef struct { uaes_keysched k1, k1e, k2; int encrypt; u64 first_sector; size_t sector_bytes; } xts_job;

static int xts_keys(int keybits, const u8 *keys, int encrypt, xts_job *j)
{
    if (keybits != 128 && keybits != 256)
        return fail(UAES_E_BAD_ARGUMENT, "XTS is defined for 128- and 256-bit keys only", 0);
    expand_key(keybits, keys, &j->k1e);                       /* key1 = first half: data key   */
    expand_key(keybits, keys + keybits / 8, &j->k2);          /* key2 = second half: tweak key */
    if (encrypt) j->k1 = j->k1e; else invert_schedule(&j->k1e, &j->k1);
    j->encrypt = encrypt;
    return 0;
}

static int xts_chunk(void *user, u64 offset, void *dev, size_t bytes, size_t obytes, void *stream)
{
    xts_job *j = (xts_job *)user;
    int e;
    (void)obytes;
    e = uaes_launch_xts_sectors(&j->k1, &j->k2, j->encrypt, j->first_sector + offset / j->sector_bytes,
                                j->sector_bytes / 16, bytes / j->sector_bytes, dev, dev, stream);
    return e ? fail(UAES_E_CUDA, "xts kernel launch", e) : 0;
}

int uaes_xts_sectors(int keybits, const uaes_u8 *keys, uaes_u64 first_sector, size_t sector_bytes,
                     const void *in, size_t len, void *out, int encrypt)
{
    devctx *c;
    xts_job j;
    int rc;
    if ((rc = xts_keys(keybits, keys, encrypt, &j)) != 0) return rc;
    if (sector_bytes < 16 || sector_bytes % 16 || sector_bytes > CHUNK_BYTES || len % sector_bytes)
        return fail(UAES_E_BAD_ARGUMENT, "sector size must be a multiple of 16 and divide the length", 0);
    if (len == 0) return 0;
    j.first_sector = first_sector; j.sector_bytes = sector_bytes;
    pthread_mutex_lock(&g_lock);
    if ((rc = get_ctx(&c)) != 0) goto done;
    if (is_direct(in) && is_direct(out)) {
        LAUNCH(uaes_launch_xts_sectors(&j.k1, &j.k2, encrypt, first_sector, sector_bytes / 16,
                                       len / sector_bytes, in, out, tls_stream));
        rc = finish_direct();
    } else {
        rc = run_chunked(c, in, out, len, sector_bytes, 0, xts_chunk, &j);
    }
done:
    pthread_mutex_unlock(&g_lock);
    return rc;
}

static int xts_unit(int keybits, const u8 *keys, const u8 *tweak, const void *in, size_t len, void *out,
                    int encrypt)
{
    static const u8 sector0[16] = {0};
    devctx *c;
    xts_job j;
    int rc;
    if (len < 16) return UAES_DATALENGTH_ERROR;               /* micro_aes.c:1069, 1088 */
    if ((rc = xts_keys(keybits, keys, encrypt, &j)) != 0) return rc;
    if (!tweak) tweak = sector0;                              /* micro_aes.c:1017-1021 */
    pthread_mutex_lock(&g_lock);
    if ((rc = get_ctx(&c)) != 0) goto done;
    if (is_direct(in) && is_direct(out)) {
        LAUNCH(uaes_launch_xts_unit(&j.k1, &j.k1e, &j.k2, encrypt, tweak, in, out, len, tls_stream));
        rc = finish_direct();
    } else {
        /* one data unit cannot be cut at chunk borders without the tweak chain: stage it whole */
        if ((rc = grow(&c->big, &c->big_bytes, len + 16, "cudaMalloc(XTS unit staging)")) != 0) goto done;
        CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
        CU(cudaMemcpyAsync(c->big, in, len, cudaMemcpyDefault, c->st[0]));
        LAUNCH(uaes_launch_xts_unit(&j.k1, &j.k1e, &j.k2, encrypt, tweak, c->big, c->big, len, c->st[0]));
        CU(cudaMemcpyAsync(out, c->big, len, cudaMemcpyDefault, c->st[0]));
        CU(cudaStreamSynchronize(c->st[0]));
    }
done:
    pthread_mutex_unlock(&g_lock);
    return rc;
}

int uaes_xts_encrypt(int keybits, const uaes_u8 *keys, const uaes_u8 *tweak, const void *in, size_t len, void *out)
{
    return xts_unit(keybits, keys, tweak, in, len, out, 1);
}

int uaes_xts_decrypt(int keybits, const uaes_u8 *keys, const uaes_u8 *tweak, const void *in, size_t len, void *out)
{
    return xts_unit(keybits, keys, tweak, in, len, out, 0);
}

/* ------------------------------------------------------------------ GCM */

#define GCM_WORK_HEAD 1024   /* tag scratch lives in front of the kernels' work area */
#define AAD_STATE_OFF 128    /* 16 bytes inside the head: GHASH state of a bulk-hashed AAD */
#define AAD_BULK_MIN  4096   /* larger AADs are hashed by the bulk kernel instead of one lane */

/* common part: returns with data resident on the device (din/dout), AAD on the device, work ready */
static int gcm_common(int keybits, const u8 *key, const u8 *nonce, const void *aad, size_t aadlen,
                      const void *in, size_t len, void *out, int decrypt)
{
    devctx *c;
    uaes_keysched ks;
    int rc, direct;
    const void *din, *daad;
    void *dout;
    u8 *work, *dtag, *dstate;
    size_t wbytes;
    cudaStream_t st;

    if (expand_key(keybits, key, &ks)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    pthread_mutex_lock(&g_lock);
    if ((rc = get_ctx(&c)) != 0) goto done;

    /* the kernels write ciphertext and tag through `out` and read `in`: both must be device memory
     * for the zero-copy path (an empty message still has a tag to write when encrypting) */
    if (decrypt) direct = len == 0 || (is_direct(in) && is_direct(out));
    else         direct = is_direct(out) && (len == 0 || is_direct(in));
    st = direct ? (cudaStream_t)tls_stream : c->st[0];
    wbytes = GCM_WORK_HEAD + uaes_gcm_work_bytes(len > aadlen ? len : aadlen) + aadlen + 64;
    if ((rc = grow(&c->work, &c->work_bytes, wbytes, "cudaMalloc(GCM work)")) != 0) goto done;
    dtag = (u8 *)c->work;
    work = (u8 *)c->work + GCM_WORK_HEAD;
    daad = NULL;
    if (aadlen) {
        u8 *a = work + uaes_gcm_work_bytes(len > aadlen ? len : aadlen);
        a += (16 - ((size_t)a & 15)) & 15;
        CU(cudaMemcpyAsync(a, aad, aadlen, cudaMemcpyDefault, st));
        daad = a;
    }
    if (direct) {
        din = in; dout = out;
    } else {
        if ((rc = grow(&c->big, &c->big_bytes, len + 32, "cudaMalloc(GCM staging)")) != 0) goto done;
        CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
        CU(cudaMemcpyAsync(c->big, in, len + (decrypt ? 16 : 0), cudaMemcpyDefault, st));
        din = c->big; dout = c->big;
    }
    dstate = NULL;
    if (aadlen >= AAD_BULK_MIN) {                             /* xMac over the AAD (micro_aes.c:1134) in bulk */
        dstate = (u8 *)c->work + AAD_STATE_OFF;
        LAUNCH(uaes_launch_gcm(&ks, nonce, NULL, 0, NULL, daad, NULL, aadlen, 1, 0, 1, dstate, work, st));
    }

    if (!decrypt) {
        /* one fused pass: CTR + GHASH, tag appended at out + len (micro_aes.c:1168,1178) */
        LAUNCH(uaes_launch_gcm(&ks, nonce, daad, aadlen, dstate, din, dout, len, 0, 0, 0, (u8 *)dout + len, work, st));
        if (!direct) CU(cudaMemcpyAsync(out, c->big, len + 16, cudaMemcpyDefault, st));
        if (!direct || !tls_async) CU(cudaStreamSynchronize(st));
    } else {
        /* verify first, decrypt only on success; `out` stays untouched otherwise (micro_aes.c:1199-1209) */
        u8 t1[16], t2[16];
        uaes_ctrblock cb;
        LAUNCH(uaes_launch_gcm(&ks, nonce, daad, aadlen, dstate, din, NULL, len, 1, 0, 0, dtag, work, st));
        CU(cudaMemcpyAsync(t1, dtag, 16, cudaMemcpyDefault, st));
        CU(cudaMemcpyAsync(t2, (const u8 *)din + len, 16, cudaMemcpyDefault, st));
        CU(cudaStreamSynchronize(st));
        if (memcmp(t1, t2, 16)) { rc = UAES_AUTH_ERROR; goto done; }
        if (len) {
            make_ctrblock(nonce, 1, 1, &cb);                  /* J0 = nonce || 1, data from J0 + 1 */
            LAUNCH(uaes_launch_ctr(&ks, &cb, din, dout, len, st));
            if (!direct) CU(cudaMemcpyAsync(out, c->big, len, cudaMemcpyDefault, st));
            if (!direct || !tls_async) CU(cudaStreamSynchronize(st));
        }
    }
done:
    pthread_mutex_unlock(&g_lock);
    return rc;
}

int uaes_gcm_encrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return gcm_common(keybits, key, nonce, aad, aadlen, in, len, out, 0);
}

int uaes_gcm_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return gcm_common(keybits, key, nonce, aad, aadlen, in, len, out, 1);
}

/* ---- a GCM message sharded over several GPUs / calls ---------------------------
 * Every shard runs the fused CTR+GHASH pass over its own byte range and returns 16 bytes; one
 * caller gathers them (an all-gather of 16 B per rank) and folds them into the tag. */
int uaes_gcm_shard(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, uaes_u64 first_block,
                   const void *in, size_t len, void *out, int decrypt, uaes_u8 *partial)
{
    devctx *c;
    uaes_keysched ks;
    int rc, direct;
    const void *din;
    void *dout;
    u8 *work, *dpart;
    cudaStream_t st;

    if (expand_key(keybits, key, &ks)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    pthread_mutex_lock(&g_lock);
    if ((rc = get_ctx(&c)) != 0) goto done;
    direct = len == 0 || (is_direct(in) && is_direct(out));
    st = direct ? (cudaStream_t)tls_stream : c->st[0];
    if ((rc = grow(&c->work, &c->work_bytes, GCM_WORK_HEAD + uaes_gcm_work_bytes(len) + 64, "cudaMalloc(GCM work)")) != 0) goto done;
    dpart = (u8 *)c->work;
    work = (u8 *)c->work + GCM_WORK_HEAD;
    if (direct) {
        din = in; dout = out;
    } else {
        if ((rc = grow(&c->big, &c->big_bytes, len + 32, "cudaMalloc(GCM staging)")) != 0) goto done;
        CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
        CU(cudaMemcpyAsync(c->big, in, len, cudaMemcpyDefault, st));
        din = c->big; dout = c->big;
    }
    LAUNCH(uaes_launch_gcm(&ks, nonce, NULL, 0, NULL, din, dout, len, decrypt ? 2 : 0, first_block, 1, dpart, work, st));
    if (!direct) CU(cudaMemcpyAsync(out, c->big, len, cudaMemcpyDefault, st));
    CU(cudaMemcpyAsync(partial, dpart, 16, cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));                            /* the 16 bytes are a host result */
done:
    pthread_mutex_unlock(&g_lock);
    return rc;
}

int uaes_gcm_combine(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const uaes_u8 *partials, const uaes_u64 *blocks_after, int nshards,
                     uaes_u64 total_len, uaes_u8 *tag)
{
    devctx *c;
    uaes_keysched ks;
    int rc;
    u8 *w;
    cudaStream_t st = (cudaStream_t)tls_stream;

    if (expand_key(keybits, key, &ks)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    if (nshards < 0 || nshards > 31) return fail(UAES_E_BAD_ARGUMENT, "at most 31 shards", 0);
    pthread_mutex_lock(&g_lock);
    if ((rc = get_ctx(&c)) != 0) goto done;
    if ((rc = grow(&c->work, &c->work_bytes, GCM_WORK_HEAD + 32 * 24 + aadlen + 64, "cudaMalloc(GCM work)")) != 0) goto done;
    w = (u8 *)c->work;                                        /* [tag 16][pad][partials][after][aad] */
    if (nshards) {
        CU(cudaMemcpyAsync(w + 64, partials, (size_t)nshards * 16, cudaMemcpyDefault, st));
        CU(cudaMemcpyAsync(w + 64 + 32 * 16, blocks_after, (size_t)nshards * 8, cudaMemcpyDefault, st));
    }
    if (aadlen) CU(cudaMemcpyAsync(w + GCM_WORK_HEAD, aad, aadlen, cudaMemcpyDefault, st));
    LAUNCH(uaes_launch_gcm_combine(&ks, nonce, aadlen ? w + GCM_WORK_HEAD : NULL, aadlen, total_len,
                                   w + 64, w + 64 + 32 * 16, (unsigned)nshards, w, st));
    CU(cudaMemcpyAsync(tag, w, 16, cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
done:
    pthread_mutex_unlock(&g_lock);
    return rc;
}

/* ------------------------------------------------------------------ GCM-SIV */

/* micro_aes.c:1474-1516.  Two passes by construction (the tag is the CTR seed): encrypt =
 * POLYVAL over the plaintext, tag, then CTR; decrypt = CTR seeded by the received tag, POLYVAL
 * over the result, compare (like the reference, the plaintext is written before the check). */
static int gcmsiv_common(int keybits, const u8 *key, const u8 *nonce, const void *aad, size_t aadlen,
                         const void *in, size_t len, void *out, int decrypt)
{
    devctx *c;
    uaes_keysched master, enc;
    int rc, direct;
    const void *din, *daad;
    void *dout;
    u8 *work, *dtag, *dderived, *dstate, derived[48];
    size_t wbytes;
    cudaStream_t st;

    if (expand_key(keybits, key, &master)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    pthread_mutex_lock(&g_lock);
    if ((rc = get_ctx(&c)) != 0) goto done;

    if (decrypt) direct = len == 0 || (is_direct(in) && is_direct(out));
    else         direct = is_direct(out) && (len == 0 || is_direct(in));
    st = direct ? (cudaStream_t)tls_stream : c->st[0];
    wbytes = GCM_WORK_HEAD + uaes_gcm_work_bytes(len > aadlen ? len : aadlen) + aadlen + 64;
    if ((rc = grow(&c->work, &c->work_bytes, wbytes, "cudaMalloc(GCM-SIV work)")) != 0) goto done;
    dtag = (u8 *)c->work;                 /* [0,16) computed tag, [64,112) derived key material */
    dderived = (u8 *)c->work + 64;
    work = (u8 *)c->work + GCM_WORK_HEAD;

    /* message keys: E_K(LE32(i) || nonce) on the device, KeyExpansion of the result on the host */
    LAUNCH(uaes_launch_gcmsiv_derive(&master, nonce, dderived, st));
    CU(cudaMemcpyAsync(derived, dderived, 48, cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
    expand_key(keybits, derived + 16, &enc);

    daad = NULL;
    if (aadlen) {
        u8 *a = work + uaes_gcm_work_bytes(len > aadlen ? len : aadlen);
        a += (16 - ((size_t)a & 15)) & 15;
        CU(cudaMemcpyAsync(a, aad, aadlen, cudaMemcpyDefault, st));
        daad = a;
    }
    if (direct) {
        din = in; dout = out;
    } else {
        if ((rc = grow(&c->big, &c->big_bytes, len + 32, "cudaMalloc(GCM-SIV staging)")) != 0) goto done;
        CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
        CU(cudaMemcpyAsync(c->big, in, len + (decrypt ? 16 : 0), cudaMemcpyDefault, st));
        din = c->big; dout = c->big;
    }
    dstate = NULL;
    if (aadlen >= AAD_BULK_MIN) {                             /* POLYVAL state of a large AAD, in bulk */
        dstate = (u8 *)c->work + AAD_STATE_OFF;
        LAUNCH(uaes_launch_gcmsiv_tag(&enc, derived, nonce, NULL, 0, NULL, daad, aadlen, 1, dstate, work, st));
    }

    if (!decrypt) {
        u8 *tagpos = (u8 *)dout + len;
        LAUNCH(uaes_launch_gcmsiv_tag(&enc, derived, nonce, daad, aadlen, dstate, din, len, 0, dtag, work, st));
        LAUNCH(uaes_launch_ctr32(&enc, dtag, din, dout, len, st));
        CU(cudaMemcpyAsync(tagpos, dtag, 16, cudaMemcpyDefault, st));
        if (!direct) CU(cudaMemcpyAsync(out, c->big, len + 16, cudaMemcpyDefault, st));
        if (!direct || !tls_async) CU(cudaStreamSynchronize(st));
    } else {
        u8 t1[16], t2[16];
        const u8 *rtag = (const u8 *)din + len;
        /* the received tag seeds the counter; keep a copy, in-place decryption may overwrite nothing
         * beyond len but the staging buffer is shared */
        CU(cudaMemcpyAsync(dtag + 16, rtag, 16, cudaMemcpyDefault, st));
        LAUNCH(uaes_launch_ctr32(&enc, dtag + 16, din, dout, len, st));
        LAUNCH(uaes_launch_gcmsiv_tag(&enc, derived, nonce, daad, aadlen, dstate, dout, len, 0, dtag, work, st));
        if (!direct && len) CU(cudaMemcpyAsync(out, c->big, len, cudaMemcpyDefault, st));
        CU(cudaMemcpyAsync(t1, dtag, 16, cudaMemcpyDefault, st));
        CU(cudaMemcpyAsync(t2, dtag + 16, 16, cudaMemcpyDefault, st));
        CU(cudaStreamSynchronize(st));
        if (memcmp(t1, t2, 16)) rc = UAES_AUTH_ERROR;         /* micro_aes.c:1510-1514 */
    }
done:
    pthread_mutex_unlock(&g_lock);
    return rc;
}

int uaes_gcmsiv_encrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                        const void *in, size_t len, void *out)
{
    return gcmsiv_common(keybits, key, nonce, aad, aadlen, in, len, out, 0);
}

int uaes_gcmsiv_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                        const void *in, size_t len, void *out)
{
    return gcmsiv_common(keybits, key, nonce, aad, aadlen, in, len, out, 1);
}

/* ------------------------------------------------------------------ CBC / CFB decrypt (8f, row 2) */

/* Both directions read two input blocks per output block, so in and out must be different
 * buffers on the device: the staged path uses the two halves of the full-size staging area. */
static int chain_common(int keybits, const u8 *key, const u8 *iv, const void *in, size_t len, void *out, int cbc)
{
    devctx *c;
    uaes_keysched enc, dec;
    int rc;
    size_t n = len / 16, r = len % 16;
    cudaStream_t st;

    if (expand_key(keybits, key, &enc)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    if (cbc) {                                                /* CS3 rules of micro_aes.c:751-760 */
        if (n > 1 && !r) { --n; r = 16; }
        if (n == 0) return UAES_DATALENGTH_ERROR;
        n -= r > 0;
        invert_schedule(&enc, &dec);
    } else if (len == 0) {
        return 0;
    }
    pthread_mutex_lock(&g_lock);
    if ((rc = get_ctx(&c)) != 0) goto done;
    if (is_direct(in) && is_direct(out) && in != out) {
        LAUNCH(uaes_launch_chain_dec(cbc ? &dec : &enc, &enc, cbc, iv, in, out, n, (unsigned)r, tls_stream));
        rc = finish_direct();
    } else {
        const size_t half = (len + 255) & ~(size_t)255;
        u8 *din, *dout;
        if ((rc = grow(&c->big, &c->big_bytes, 2 * half + 32, "cudaMalloc(CBC/CFB staging)")) != 0) goto done;
        din = (u8 *)c->big; dout = din + half;
        st = c->st[0];
        CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
        CU(cudaMemcpyAsync(din, in, len, cudaMemcpyDefault, st));
        LAUNCH(uaes_launch_chain_dec(cbc ? &dec : &enc, &enc, cbc, iv, din, dout, n, (unsigned)r, st));
        CU(cudaMemcpyAsync(out, dout, len, cudaMemcpyDefault, st));
        CU(cudaStreamSynchronize(st));
    }
done:
    pthread_mutex_unlock(&g_lock);
    return rc;
}

int uaes_cbc_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *iv, const void *in, size_t len, void *out)
{
    return chain_common(keybits, key, iv, in, len, out, 1);
}

int uaes_cfb_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *iv, const void *in, size_t len, void *out)
{
    return chain_common(keybits, key, iv, in, len, out, 0);
}

/* ------------------------------------------------------------------ OCB */

/* micro_aes.c:1779-1813.  One pass; decrypt writes the plaintext and then reports the tag
 * comparison, exactly like the reference (micro_aes.c:1806-1812). */
static int ocb_common(int keybits, const u8 *key, const u8 *nonce, const void *aad, size_t aadlen,
                      const void *in, size_t len, void *out, int decrypt)
{
    devctx *c;
    uaes_keysched enc, dec;
    int rc, direct;
    const void *din, *daad;
    void *dout;
    u8 *work, *dtag;
    cudaStream_t st;

    if (expand_key(keybits, key, &enc)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    if (decrypt) invert_schedule(&enc, &dec);
    pthread_mutex_lock(&g_lock);
    if ((rc = get_ctx(&c)) != 0) goto done;
    if (decrypt) direct = len == 0 || (is_direct(in) && is_direct(out));
    else         direct = is_direct(out) && (len == 0 || is_direct(in));
    st = direct ? (cudaStream_t)tls_stream : c->st[0];
    if ((rc = grow(&c->work, &c->work_bytes, GCM_WORK_HEAD + uaes_ocb_work_bytes() + aadlen + 64,
                   "cudaMalloc(OCB work)")) != 0) goto done;
    dtag = (u8 *)c->work;
    work = (u8 *)c->work + GCM_WORK_HEAD;
    daad = NULL;
    if (aadlen) {
        u8 *a = work + ((uaes_ocb_work_bytes() + 15) & ~(size_t)15);
        CU(cudaMemcpyAsync(a, aad, aadlen, cudaMemcpyDefault, st));
        daad = a;
    }
    if (direct) {
        din = in; dout = out;
    } else {
        if ((rc = grow(&c->big, &c->big_bytes, len + 32, "cudaMalloc(OCB staging)")) != 0) goto done;
        CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
        CU(cudaMemcpyAsync(c->big, in, len + (decrypt ? 16 : 0), cudaMemcpyDefault, st));
        din = c->big; dout = c->big;
    }
    if (!decrypt) {
        LAUNCH(uaes_launch_ocb(&enc, &enc, 1, nonce, daad, aadlen, din, dout, len, (u8 *)dout + len, work, st));
        if (!direct) CU(cudaMemcpyAsync(out, c->big, len + 16, cudaMemcpyDefault, st));
        if (!direct || !tls_async) CU(cudaStreamSynchronize(st));
    } else {
        u8 t1[16], t2[16];
        CU(cudaMemcpyAsync(t2, (const u8 *)din + len, 16, cudaMemcpyDefault, st));   /* before it can be overwritten */
        LAUNCH(uaes_launch_ocb(&enc, &dec, 0, nonce, daad, aadlen, din, dout, len, dtag, work, st));
        if (!direct && len) CU(cudaMemcpyAsync(out, c->big, len, cudaMemcpyDefault, st));
        CU(cudaMemcpyAsync(t1, dtag, 16, cudaMemcpyDefault, st));
        CU(cudaStreamSynchronize(st));
        if (memcmp(t1, t2, 16)) rc = UAES_AUTH_ERROR;
    }
done:
    pthread_mutex_unlock(&g_lock);
    return rc;
}

int uaes_ocb_encrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return ocb_common(keybits, key, nonce, aad, aadlen, in, len, out, 0);
}

int uaes_ocb_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return ocb_common(keybits, key, nonce, aad, aadlen, in, len, out, 1);
}

/* ------------------------------------------------------------------ CCM, batched */

/* One launch for n messages.  Host-side work is bookkeeping only: where the descriptors and the
 * three byte ranges live, staging whatever is host memory through the grow-only device buffers. */
#define BATCH_CCM 0
#define BATCH_EAX 1
#define BATCH_SIV 2
#define BATCH_GCM 3

static int launch_batch(int mode, const uaes_keysched *ks, const uaes_keysched *ks2, int decrypt, void *msgs_dev,
                        u64 n, const void *aad, const void *in, void *out, void *stream)
{
    if (mode == BATCH_CCM) return uaes_launch_ccm_batch(ks, decrypt, msgs_dev, n, aad, in, out, stream);
    return uaes_launch_mac_batch(mode, ks, ks2, decrypt, msgs_dev, n, aad, in, out, stream);
}

static int mac_batch(int mode, int keybits, const u8 *key, uaes_msg *msgs, size_t n,
                     const void *aad, const void *in, void *out, int decrypt)
{
    devctx *c;
    uaes_keysched ks, ks2;
    int rc = 0, msgs_dev;
    size_t i, in_ext = 0, out_ext = 0, aad_ext = 0, off;
    const void *din, *daad;
    void *dout, *dmsgs;
    cudaStream_t st;
    struct cudaPointerAttributes at;

    if (expand_key(keybits, key, &ks)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    ks2 = ks;
    if (mode == BATCH_SIV) expand_key(keybits, key + keybits / 8, &ks2);     /* keys = K1 || K2, micro_aes.c:1378 */
    if (n == 0) return 0;
    pthread_mutex_lock(&g_lock);
    if ((rc = get_ctx(&c)) != 0) goto done;
    msgs_dev = cudaPointerGetAttributes(&at, msgs) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    if (msgs_dev) {
        if (!is_direct(out) || !is_direct(in) || (aad && !is_direct(aad))) {
            rc = fail(UAES_E_BAD_ARGUMENT, "device descriptors need device (16-byte aligned) aad/in/out", 0);
            goto done;
        }
        st = (cudaStream_t)tls_stream;
        LAUNCH(launch_batch(mode, &ks, &ks2, decrypt, msgs, n, aad, in, out, st));
        if (!tls_async) CU(cudaStreamSynchronize(st));
        goto done;                                   /* per-message results stay on the device */
    }
    for (i = 0; i < n; ++i) {
        const size_t tag_in = decrypt ? 16 : 0, tag_out = decrypt ? 0 : 16;
        if (msgs[i].in_off + msgs[i].len + tag_in > in_ext) in_ext = msgs[i].in_off + msgs[i].len + tag_in;
        if (msgs[i].out_off + msgs[i].len + tag_out > out_ext) out_ext = msgs[i].out_off + msgs[i].len + tag_out;
        if (msgs[i].aad_len && msgs[i].aad_off + msgs[i].aad_len > aad_ext) aad_ext = msgs[i].aad_off + msgs[i].aad_len;
    }
    st = c->st[0];
    CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
    /* work area: descriptors, then (if host) the associated data; big: input and output ranges */
    off = (n * sizeof(uaes_msg) + 255) & ~(size_t)255;
    if ((rc = grow(&c->work, &c->work_bytes, off + aad_ext + 64, "cudaMalloc(batch descriptors)")) != 0) goto done;
    dmsgs = c->work;
    CU(cudaMemcpyAsync(dmsgs, msgs, n * sizeof(uaes_msg), cudaMemcpyDefault, st));
    daad = aad;
    if (aad_ext && !is_direct(aad)) {
        CU(cudaMemcpyAsync((u8 *)c->work + off, aad, aad_ext, cudaMemcpyDefault, st));
        daad = (u8 *)c->work + off;
    }
    din = in; dout = out;
    if (!is_direct(in) || !is_direct(out)) {
        const size_t ioff = (in_ext + 255) & ~(size_t)255;
        if ((rc = grow(&c->big, &c->big_bytes, ioff + out_ext + 64, "cudaMalloc(batch staging)")) != 0) goto done;
        if (in_ext) CU(cudaMemcpyAsync(c->big, in, in_ext, cudaMemcpyDefault, st));
        din = c->big; dout = (u8 *)c->big + ioff;
        /* bytes of the output range that no message covers must survive the copy back */
        if (out_ext) CU(cudaMemcpyAsync(dout, out, out_ext, cudaMemcpyDefault, st));
    }
    LAUNCH(launch_batch(mode, &ks, &ks2, decrypt, dmsgs, n, daad, din, dout, st));
    if (dout != out && out_ext) CU(cudaMemcpyAsync(out, dout, out_ext, cudaMemcpyDefault, st));
    CU(cudaMemcpyAsync(msgs, dmsgs, n * sizeof(uaes_msg), cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
    for (i = 0; i < n; ++i) if (msgs[i].result) rc = UAES_AUTH_ERROR;
done:
    pthread_mutex_unlock(&g_lock);
    return rc;
}

#define BATCH_ENTRY(name, mode, dec) \
    int name(int keybits, const uaes_u8 *key, uaes_msg *msgs, size_t n, const void *aad, const void *in, void *out) \
    { return mac_batch(mode, keybits, key, msgs, n, aad, in, out, dec); }
BATCH_ENTRY(uaes_ccm_encrypt_batch, BATCH_CCM, 0)
BATCH_ENTRY(uaes_ccm_decrypt_batch, BATCH_CCM, 1)
BATCH_ENTRY(uaes_eax_encrypt_batch, BATCH_EAX, 0)
BATCH_ENTRY(uaes_eax_decrypt_batch, BATCH_EAX, 1)
BATCH_ENTRY(uaes_siv_encrypt_batch, BATCH_SIV, 0)
BATCH_ENTRY(uaes_siv_decrypt_batch, BATCH_SIV, 1)
BATCH_ENTRY(uaes_gcm_encrypt_batch, BATCH_GCM, 0)
BATCH_ENTRY(uaes_gcm_decrypt_batch, BATCH_GCM, 1)

/* one message with the reference's argument list = a batch of one */
static int mac_single(int mode, int keybits, const u8 *key, const u8 *nonce, size_t noncelen, const void *aad,
                      size_t aadlen, const void *in, size_t len, void *out, int decrypt)
{
    uaes_msg m;
    if (len > 0xFFFFFFFFu - 16 || aadlen > 0xFFFFFFFFu) return fail(UAES_E_BAD_ARGUMENT, "message longer than 4 GiB", 0);
    memset(&m, 0, sizeof m);
    m.len = (unsigned int)len; m.aad_len = (unsigned int)aadlen;
    if (noncelen) memcpy(m.nonce, nonce, noncelen);
    return mac_batch(mode, keybits, key, &m, 1, aad, in, out, decrypt);
}

int uaes_ccm_encrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return mac_single(BATCH_CCM, keybits, key, nonce, 11, aad, aadlen, in, len, out, 0);
}

int uaes_ccm_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return mac_single(BATCH_CCM, keybits, key, nonce, 11, aad, aadlen, in, len, out, 1);
}

int uaes_eax_encrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return mac_single(BATCH_EAX, keybits, key, nonce, 16, aad, aadlen, in, len, out, 0);
}

int uaes_eax_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return mac_single(BATCH_EAX, keybits, key, nonce, 16, aad, aadlen, in, len, out, 1);
}

/* The reference keeps the synthetic IV and the ciphertext in separate buffers (micro_aes.c:1372,
 * 1394); the batch layout is IV || ciphertext.  A host temporary bridges the two. */
int uaes_siv_encrypt(int keybits, const uaes_u8 *keys, const void *aad, size_t aadlen,
                     const void *in, size_t len, uaes_u8 *iv, void *out)
{
    int rc;
    u8 *tmp = (u8 *)malloc(len + 16);
    if (!tmp) return fail(UAES_E_NO_MEMORY, "malloc(SIV temporary)", 0);
    rc = mac_single(BATCH_SIV, keybits, keys, NULL, 0, aad, aadlen, in, len, tmp, 0);
    if (rc == 0) {
        memcpy(iv, tmp, 16);
        if (len && cudaMemcpy(out, tmp + 16, len, cudaMemcpyDefault) != cudaSuccess)
            rc = fail(UAES_E_CUDA, "cudaMemcpy(SIV ciphertext)", (int)cudaGetLastError());
    }
    free(tmp);
    return rc;
}

int uaes_siv_decrypt(int keybits, const uaes_u8 *keys, const uaes_u8 *iv, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    int rc;
    u8 *tmp = (u8 *)malloc(len + 16);
    if (!tmp) return fail(UAES_E_NO_MEMORY, "malloc(SIV temporary)", 0);
    memcpy(tmp, iv, 16);
    if (len && cudaMemcpy(tmp + 16, in, len, cudaMemcpyDefault) != cudaSuccess) {
        free(tmp);
        return fail(UAES_E_CUDA, "cudaMemcpy(SIV ciphertext)", (int)cudaGetLastError());
    }
    rc = mac_single(BATCH_SIV, keybits, keys, NULL, 0, aad, aadlen, tmp, len, out, 1);
    free(tmp);
    return rc;
}

/* ------------------------------------------------------------------ streaming */

/* init / update / final on top of the range primitives: CTR is position bookkeeping around
 * uaes_ctr_crypt_range; GCM runs one fused CTR+GHASH shard pass per update and keeps the 16-byte
 * contributions, folding them on the device whenever 30 have piled up, so any number of updates
 * costs O(1) host memory.  The reference has no streaming API. */
#define STREAM_CTR 1
#define STREAM_GCM 2
#define STREAM_MAX_PENDING 30

struct uaes_stream {
    int kind, keybits, decrypt, ragged, npending;
    u8 key[32], nonce[12];
    u8 *aad; size_t aadlen;
    u64 pos_blocks, total_len;
    u8 partial[STREAM_MAX_PENDING + 1][16];
    u64 end_block[STREAM_MAX_PENDING + 1];
};

static uaes_stream *stream_new(int kind, int keybits, const u8 *key, const u8 *nonce)
{
    uaes_stream *s;
    uaes_keysched ks;
    if (expand_key(keybits, key, &ks)) { fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0); return NULL; }
    s = (uaes_stream *)calloc(1, sizeof *s);
    if (!s) { fail(UAES_E_NO_MEMORY, "calloc(stream)", 0); return NULL; }
    s->kind = kind; s->keybits = keybits;
    memcpy(s->key, key, (size_t)keybits / 8);
    memcpy(s->nonce, nonce, 12);
    return s;
}

uaes_stream *uaes_stream_ctr(int keybits, const uaes_u8 *key, const uaes_u8 *iv)
{
    return stream_new(STREAM_CTR, keybits, key, iv);
}

uaes_stream *uaes_stream_gcm(int keybits, const uaes_u8 *key, const uaes_u8 *nonce,
                             const void *aad, size_t aadlen, int decrypt)
{
    uaes_stream *s = stream_new(STREAM_GCM, keybits, key, nonce);
    if (!s) return NULL;
    s->decrypt = decrypt;
    if (aadlen) {
        s->aad = (u8 *)malloc(aadlen);
        if (!s->aad || cudaMemcpy(s->aad, aad, aadlen, cudaMemcpyDefault) != cudaSuccess) {
            cudaGetLastError();
            fail(UAES_E_NO_MEMORY, "stream AAD copy", 0);
            free(s->aad); free(s);
            return NULL;
        }
        s->aadlen = aadlen;
    }
    return s;
}

void uaes_stream_free(uaes_stream *s)
{
    if (!s) return;
    free(s->aad);
    memset(s, 0, sizeof *s);
    free(s);
}

/* sum_r partial[r] * H^(end_last - end_r) -> one pending entry */
static int stream_fold(uaes_stream *s)
{
    devctx *c;
    uaes_keysched ks;
    int rc = 0, i;
    u64 after[STREAM_MAX_PENDING + 1];
    u8 *w;
    cudaStream_t st = (cudaStream_t)tls_stream;
    const u64 end = s->end_block[s->npending - 1];

    expand_key(s->keybits, s->key, &ks);
    for (i = 0; i < s->npending; ++i) after[i] = end - s->end_block[i];
    pthread_mutex_lock(&g_lock);
    if ((rc = get_ctx(&c)) != 0) goto done;
    if ((rc = grow(&c->work, &c->work_bytes, GCM_WORK_HEAD + 32 * 24 + 64, "cudaMalloc(GCM work)")) != 0) goto done;
    w = (u8 *)c->work;
    CU(cudaMemcpyAsync(w + 64, s->partial, (size_t)s->npending * 16, cudaMemcpyDefault, st));
    CU(cudaMemcpyAsync(w + 64 + 32 * 16, after, (size_t)s->npending * 8, cudaMemcpyDefault, st));
    LAUNCH(uaes_launch_gcm_fold(&ks, w + 64, w + 64 + 32 * 16, (unsigned)s->npending, w, st));
    CU(cudaMemcpyAsync(s->partial[0], w, 16, cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
    s->end_block[0] = end;
    s->npending = 1;
done:
    pthread_mutex_unlock(&g_lock);
    return rc;
}

int uaes_stream_update(uaes_stream *s, const void *in, size_t len, void *out)
{
    int rc;
    const u64 nb = ((u64)len + 15) / 16;
    if (!s) return fail(UAES_E_BAD_ARGUMENT, "null stream", 0);
    if (len == 0) return 0;
    if (s->ragged) return fail(UAES_E_BAD_ARGUMENT, "only the last update may have a length that is not a multiple of 16", 0);
    if (s->kind == STREAM_CTR) {
        rc = uaes_ctr_crypt_range(s->keybits, s->key, s->nonce, s->pos_blocks, in, len, out);
    } else {
        if (s->npending == STREAM_MAX_PENDING && (rc = stream_fold(s)) != 0) return rc;
        rc = uaes_gcm_shard(s->keybits, s->key, s->nonce, s->pos_blocks, in, len, out, s->decrypt,
                            s->partial[s->npending]);
        if (rc == 0) { s->end_block[s->npending] = s->pos_blocks + nb; ++s->npending; }
    }
    if (rc) return rc;
    s->pos_blocks += nb;
    s->total_len += len;
    if (len % 16) s->ragged = 1;
    return 0;
}

/* GCM encrypt: writes the 16-byte tag.  GCM decrypt: tag = the received tag; UAES_AUTH_ERROR means
 * everything the updates produced must be discarded.  CTR: nothing to do (tag may be NULL). */
int uaes_stream_final(uaes_stream *s, uaes_u8 *tag)
{
    int rc, i;
    u64 after[STREAM_MAX_PENDING + 1];
    u8 t[16];
    if (!s) return fail(UAES_E_BAD_ARGUMENT, "null stream", 0);
    if (s->kind == STREAM_CTR) return 0;
    for (i = 0; i < s->npending; ++i) after[i] = s->pos_blocks - s->end_block[i];
    rc = uaes_gcm_combine(s->keybits, s->key, s->nonce, s->aad, s->aadlen, &s->partial[0][0], after,
                          s->npending, s->total_len, t);
    if (rc) return rc;
    if (!s->decrypt) { memcpy(tag, t, 16); return 0; }
    return memcmp(t, tag, 16) ? UAES_AUTH_ERROR : 0;
}

/* ------------------------------------------------------------------ synthetic data */

int uaes_fill_splitmix64(uaes_u64 seed, uaes_u64 first_word, void *dst, size_t nwords)
{
    devctx *c;
    int rc;
    pthread_mutex_lock(&g_lock);
    if ((rc = get_ctx(&c)) != 0) goto done;
    LAUNCH(uaes_launch_fill(seed, first_word, dst, nwords, tls_stream));
    rc = finish_direct();
done:
    pthread_mutex_unlock(&g_lock);
    return rc;
}

int uaes_xor_fold64(const void *src, size_t nwords, uaes_u64 *result)
{
    devctx *c;
    int rc;
    pthread_mutex_lock(&g_lock);
    if ((rc = get_ctx(&c)) != 0) goto done;
    if ((rc = grow(&c->work, &c->work_bytes, GCM_WORK_HEAD, "cudaMalloc(work)")) != 0) goto done;
    LAUNCH(uaes_launch_xor_fold(src, nwords, c->work, tls_stream));
    CU(cudaMemcpyAsync(result, c->work, 8, cudaMemcpyDefault, (cudaStream_t)tls_stream));
    CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
done:
    pthread_mutex_unlock(&g_lock);
    return rc;
}
