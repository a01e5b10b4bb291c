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
 *   pinned host memory                 -> staged through three device chunks on three streams so
 *                                         that H2D, kernel and D2H of consecutive chunks overlap;
 *   pageable host memory               -> the same pipeline behind pinned bounce chunks that a
 *                                         few helper threads fill and drain (memcpy only);
 *   misaligned device memory           -> the same pipeline with device-to-device copies.
 * Every mode of the hot path is cut into chunks: CTR / ECB / XTS sectors by block or sector
 * range, one XTS data unit by tweak jump-ahead, GCM as a sequence of shards whose 16-byte GHASH
 * contributions are folded at the end.  With uaes_set_devices(n) a host-buffer call is spread
 * over n GPUs, one part and one host thread per device, each under its own device's lock.
 */
#define _POSIX_C_SOURCE 200809L      /* clock_gettime, pthread_cond_timedwait under -std=c99 */
#include <cuda_runtime_api.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

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

int uaes_ctr_queue_stats(uaes_u64 *tt_units, uaes_u64 *bs_units, uaes_u64 *unit_blocks)
{
    const int e = uaes_launch_ctr_queue_stats(tt_units, bs_units, unit_blocks);
    return e ? fail(UAES_E_CUDA, "no work-queue CTR launch on this thread yet", 0) : 0;
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

/* a caller-supplied 16-byte counter block (PRESET_COUNTER, micro_aes.c:964-966; GCM's J0) advanced
 * by `add` blocks: bytes 9..15 are the 56-bit big-endian counter (micro_aes.c:421-427) */
static void block_to_ctrblock(const u8 *blk, u64 add, uaes_ctrblock *cb)
{
    u64 v = 0;
    int i;
    for (i = 9; i < 16; ++i) v = v << 8 | blk[i];
    cb->w0 = (u32)blk[0] | (u32)blk[1] << 8 | (u32)blk[2] << 16 | (u32)blk[3] << 24;
    cb->w1 = (u32)blk[4] | (u32)blk[5] << 8 | (u32)blk[6] << 16 | (u32)blk[7] << 24;
    cb->b8 = blk[8];
    cb->v0 = (v + add) & (((u64)1 << 56) - 1);
}

/* ------------------------------------------------------------------ per-device resources */

/* One devctx per CUDA device, each with its OWN lock: threads driving different GPUs never wait for
 * each other.  The lock covers what is shared per device -- the staging slots (a staged call owns
 * them for its whole H2D / kernel / D2H pipeline), the full-size staging buffer and the
 * bookkeeping of the scratch pool.  Calls on device-resident buffers take it only while they pick a
 * scratch block; plain CTR / ECB / XTS launches on device memory take no lock at all. */
#define MAX_SCRATCH 16

typedef struct {
    void *p; size_t bytes;
    cudaEvent_t ev;                     /* recorded after the last kernel that used the block */
    cudaStream_t last;                  /* ... on this stream */
    int busy, used, have_ev;
} scratch;

typedef struct {
    int ready, dev;
    pthread_mutex_t lock;               /* staging: slots, streams st[], big */
    pthread_mutex_t plock;              /* scratch pool bookkeeping (taken briefly, may nest inside lock) */
    void *slot[MAX_SLOT];               /* device chunks of the staging pipeline */
    void *hslot[MAX_SLOT];              /* pinned host chunks: bounce buffers for PAGEABLE caller memory */
    cudaStream_t st[MAX_SLOT];
    cudaEvent_t ev[MAX_SLOT];
    cudaEvent_t rev[16]; int rev_ok[16];  /* one event per ring slot of the pageable (bounce) pipeline */
    void *big;  size_t big_bytes;       /* full-size staging (GCM decrypt, CBC/CFB, batches); trimmed after use */
    scratch pool[MAX_SCRATCH];          /* GCM / OCB / batch work areas: one block per call in flight */
} devctx;

static devctx g_dev[MAX_DEV];
static pthread_once_t g_dev_once = PTHREAD_ONCE_INIT;
static pthread_mutex_t g_cfg_lock = PTHREAD_MUTEX_INITIALIZER;     /* first-use initialisation only */
static int g_cfg_ready;

/* process-wide settings (uaes_set_devices, uaes_set_burn, ...) */
static int    g_fan_devices = 1;                      /* devices a host-buffer call may be spread over */
static size_t g_fan_min = (size_t)256 << 20;          /* ... when every device gets at least this much */
static int    g_fan_oversubscribe;                    /* tests: more parts than devices (UAES_FANOUT_OVERSUBSCRIBE) */
static int    g_burn;                                 /* wipe staging memory and scratch after each call */
static size_t g_big_keep = (size_t)256 << 20;         /* full-size staging above this is freed after the call */
static int    g_copy_threads = 8;
static int    g_in_ahead = 2;                         /* chunks of pageable input copied ahead of the GPU stage */                     /* helpers that move pageable memory to / from the pinned chunks */

static void dev_locks_init(void)
{
    int i;
    for (i = 0; i < MAX_DEV; ++i) {
        pthread_mutex_init(&g_dev[i].lock, NULL);
        pthread_mutex_init(&g_dev[i].plock, NULL);
        g_dev[i].dev = i;
    }
}

static void cfg_init(void)
{
    const char *e;
    pthread_mutex_lock(&g_cfg_lock);
    if (!g_cfg_ready) {
        if ((e = getenv("UAES_STAGE_SLOTS")) != NULL && atoi(e) >= 1 && atoi(e) <= MAX_SLOT) g_nslot = atoi(e);
        if ((e = getenv("UAES_STAGE_CHUNK_MIB")) != NULL && atoi(e) >= 1 && (size_t)atoi(e) <= (MAX_CHUNK >> 20))
            g_chunk = (size_t)atoi(e) << 20;
        /* more parts than devices: the parts share devices and run one after the other there -- no
         * gain, but the whole splitting / threading path runs on a one-GPU box (tests) */
        if ((e = getenv("UAES_FANOUT_OVERSUBSCRIBE")) != NULL) g_fan_oversubscribe = atoi(e) != 0;
        if ((e = getenv("UAES_DEVICES")) != NULL) {
            int n = atoi(e), have = uaes_device_count();
            g_fan_devices = (n <= 0 || (n > have && !g_fan_oversubscribe)) ? have : n;
            if (g_fan_devices < 1) g_fan_devices = 1;
        }
        if ((e = getenv("UAES_FANOUT_MIN_MIB")) != NULL && atoi(e) >= 1) g_fan_min = (size_t)atoi(e) << 20;
        if ((e = getenv("UAES_BURN")) != NULL) g_burn = atoi(e) != 0;
        if ((e = getenv("UAES_IN_AHEAD")) != NULL && atoi(e) >= 1) g_in_ahead = atoi(e);
        if ((e = getenv("UAES_COPY_THREADS")) != NULL && atoi(e) >= 1 && atoi(e) <= 32) g_copy_threads = atoi(e);
        g_cfg_ready = 1;
    }
    pthread_mutex_unlock(&g_cfg_lock);
}

/* context of the calling thread's current device; creates its streams on first use */
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
    pthread_once(&g_dev_once, dev_locks_init);
    if (!g_cfg_ready) cfg_init();
    c = &g_dev[dev];
    if (!c->ready) {
        pthread_mutex_lock(&c->lock);
        if (!c->ready) {
            for (i = 0; i < MAX_SLOT; ++i)
                if (cudaStreamCreateWithFlags(&c->st[i], cudaStreamNonBlocking) != cudaSuccess ||
                    cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming) != cudaSuccess) {
                    pthread_mutex_unlock(&c->lock);
                    return fail(UAES_E_CUDA, "cudaStreamCreate", (int)cudaGetLastError());
                }
            c->ready = 1;
        }
        pthread_mutex_unlock(&c->lock);
    }
    *out = c;
    return 0;
}

/* lock held */
static int need_slots(devctx *c, int pinned_too)
{
    int i;
    for (i = 0; i < NSLOT; ++i) {
        if (!c->slot[i] && cudaMalloc(&c->slot[i], CHUNK_BYTES + 32) != cudaSuccess)
            return fail(UAES_E_NO_MEMORY, "cudaMalloc(staging chunk)", (int)cudaGetLastError());
        if (pinned_too && !c->hslot[i] && cudaHostAlloc(&c->hslot[i], CHUNK_BYTES + 32, cudaHostAllocDefault) != cudaSuccess)
            return fail(UAES_E_NO_MEMORY, "cudaHostAlloc(bounce chunk)", (int)cudaGetLastError());
    }
    return 0;
}

/* lock held */
static int grow(void **p, size_t *have, size_t want, const char *what)
{
    if (*have >= want) return 0;
    if (*p) { cudaFree(*p); *p = NULL; *have = 0; }
    want += want / 8 + 4096;
    if (cudaMalloc(p, want) != cudaSuccess) return fail(UAES_E_NO_MEMORY, what, (int)cudaGetLastError());
    *have = want;
    return 0;
}

/* lock held, all work on `big` complete: a multi-GiB staging buffer does not outlive its call
 * (one 16 GiB host GCM call used to pin 16 GiB of HBM for the life of the process) */
static void big_done(devctx *c)
{
    if (c->big && g_burn) cudaMemset(c->big, 0, c->big_bytes);
    if (c->big && c->big_bytes > g_big_keep) { cudaFree(c->big); c->big = NULL; c->big_bytes = 0; }
}

/* A work area for ONE call: never shared by two calls in flight.  A block is handed out again only
 * to the stream that used it last (stream order protects it) or once its event has completed, so
 * asynchronous calls on different user streams cannot trample each other's H, E(J0), partials or
 * tag scratch. */
static int scratch_get_locked(devctx *c, size_t want, cudaStream_t st, scratch **out)
{
    int i, pick = -1, idle = -1, empty = -1, waitable = -1;
    for (i = 0; i < MAX_SCRATCH; ++i) {
        scratch *s = &c->pool[i];
        if (s->busy) continue;
        if (!s->p) { if (empty < 0) empty = i; continue; }
        if (!s->used || s->last == st || cudaEventQuery(s->ev) == cudaSuccess) {
            if (s->bytes >= want) { pick = i; break; }
            if (idle < 0) idle = i;
        } else if (waitable < 0) waitable = i;
    }
    cudaGetLastError();                                   /* cudaErrorNotReady from the queries */
    if (pick < 0) {
        scratch *s;
        if (empty >= 0) pick = empty;
        else if (idle >= 0) pick = idle;
        else if (waitable >= 0) { pick = waitable; cudaEventSynchronize(c->pool[pick].ev); }
        else return fail(UAES_E_NO_MEMORY, "more than 16 calls in flight on one device", 0);
        s = &c->pool[pick];
        if (s->p && s->bytes < want) {
            if (s->used && s->last == st) cudaStreamSynchronize(st);      /* still queued work may read it */
            cudaFree(s->p); s->p = NULL; s->bytes = 0;
        }
        if (!s->p) {
            const size_t b = want + want / 8 + 4096;
            if (cudaMalloc(&s->p, b) != cudaSuccess) return fail(UAES_E_NO_MEMORY, "cudaMalloc(work area)", (int)cudaGetLastError());
            s->bytes = b; s->used = 0;
        }
        if (!s->have_ev) {
            if (cudaEventCreateWithFlags(&s->ev, cudaEventDisableTiming) != cudaSuccess)
                return fail(UAES_E_CUDA, "cudaEventCreate", (int)cudaGetLastError());
            s->have_ev = 1;
        }
    }
    c->pool[pick].busy = 1;
    *out = &c->pool[pick];
    return 0;
}

static int scratch_get(devctx *c, size_t want, cudaStream_t st, scratch **out)
{
    int rc;
    pthread_mutex_lock(&c->plock);
    rc = scratch_get_locked(c, want, st, out);
    pthread_mutex_unlock(&c->plock);
    return rc;
}

/* everything that uses the block has been enqueued on st */
static void scratch_put(devctx *c, scratch *s, cudaStream_t st)
{
    if (!s) return;
    pthread_mutex_lock(&c->plock);
    if (g_burn) cudaMemsetAsync(s->p, 0, s->bytes, st);   /* H, E(J0), derived keys, tag scratch */
    cudaEventRecord(s->ev, st);
    s->last = st; s->used = 1; s->busy = 0;
    pthread_mutex_unlock(&c->plock);
}

#define PTR_DEVICE   0   /* device or managed memory, 16-byte aligned: kernels touch it directly */
#define PTR_PINNED   1   /* page-locked host memory (cudaHostAlloc / cudaHostRegister): DMA at full speed */
#define PTR_PAGEABLE 2   /* plain malloc'd memory: moved through pinned bounce chunks by helper threads */
#define PTR_OTHER    3   /* misaligned device memory: staged with device-to-device copies */

static int ptr_class(const void *p)
{
    struct cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return PTR_PAGEABLE; }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)
        return ((size_t)p & 15) == 0 ? PTR_DEVICE : PTR_OTHER;
    if (at.type == cudaMemoryTypeHost) return PTR_PINNED;
    return PTR_PAGEABLE;
}

/* device (or managed) memory that the kernels may touch directly with 128-bit accesses */
static int is_direct(const void *p) { return ptr_class(p) == PTR_DEVICE; }

static int finish_direct(void)
{
    if (!tls_async) {
        cudaError_t e = cudaStreamSynchronize((cudaStream_t)tls_stream);
        if (e != cudaSuccess) return fail(UAES_E_CUDA, "cudaStreamSynchronize", (int)e);
    }
    return 0;
}

/* ------------------------------------------------------------------ helpers for pageable memory */

/* cudaMemcpyAsync on pageable memory is synchronous and single-threaded inside the driver, which
 * serialises the whole pipeline (H2D of chunk k+1 cannot overlap D2H of chunk k).  For pageable
 * caller buffers the library therefore owns pinned bounce chunks and moves the bytes itself with a
 * few helper threads: memcpy into the pinned chunk, DMA, kernel, DMA, memcpy out -- the memcpys of
 * neighbouring chunks overlap each other and the DMA.  This is data movement only; no cipher work
 * happens on the host. */
typedef struct { u8 *dst; const u8 *src; size_t n; int *left; } copy_piece;

#define COPY_QUEUE 1024
#define COPY_PIECE ((size_t)1 << 20)
typedef struct copy_pool {
    pthread_mutex_t m;
    pthread_cond_t work, done;
    copy_piece q[COPY_QUEUE];
    unsigned head, tail;
    int nthreads;
    pthread_t th[32];
} copy_pool;

static copy_pool g_cp = { PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER,
                          {{0, 0, 0, 0}}, 0, 0, 0, {0} };

static void *copy_worker(void *arg)
{
    copy_pool *p = (copy_pool *)arg;
    pthread_mutex_lock(&p->m);
    for (;;) {
        copy_piece w;
        while (p->head == p->tail) pthread_cond_wait(&p->work, &p->m);
        w = p->q[p->head % COPY_QUEUE]; ++p->head;
        pthread_mutex_unlock(&p->m);
        memcpy(w.dst, w.src, w.n);
        pthread_mutex_lock(&p->m);
        --*w.left;
        pthread_cond_broadcast(&p->done);
    }
    return NULL;
}

/* queue a copy of n bytes for the helper threads (shared by all devices' pipelines) in pieces of
 * 1 MiB; *left counts the pieces still to arrive (read it with copy_left / copy_wait) */
static void copy_async(void *dst, const void *src, size_t n, int *left)
{
    copy_pool *p = &g_cp;
    size_t off;
    pthread_mutex_lock(&p->m);
    while (p->nthreads < g_copy_threads && p->nthreads < 32) {
        pthread_attr_t at;
        pthread_attr_init(&at);
        pthread_attr_setdetachstate(&at, PTHREAD_CREATE_DETACHED);
        if (pthread_create(&p->th[p->nthreads], &at, copy_worker, p) != 0) { pthread_attr_destroy(&at); break; }
        pthread_attr_destroy(&at);
        ++p->nthreads;
    }
    if (p->nthreads == 0) {                               /* no helper could be started: copy here */
        pthread_mutex_unlock(&p->m);
        memcpy(dst, src, n);
        return;
    }
    for (off = 0; off < n; off += COPY_PIECE) {
        copy_piece *q;
        while (p->tail - p->head >= COPY_QUEUE) pthread_cond_wait(&p->done, &p->m);
        q = &p->q[p->tail % COPY_QUEUE];
        q->dst = (u8 *)dst + off; q->src = (const u8 *)src + off; q->n = n - off < COPY_PIECE ? n - off : COPY_PIECE; q->left = left;
        ++p->tail; ++*left;
        pthread_cond_signal(&p->work);
    }
    pthread_mutex_unlock(&p->m);
}

static int copy_left(int *left)
{
    int v;
    pthread_mutex_lock(&g_cp.m);
    v = *left;
    pthread_mutex_unlock(&g_cp.m);
    return v;
}

/* The pipeline's own thread has nothing to enqueue: it copies a queued piece itself (waking a sleeping
 * helper costs ~200 us on the measured hosts -- more than copying 1 MiB), or, when the queue is empty,
 * waits until some piece of anybody has arrived, for 200 us at most */
static void copy_wait_any(void)
{
    copy_pool *p = &g_cp;
    pthread_mutex_lock(&p->m);
    if (p->head != p->tail) {
        const copy_piece w = p->q[p->head % COPY_QUEUE];
        ++p->head;
        pthread_mutex_unlock(&p->m);
        memcpy(w.dst, w.src, w.n);
        pthread_mutex_lock(&p->m);
        --*w.left;
        pthread_cond_broadcast(&p->done);
    } else {
        struct timespec ts;
        clock_gettime(CLOCK_REALTIME, &ts);
        ts.tv_nsec += 200000;
        if (ts.tv_nsec >= 1000000000L) { ts.tv_nsec -= 1000000000L; ++ts.tv_sec; }
        pthread_cond_timedwait(&p->done, &p->m, &ts);
    }
    pthread_mutex_unlock(&p->m);
}

/* UAES_TRACE=1: timestamps of the pipeline stages on stderr (debugging / tuning only) */
static int g_trace = -1;
static double now_us(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec * 1e6 + (double)ts.tv_nsec * 1e-3;
}
#define TRACE(...) do { if (g_trace > 0) fprintf(stderr, __VA_ARGS__); } while (0)

/* ------------------------------------------------------------------ chunked staging pipeline */

/* one chunk, resident in a device slot: offset = byte offset inside the WHOLE call (not the part),
 * index = number of the chunk inside this device's part, slot = which slot / work area it uses */
typedef int (*chunk_fn)(void *user, u64 offset, size_t index, int slot, void *dev, size_t bytes, size_t out_bytes,
                        void *stream);

typedef struct {
    const u8 *in; u8 *out;       /* this part's first byte (host or device memory, any mix) */
    size_t len;                  /* bytes of this part */
    u64 base;                    /* offset of the part inside the whole call */
    size_t unit;                 /* a chunk is a multiple of this */
    size_t out_extra;            /* bytes the LAST chunk writes beyond its input size (ECB padding) */
    size_t absorb_tail;          /* a final piece of fewer than this many bytes joins the chunk before it
                                    (XTS stealing needs the last full block and the ragged tail together) */
    u8 *resident;                /* non-NULL: chunk k stays at resident + offset (device memory holding the whole
                                    part) instead of a slot, and nothing is copied back (out is NULL) */
    chunk_fn fn; void *user;     /* fn may be NULL: a pure copy */
} pipe_part;

/* Geometry of one part's pipeline.  Pinned or device memory: NSLOT chunks of CHUNK_BYTES, DMA straight
 * from / to the caller's buffer.  Pageable memory on either side: the same device and pinned chunks
 * are cut into a ring of smaller pieces (BOUNCE_BYTES) so that the helper threads' memcpy in, the DMA
 * both ways, the kernel and the memcpy out of different pieces all overlap (measured on the B200 box,
 * profiles/r2_hostmem_probe.txt: cudaMemcpy of pageable memory 10.7 / 20.1 GiB/s, cudaHostRegister
 * 10.9 GiB/s, 8 memcpy threads 74 GiB/s). */
#define BOUNCE_BYTES ((size_t)8 << 20)
#define MAX_RING 16                   /* NSLOT <= 8: at least 2 pieces per chunk */

#define BOUNCE_MIN ((size_t)64 << 10)  /* smaller pageable calls go through plain cudaMemcpyAsync (one stream, no
                                         helpers): the fixed costs dominate there */

static int pipe_bounces(const pipe_part *p)
{
    if (p->len <= BOUNCE_MIN) return 0;
    return ptr_class(p->in) == PTR_PAGEABLE || (p->out && ptr_class(p->out) == PTR_PAGEABLE);
}

static size_t pipe_chunk_bytes(const pipe_part *p)
{
    size_t chunk = CHUNK_BYTES - CHUNK_BYTES % p->unit;
    if (pipe_bounces(p) && chunk > BOUNCE_BYTES && BOUNCE_BYTES >= p->unit) chunk = BOUNCE_BYTES - BOUNCE_BYTES % p->unit;
    return chunk;
}

/* pieces per staging chunk: 1 for pinned / device memory */
static size_t pipe_per(const pipe_part *p)
{
    const size_t chunk = pipe_chunk_bytes(p);
    size_t per;
    if (!pipe_bounces(p) || chunk == 0) return 1;
    per = (CHUNK_BYTES + 32) / (chunk + 32);               /* 32 bytes of slack behind each piece */
    if (per > (size_t)MAX_RING / (size_t)NSLOT) per = (size_t)MAX_RING / (size_t)NSLOT;
    return per < 1 ? 1 : per;
}

/* ring slots (= chunks in flight, = work areas a chunk function may index with `slot`) */
static int pipe_slots(const pipe_part *p) { return (int)(pipe_per(p) * (size_t)NSLOT); }

static size_t pipe_nchunks(const pipe_part *p)
{
    const size_t chunk = pipe_chunk_bytes(p);
    size_t n;
    if (p->len == 0 || chunk == 0) return 0;
    n = (p->len + chunk - 1) / chunk;
    if (n > 1 && p->len - (n - 1) * chunk < p->absorb_tail) --n;
    return n;
}

/* Runs one part through the device's ring.  Caller holds c->lock and has made `c` current.
 * Four stages per chunk, each started as soon as its predecessor has finished and a ring slot is free:
 *   IN   helper threads copy pageable input into the slot's pinned piece     (skipped: pinned / device input)
 *   GPU  H2D, the mode's kernel(s), D2H, an event -- all on the slot's stream
 *   OUT  helper threads copy the pinned piece to pageable output             (skipped: pinned / device output)
 *   retire: the slot may be used again */
static int run_chunked(devctx *c, const pipe_part *p)
{
    int rc = 0, i;
    const size_t chunk = pipe_chunk_bytes(p), n = pipe_nchunks(p);
    const int R = pipe_slots(p);
    const int bounce = pipe_bounces(p);
    const int bounce_in = bounce && ptr_class(p->in) == PTR_PAGEABLE, bounce_out = bounce && p->out && ptr_class(p->out) == PTR_PAGEABLE;
    const size_t per = pipe_per(p);                          /* ring slots per staging chunk */
    size_t kin = 0, kgpu = 0, kout = 0, kdone = 0;
    int in_left[MAX_RING], out_left[MAX_RING];

    double t0;
    if (g_trace < 0) g_trace = getenv("UAES_TRACE") != NULL;
    t0 = now_us();
    if (chunk == 0) return fail(UAES_E_BAD_ARGUMENT, "unit larger than the staging chunk", 0);
    if ((rc = need_slots(c, bounce)) != 0) return rc;
    for (i = 0; i < MAX_RING; ++i) in_left[i] = out_left[i] = 0;
    TRACE("[uaes] run_chunked len %zu chunk %zu n %zu R %d bounce %d/%d  (+%.0f us)\n", p->len, chunk, n, R, bounce_in, bounce_out, now_us() - t0);
    if (bounce)
        for (i = 0; i < R; ++i)
            if (!c->rev_ok[i]) {
                CU(cudaEventCreateWithFlags(&c->rev[i], cudaEventDisableTiming));
                c->rev_ok[i] = 1;
            }

#define SLOT_OF(k)   ((int)((k) % (size_t)R))
#define DEV_OF(r)    ((u8 *)c->slot[(size_t)(r) / per] + ((size_t)(r) % per) * (chunk + 32))
#define HOST_OF(r)   ((u8 *)c->hslot[(size_t)(r) / per] + ((size_t)(r) % per) * (chunk + 32))
#define STREAM_OF(r) (c->st[(r) % NSLOT])
#define OFF_OF(k)    ((k) * chunk)
#define BYTES_OF(k)  ((k) + 1 == n ? p->len - OFF_OF(k) : chunk)

    if (!bounce) {
        /* pinned / device memory: everything is stream ordered, nothing to wait for on the host */
        size_t k;
        for (k = 0; k < n; ++k) {
            const size_t off = OFF_OF(k), bytes = BYTES_OF(k), obytes = bytes + (k + 1 == n ? p->out_extra : 0);
            const int r = SLOT_OF(k);
            u8 *d = p->resident ? p->resident + off : (u8 *)c->slot[r];
            CU(cudaMemcpyAsync(d, p->in + off, bytes, cudaMemcpyDefault, c->st[r]));
            if (p->fn && (rc = p->fn(p->user, p->base + off, k, r, d, bytes, obytes, c->st[r])) != 0) goto done;
            if (p->out) CU(cudaMemcpyAsync(p->out + off, d, obytes, cudaMemcpyDefault, c->st[r]));
            if (g_burn && !p->resident) CU(cudaMemsetAsync(c->slot[r], 0, CHUNK_BYTES, c->st[r]));
        }
        goto done;
    }

    while (kdone < n) {
        int progressed = 0;
        /* IN: a ring slot must be free; pageable input is copied only a little ahead of the GPU stage, or the
         * helpers would fill the whole ring first and the copies out (which free slots) would queue behind */
        if (kin < n && kin - kdone < (size_t)R && (!bounce_in || kin - kgpu < (size_t)g_in_ahead)) {
            const int r = SLOT_OF(kin);
            if (bounce_in) copy_async(HOST_OF(r), p->in + OFF_OF(kin), BYTES_OF(kin), &in_left[r]);
            TRACE("[uaes] %8.0f IN  %zu\n", now_us() - t0, kin);
            ++kin; progressed = 1;
        }
        if (kgpu < kin && (!bounce_in || copy_left(&in_left[SLOT_OF(kgpu)]) == 0)) {      /* GPU */
            const size_t k = kgpu, off = OFF_OF(k), bytes = BYTES_OF(k), obytes = bytes + (k + 1 == n ? p->out_extra : 0);
            const int r = SLOT_OF(k);
            u8 *d = p->resident ? p->resident + off : DEV_OF(r);
            cudaStream_t st = STREAM_OF(r);
            if (bounce_in) CU(cudaMemcpyAsync(d, HOST_OF(r), bytes, cudaMemcpyHostToDevice, st));
            else           CU(cudaMemcpyAsync(d, p->in + off, bytes, cudaMemcpyDefault, st));
            if (p->fn && (rc = p->fn(p->user, p->base + off, k, r, d, bytes, obytes, st)) != 0) goto done;
            if (p->out) {
                if (bounce_out) CU(cudaMemcpyAsync(HOST_OF(r), d, obytes, cudaMemcpyDeviceToHost, st));
                else            CU(cudaMemcpyAsync(p->out + off, d, obytes, cudaMemcpyDefault, st));
            }
            if (g_burn && !p->resident) CU(cudaMemsetAsync(d, 0, chunk, st));
            CU(cudaEventRecord(c->rev[r], st));
            TRACE("[uaes] %8.0f GPU %zu\n", now_us() - t0, kgpu);
            ++kgpu; progressed = 1;
        }
        if (kout < kgpu) {                                                  /* OUT */
            const int r = SLOT_OF(kout);
            const cudaError_t q = cudaEventQuery(c->rev[r]);
            if (q == cudaSuccess) {
                const size_t k = kout, obytes = BYTES_OF(k) + (k + 1 == n ? p->out_extra : 0);
                if (bounce_out && p->out) copy_async(p->out + OFF_OF(k), HOST_OF(r), obytes, &out_left[r]);
                TRACE("[uaes] %8.0f OUT %zu\n", now_us() - t0, kout);
                ++kout; progressed = 1;
            } else if (q != cudaErrorNotReady) {
                rc = fail(UAES_E_CUDA, "staging pipeline", (int)q);
                goto done;
            } else {
                cudaGetLastError();
            }
        }
        if (kdone < kout && (!bounce_out || copy_left(&out_left[SLOT_OF(kdone)]) == 0)) {    /* retire */
            TRACE("[uaes] %8.0f RET %zu\n", now_us() - t0, kdone);
            ++kdone; progressed = 1;
        }
        if (!progressed) {
            /* waiting for helper threads (IN of kgpu, OUT of kdone) or for the GPU (event of kout) */
            if ((kgpu < kin && bounce_in) || (kdone < kout && bounce_out)) copy_wait_any();
            else if (kout < kgpu) CU(cudaEventSynchronize(c->rev[SLOT_OF(kout)]));
        }
    }
done:
    /* on an error pieces may still be queued for the helpers: they must land before the ring is reused */
    for (i = 0; i < MAX_RING; ++i)
        while (copy_left(&in_left[i]) || copy_left(&out_left[i])) copy_wait_any();
    for (i = 0; i < NSLOT; ++i) {
        cudaError_t e = cudaStreamSynchronize(c->st[i]);
        if (e != cudaSuccess && !rc) rc = fail(UAES_E_CUDA, "cudaStreamSynchronize(staging)", (int)e);
    }
    if (g_burn && bounce)
        for (i = 0; i < NSLOT; ++i) if (c->hslot[i]) memset(c->hslot[i], 0, CHUNK_BYTES);
    TRACE("[uaes] %8.0f done rc %d\n", now_us() - t0, rc);
    return rc;
#undef SLOT_OF
#undef DEV_OF
#undef HOST_OF
#undef STREAM_OF
#undef OFF_OF
#undef BYTES_OF
}

/* ------------------------------------------------------------------ spreading a call over devices */

/* A host-buffer call may use every GPU of the box (extension 5): the byte range is cut
 * into one contiguous part per device (block / sector ranges are independent), each part
 * runs the staging pipeline of ITS device on its own host thread, with that device's lock only.
 * uaes_set_devices(n) (or UAES_DEVICES) turns it on; a call is spread only as far as every device
 * still gets uaes_set_fanout_min() bytes. */
typedef struct {
    int dev, rc, err;
    char msg[160];
    pipe_part part;
    int (*before)(devctx *c, void *user, const pipe_part *p);    /* per-device set-up under the lock (may be NULL) */
    int (*after)(devctx *c, void *user, const pipe_part *p);     /* per-device wrap-up, pipeline drained (may be NULL) */
    pthread_t th;
    int threaded;
} fan_part;

static int fan_run_one(fan_part *f)
{
    devctx *c;
    int rc;
    if ((rc = get_ctx(&c)) != 0) return rc;
    pthread_mutex_lock(&c->lock);
    rc = f->before ? f->before(c, f->part.user, &f->part) : 0;
    if (!rc) rc = run_chunked(c, &f->part);
    if (!rc && f->after) rc = f->after(c, f->part.user, &f->part);
    pthread_mutex_unlock(&c->lock);
    return rc;
}

static void *fan_thread(void *arg)
{
    fan_part *f = (fan_part *)arg;
    cudaError_t e = cudaSetDevice(f->dev);
    tls_stream = NULL; tls_async = 0;
    f->rc = e == cudaSuccess ? fan_run_one(f) : fail(UAES_E_CUDA, "cudaSetDevice", (int)e);
    f->err = tls_err;
    memcpy(f->msg, tls_msg, sizeof f->msg);
    return NULL;
}

/* number of devices a staged call of `len` bytes is spread over */
static int fan_width(size_t len, const void *in, const void *out)
{
    int n;
    if (!g_cfg_ready) cfg_init();
    n = g_fan_devices;
    /* device-resident data stays on its device: only host memory is worth spreading */
    if (n <= 1 || ptr_class(in) == PTR_DEVICE || ptr_class(in) == PTR_OTHER ||
        ptr_class(out) == PTR_DEVICE || ptr_class(out) == PTR_OTHER) return 1;
    while (n > 1 && len / (size_t)n < g_fan_min) --n;
    return n;
}

/* Cuts `whole` into n parts at multiples of `align` bytes: part 0 for the calling thread's current
 * device, part i for device (current + i) mod count.  Returns the number of parts. */
static int fan_plan(const pipe_part *whole, int n, size_t align, fan_part f[MAX_DEV])
{
    int i, cur = 0, count = uaes_device_count();
    size_t per, off = 0;
    if (n > count && !g_fan_oversubscribe) n = count;
    if (n > MAX_DEV) n = MAX_DEV;
    if (n < 1) n = 1;
    if (n > 1 && cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); n = 1; }
    per = (whole->len / (size_t)n + align - 1) / align * align;
    for (i = 0; i < n; ++i) {
        const size_t bytes = i + 1 == n ? whole->len - off : (whole->len - off < per ? whole->len - off : per);
        memset(&f[i], 0, sizeof f[i]);
        f[i].dev = count ? (cur + i) % count : 0;
        f[i].part = *whole;
        f[i].part.in = whole->in + off;
        f[i].part.out = whole->out ? whole->out + off : NULL;
        f[i].part.len = bytes; f[i].part.base = whole->base + off;
        f[i].part.out_extra = i + 1 == n ? whole->out_extra : 0;
        f[i].part.absorb_tail = i + 1 == n ? whole->absorb_tail : 0;
        off += bytes;
    }
    return n;
}

/* part 0 on the calling thread, every other non-empty part on a thread of its own */
static int fan_exec(fan_part f[], int n)
{
    int i, rc;
    for (i = 1; i < n; ++i) {
        f[i].threaded = 0; f[i].rc = 0;
        if (f[i].part.len == 0) continue;
        if (pthread_create(&f[i].th, NULL, fan_thread, &f[i]) == 0) f[i].threaded = 1;
        else { f[i].rc = fail(UAES_E_NO_MEMORY, "pthread_create(device worker)", 0); f[i].err = tls_err; memcpy(f[i].msg, tls_msg, sizeof f[i].msg); }
    }
    rc = f[0].rc = f[0].part.len ? fan_run_one(&f[0]) : 0;
    for (i = 1; i < n; ++i) {
        if (f[i].threaded) pthread_join(f[i].th, NULL);
        if (f[i].rc && !rc) {                          /* hand the worker's error to the caller's latch */
            rc = f[i].rc; tls_err = f[i].err; memcpy(tls_msg, f[i].msg, sizeof tls_msg);
        }
    }
    return rc;
}

/* the common case: every part shares one read-only job description */
static int fan_out(const pipe_part *whole, int n, size_t align)
{
    fan_part f[MAX_DEV];
    n = fan_plan(whole, n, align, f);
    return fan_exec(f, n);
}

/* ------------------------------------------------------------------ CTR */

typedef struct { uaes_keysched ks; uaes_ctrblock cb0; } ctr_job;    /* cb0 = counter block of byte 0 of the call */

static int ctr_chunk(void *user, u64 offset, size_t index, int slot, void *dev, size_t bytes, size_t obytes, void *stream)
{
    ctr_job *j = (ctr_job *)user;
    uaes_ctrblock cb = j->cb0;
    int e;
    (void)obytes; (void)index; (void)slot;
    cb.v0 = (cb.v0 + offset / 16) & (((u64)1 << 56) - 1);
    e = uaes_launch_ctr(&j->ks, &cb, dev, dev, bytes, stream);
    return e ? fail(UAES_E_CUDA, "ctr kernel launch", e) : 0;
}

static int ctr_run(ctr_job *j, const void *in, size_t len, void *out)
{
    devctx *c;
    int rc;
    if (len == 0) return 0;                         /* NULL data is fine when there is none */
    if ((rc = get_ctx(&c)) != 0) return rc;
    if (is_direct(in) && is_direct(out)) {
        LAUNCH(uaes_launch_ctr(&j->ks, &j->cb0, in, out, len, tls_stream));
        rc = finish_direct();
    } else {
        pipe_part p;
        memset(&p, 0, sizeof p);
        p.in = (const u8 *)in; p.out = (u8 *)out; p.len = len; p.unit = 16; p.fn = ctr_chunk; p.user = j;
        /* inputs produced on the caller's stream (mixed host/device calls) must be complete */
        CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
        rc = fan_out(&p, fan_width(len, in, out), 16384);
    }
done:
    if (g_burn) memset(j, 0, sizeof *j);            /* BURN(RoundKey), micro_aes.c:975 */
    return rc;
}

int uaes_ctr_crypt_range(int keybits, const uaes_u8 *key, const uaes_u8 *iv, uaes_u64 first_block,
                         const void *in, size_t len, void *out)
{
    ctr_job j;
    if (expand_key(keybits, key, &j.ks)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    make_ctrblock(iv, 1, first_block, &j.cb0);
    return ctr_run(&j, in, len, out);
}

int uaes_ctr_crypt(int keybits, const uaes_u8 *key, const uaes_u8 *iv, const void *in, size_t len, void *out)
{
    return uaes_ctr_crypt_range(keybits, key, iv, 0, in, len, out);
}

/* the reference built with PRESET_COUNTER = 1 (micro_aes.c:964-966): the caller's 16 bytes ARE
 * counter block 0; block k adds k to the 56-bit big-endian field in bytes 9..15 */
int uaes_ctr_crypt_block(int keybits, const uaes_u8 *key, const uaes_u8 *ctr, uaes_u64 first_block,
                         const void *in, size_t len, void *out)
{
    ctr_job j;
    if (expand_key(keybits, key, &j.ks)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    block_to_ctrblock(ctr, first_block, &j.cb0);
    return ctr_run(&j, in, len, out);
}

/* ------------------------------------------------------------------ ECB */

typedef struct { uaes_keysched ks; int encrypt, pad; size_t total; } ecb_job;

static int ecb_chunk(void *user, u64 offset, size_t index, int slot, void *dev, size_t bytes, size_t obytes, void *stream)
{
    ecb_job *j = (ecb_job *)user;
    int e;
    (void)obytes; (void)index; (void)slot;
    /* only the chunk that ends the message is padded (PKCS#7 / ISO 7816 always add a block there) */
    e = uaes_launch_ecb(&j->ks, j->encrypt, offset + bytes == j->total ? j->pad : 0, dev, dev, bytes, stream);
    return e ? fail(UAES_E_CUDA, "ecb kernel launch", e) : 0;
}

static int ecb_common(int keybits, const u8 *key, const void *in, size_t len, void *out, int encrypt, int pad)
{
    devctx *c;
    ecb_job j;
    uaes_keysched enc;
    int rc;
    if (expand_key(keybits, key, &enc)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    if (pad < 0 || pad > 2) return fail(UAES_E_BAD_ARGUMENT, "padding must be 0 (zeros), 1 (PKCS#7) or 2 (ISO/IEC 7816-4)", 0);
    if (encrypt) j.ks = enc; else invert_schedule(&enc, &j.ks);
    j.encrypt = encrypt; j.pad = encrypt ? pad : 0; j.total = len;
    if (len == 0 && !j.pad) return 0;
    if ((rc = get_ctx(&c)) != 0) return rc;
    if (len == 0) {
        /* an empty message still encrypts to one full padding block: stage that block alone */
        u8 *d;
        pthread_mutex_lock(&c->lock);
        if ((rc = need_slots(c, 0)) == 0) {
            d = (u8 *)c->slot[0];
            rc = uaes_launch_ecb(&j.ks, 1, j.pad, d, d, 0, c->st[0]);
            if (rc) rc = fail(UAES_E_CUDA, "ecb kernel launch", rc);
            else if (cudaMemcpyAsync(out, d, 16, cudaMemcpyDefault, c->st[0]) != cudaSuccess ||
                     cudaStreamSynchronize(c->st[0]) != cudaSuccess)
                rc = fail(UAES_E_CUDA, "cudaMemcpy(padding block)", (int)cudaGetLastError());
        }
        pthread_mutex_unlock(&c->lock);
        return rc;
    }
    if (is_direct(in) && is_direct(out)) {
        LAUNCH(uaes_launch_ecb(&j.ks, encrypt, j.pad, in, out, len, tls_stream));
        rc = finish_direct();
    } else {
        /* encrypt pads the tail to a whole block: the last chunk returns up to 16 more bytes */
        pipe_part p;
        memset(&p, 0, sizeof p);
        p.in = (const u8 *)in; p.out = (u8 *)out; p.len = len; p.unit = 16; p.fn = ecb_chunk; p.user = &j;
        p.out_extra = !encrypt ? 0 : j.pad ? 16 - len % 16 : (16 - len % 16) % 16;
        CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
        rc = fan_out(&p, fan_width(len, in, out), 16384);
    }
done:
    if (g_burn) { memset(&j, 0, sizeof j); memset(&enc, 0, sizeof enc); }
    return rc;
}

int uaes_ecb_encrypt(int keybits, const uaes_u8 *key, const void *in, size_t len, void *out)
{
    return ecb_common(keybits, key, in, len, out, 1, 0);
}

/* AES_PADDING = 1 / 2 of the reference (micro_aes.h:78-80): out holds (len / 16 + 1) * 16 bytes */
int uaes_ecb_encrypt_padded(int keybits, const uaes_u8 *key, const void *in, size_t len, void *out, int padding)
{
    return ecb_common(keybits, key, in, len, out, 1, padding);
}

int uaes_ecb_decrypt(int keybits, const uaes_u8 *key, const void *in, size_t len, void *out)
{
    const int rc = ecb_common(keybits, key, in, len, out, 0, 0);
    if (rc) return rc;
    return len % 16 ? UAES_DECRYPTION_ERROR : UAES_OK;       /* micro_aes.c:679 */
}

/* ------------------------------------------------------------------ XTS */

typedef struct {
    uaes_keysched k1, k1e, k2;
    int encrypt;
    u64 first_sector; size_t sector_bytes;          /* sector batches */
    u8 tweak[16]; u64 first_block;                  /* one data unit, or a range of it */
} xts_job;

static int xts_keys(int keybits, const u8 *keys, int encrypt, xts_job *j)
{
    /* the reference runs XTS with whatever AES___ it was built for, 192 included (two 24-byte keys) */
    if (expand_key(keybits, keys, &j->k1e))                   /* key1 = first half: data key   */
        return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    expand_key(keybits, keys + keybits / 8, &j->k2);          /* key2 = second half: tweak key */
    if (encrypt) j->k1 = j->k1e; else invert_schedule(&j->k1e, &j->k1);
    j->encrypt = encrypt;
    return 0;
}

static int xts_chunk(void *user, u64 offset, size_t index, int slot, void *dev, size_t bytes, size_t obytes, void *stream)
{
    xts_job *j = (xts_job *)user;
    int e;
    (void)obytes; (void)index; (void)slot;
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
    if (sector_bytes < 16 || sector_bytes % 16 || len % sector_bytes)
        return fail(UAES_E_BAD_ARGUMENT, "sector size must be a multiple of 16 and divide the length", 0);
    if (len == 0) return 0;
    j.first_sector = first_sector; j.sector_bytes = sector_bytes;
    if ((rc = get_ctx(&c)) != 0) return rc;
    if (is_direct(in) && is_direct(out)) {
        LAUNCH(uaes_launch_xts_sectors(&j.k1, &j.k2, encrypt, first_sector, sector_bytes / 16,
                                       len / sector_bytes, in, out, tls_stream));
        rc = finish_direct();
    } else {
        pipe_part p;
        if (sector_bytes > CHUNK_BYTES)
            return fail(UAES_E_BAD_ARGUMENT, "staged sectors must fit one staging chunk; use uaes_xts_encrypt per unit", 0);
        memset(&p, 0, sizeof p);
        p.in = (const u8 *)in; p.out = (u8 *)out; p.len = len; p.unit = sector_bytes; p.fn = xts_chunk; p.user = &j;
        CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
        rc = fan_out(&p, fan_width(len, in, out), sector_bytes);
    }
done:
    if (g_burn) memset(&j, 0, sizeof j);
    return rc;
}

static int xts_unit_chunk(void *user, u64 offset, size_t index, int slot, void *dev, size_t bytes, size_t obytes, void *stream)
{
    xts_job *j = (xts_job *)user;
    int e;
    (void)obytes; (void)index; (void)slot;
    e = uaes_launch_xts_unit(&j->k1, &j->k1e, &j->k2, j->encrypt, j->tweak, j->first_block + offset / 16,
                             dev, dev, bytes, stream);
    return e ? fail(UAES_E_CUDA, "xts kernel launch", e) : 0;
}

/* One data unit or a block range of it.  The tweak chain of the reference (micro_aes.c:1030-1036) is
 * T_0 * alpha^k by jump-ahead, so a unit can be cut anywhere at a block boundary: staged in chunks,
 * or spread over GPUs.  Only the range that ends the unit may be ragged (stealing). */
static int xts_unit(int keybits, const u8 *keys, const u8 *tweak, u64 first_block, const void *in, size_t len,
                    void *out, int encrypt)
{
    devctx *c;
    xts_job j;
    int rc;
    if (len < 16) return UAES_DATALENGTH_ERROR;               /* micro_aes.c:1069, 1088 */
    if ((rc = xts_keys(keybits, keys, encrypt, &j)) != 0) return rc;
    memset(j.tweak, 0, 16);                                   /* NULL = sector 0, micro_aes.c:1017-1021 */
    if (tweak) memcpy(j.tweak, tweak, 16);
    j.first_block = first_block;
    if ((rc = get_ctx(&c)) != 0) return rc;
    if (is_direct(in) && is_direct(out)) {
        LAUNCH(uaes_launch_xts_unit(&j.k1, &j.k1e, &j.k2, encrypt, j.tweak, first_block, in, out, len, tls_stream));
        rc = finish_direct();
    } else {
        pipe_part p;
        memset(&p, 0, sizeof p);
        p.in = (const u8 *)in; p.out = (u8 *)out; p.len = len; p.unit = 16; p.fn = xts_unit_chunk; p.user = &j;
        p.absorb_tail = 32;                       /* a last piece of < 32 bytes joins the chunk before it */
        CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
        rc = fan_out(&p, fan_width(len, in, out), 16384);
    }
done:
    if (g_burn) memset(&j, 0, sizeof j);
    return rc;
}

int uaes_xts_encrypt(int keybits, const uaes_u8 *keys, const uaes_u8 *tweak, const void *in, size_t len, void *out)
{
    return xts_unit(keybits, keys, tweak, 0, in, len, out, 1);
}

int uaes_xts_decrypt(int keybits, const uaes_u8 *keys, const uaes_u8 *tweak, const void *in, size_t len, void *out)
{
    return xts_unit(keybits, keys, tweak, 0, in, len, out, 0);
}

int uaes_xts_crypt_range(int keybits, const uaes_u8 *keys, const uaes_u8 *tweak, uaes_u64 first_block,
                         const void *in, size_t len, void *out, int encrypt)
{
    return xts_unit(keybits, keys, tweak, first_block, in, len, out, encrypt);
}

/* ------------------------------------------------------------------ GCM */

#define GCM_WORK_HEAD 1024   /* tag scratch lives in front of the kernels' work area */
#define AAD_STATE_OFF 128    /* 16 bytes inside the head: GHASH state of a bulk-hashed AAD */
#define AAD_BULK_MIN  4096   /* larger AADs are hashed by the bulk kernel instead of one lane */

/* J0 (micro_aes.c:1140-1152): nonce || 00000001 for the recommended 12-byte nonce; any other length
 * is GHASH_H({}, nonce), computed on the device and read back (16 bytes) because the launch
 * arguments of the bulk kernel are derived from it */
static int gcm_j0(devctx *c, const uaes_keysched *ks, const u8 *nonce, size_t noncelen, u8 j0[16])
{
    int rc = 0;
    scratch *w = NULL;
    cudaStream_t st = (cudaStream_t)tls_stream;
    if (noncelen == 12) {
        memcpy(j0, nonce, 12);
        j0[12] = j0[13] = j0[14] = 0; j0[15] = 1;
        return 0;
    }
    if (noncelen == 0) return fail(UAES_E_BAD_ARGUMENT, "GCM nonce must not be empty", 0);
    if ((rc = scratch_get(c, noncelen + 64, st, &w)) != 0) return rc;
    CU(cudaMemcpyAsync((u8 *)w->p + 32, nonce, noncelen, cudaMemcpyDefault, st));
    LAUNCH(uaes_launch_gcm_j0(ks, (u8 *)w->p + 32, noncelen, w->p, st));
    CU(cudaMemcpyAsync(j0, w->p, 16, cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
done:
    scratch_put(c, w, st);
    return rc;
}

/* ---- staged GCM: the message goes through the chunk pipeline as a sequence of shards ------------
 * Every chunk is one fused CTR+GHASH pass that leaves its 16-byte contribution on the device; the
 * contributions of all chunks (of all devices) are folded with powers of H into the tag at the end
 * (gcm_combine_kernel).  H2D, kernel and D2H of neighbouring chunks overlap exactly as for CTR.
 * Decryption must not hand out plaintext before the tag is checked (micro_aes.c:1204-1208): its
 * chunks decrypt into a device-resident copy of the part, and only a matching tag starts the copy
 * back to the caller. */
typedef struct {
    uaes_keysched ks;
    u8 j0[16];
    int mode;                    /* 0 encrypt; 2 decrypt (CTR + GHASH of the input) */
    u64 total_len;
    size_t chunk;                /* pipeline chunk of this call */
    scratch *w;                  /* NSLOT work areas + the part's contribution array (device) */
    size_t work_stride, nchunks, nslots;
    u8 *host_parts;              /* where this part's contributions go on the host (16 B per chunk) */
    u64 *host_after;             /* ... and the number of GHASH blocks after each chunk */
    u8 *keep;                    /* decrypt: the part's plaintext on the device until the tag is checked */
} gcm_part;

static int gcm_part_before(devctx *c, void *user, const pipe_part *p)
{
    gcm_part *g = (gcm_part *)user;
    g->nchunks = pipe_nchunks(p);
    g->work_stride = (uaes_gcm_work_bytes(pipe_chunk_bytes(p)) + 255) & ~(size_t)255;
    g->nslots = (size_t)pipe_slots(p);
    return scratch_get(c, g->nslots * g->work_stride + g->nchunks * 16 + 64, c->st[0], &g->w);
}

static int gcm_part_chunk(void *user, u64 offset, size_t index, int slot, void *dev, size_t bytes, size_t obytes, void *stream)
{
    gcm_part *g = (gcm_part *)user;
    u8 *work = (u8 *)g->w->p + (size_t)slot * g->work_stride;
    u8 *dpart = (u8 *)g->w->p + g->nslots * g->work_stride + index * 16;
    int e;
    (void)obytes;
    e = uaes_launch_gcm(&g->ks, g->j0, NULL, 0, NULL, dev, dev, bytes, g->mode, offset / 16, 1, dpart, 16, work, stream);
    if (e) return fail(UAES_E_CUDA, "gcm kernel launch", e);
    g->host_after[index] = (g->total_len + 15) / 16 - (offset + bytes + 15) / 16;
    return 0;
}

static int gcm_part_after(devctx *c, void *user, const pipe_part *p)
{
    gcm_part *g = (gcm_part *)user;
    int rc = 0;
    (void)p;
    /* the pipeline is drained: the contributions are complete */
    if (g->nchunks &&
        cudaMemcpy(g->host_parts, (u8 *)g->w->p + g->nslots * g->work_stride, g->nchunks * 16, cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = fail(UAES_E_CUDA, "cudaMemcpy(GHASH contributions)", (int)cudaGetLastError());
    scratch_put(c, g->w, c->st[0]);
    g->w = NULL;
    return rc;
}

static int gcm_keep_free(devctx *c, void *user, const pipe_part *p)
{
    gcm_part *g = (gcm_part *)user;
    (void)c;
    if (g->keep) {
        if (g_burn) cudaMemset(g->keep, 0, p->len);
        cudaFree(g->keep);
        g->keep = NULL;
    }
    return 0;
}

/* folds contributions (n x 16 B and n x u64, host or device memory) + AAD + lengths into the tag
 * (taglen bytes written to `tag`, host or device memory) on the current device */
static int gcm_fold_tag(devctx *c, const uaes_keysched *ks, const u8 j0[16], const void *aad, size_t aadlen,
                        const void *parts, const void *after, size_t n, u64 total_len, u8 *tag, size_t taglen)
{
    int rc = 0, bulk_aad = aadlen >= AAD_BULK_MIN;
    scratch *w = NULL;
    u8 *base, *dparts, *dafter, *daad;
    cudaStream_t st = (cudaStream_t)tls_stream;
    const size_t aoff = (GCM_WORK_HEAD + (n + 1) * 24 + 255) & ~(size_t)255;
    const size_t woff = aoff + ((aadlen + 255) & ~(size_t)255);

    if ((rc = scratch_get(c, woff + 64 + (bulk_aad ? uaes_gcm_work_bytes(aadlen) + 256 : 0), st, &w)) != 0) return rc;
    base = (u8 *)w->p;                   /* [tag 16][pad][parts (n+1) x 16][after (n+1) x 8][aad][aad work] */
    dparts = base + GCM_WORK_HEAD; dafter = dparts + (n + 1) * 16; daad = base + aoff;
    if (n) {
        CU(cudaMemcpyAsync(dparts, parts, n * 16, cudaMemcpyDefault, st));
        CU(cudaMemcpyAsync(dafter, after, n * 8, cudaMemcpyDefault, st));
    }
    if (aadlen) CU(cudaMemcpyAsync(daad, aad, aadlen, cudaMemcpyDefault, st));
    if (bulk_aad) {
        /* a large AAD is one more shard: hashed by the bulk kernel, it ends total_blocks before the end */
        const u64 tb = (total_len + 15) / 16;
        LAUNCH(uaes_launch_gcm(ks, j0, NULL, 0, NULL, daad, NULL, aadlen, 1, 0, 1, dparts + n * 16, 16, base + woff, st));
        CU(cudaMemcpyAsync(dafter + n * 8, &tb, 8, cudaMemcpyDefault, st));
        CU(cudaStreamSynchronize(st));                          /* tb lives on this stack frame */
    }
    LAUNCH(uaes_launch_gcm_combine(ks, j0, bulk_aad || !aadlen ? NULL : daad, aadlen, total_len, dparts, dafter,
                                   (unsigned)(n + (bulk_aad ? 1 : 0)), base, 16, st));
    CU(cudaMemcpyAsync(tag, base, taglen, cudaMemcpyDefault, st));
    /* a tag that stays on the device (and contributions that came from there) needs no host round trip:
     * in asynchronous mode the call only enqueues, like the shard calls that feed it */
    if (!(tls_async && (ptr_class(tag) == PTR_DEVICE || ptr_class(tag) == PTR_OTHER))) CU(cudaStreamSynchronize(st));
done:
    scratch_put(c, w, st);
    return rc;
}

static int gcm_staged(devctx *c, const uaes_keysched *ks, const u8 j0[16], const void *aad, size_t aadlen,
                      const void *in, size_t len, void *out, size_t taglen, int decrypt)
{
    fan_part f[MAX_DEV];
    gcm_part g[MAX_DEV];
    pipe_part whole;
    int rc = 0, n, i;
    size_t total_chunks = 0, k = 0;
    u8 *parts = NULL, tag[16], rtag[16];
    u64 *after = NULL;

    memset(&whole, 0, sizeof whole);
    whole.in = (const u8 *)in; whole.out = decrypt ? NULL : (u8 *)out; whole.len = len; whole.unit = 16;
    whole.fn = gcm_part_chunk;
    n = fan_plan(&whole, fan_width(len, in, out), 16384, f);
    for (i = 0; i < n; ++i) total_chunks += pipe_nchunks(&f[i].part);
    parts = (u8 *)malloc(total_chunks * 16 + 16);
    after = (u64 *)malloc(total_chunks * 8 + 8);
    if (!parts || !after) { rc = fail(UAES_E_NO_MEMORY, "malloc(GHASH contributions)", 0); goto done; }
    memset(g, 0, sizeof g);
    for (i = 0; i < n; ++i) {
        g[i].ks = *ks; memcpy(g[i].j0, j0, 16);
        g[i].mode = decrypt ? 2 : 0; g[i].total_len = len; g[i].chunk = pipe_chunk_bytes(&whole);
        g[i].host_parts = parts + 16 * k; g[i].host_after = after + k;
        k += pipe_nchunks(&f[i].part);
        f[i].part.user = &g[i];
        f[i].before = gcm_part_before; f[i].after = gcm_part_after;
        if (decrypt && f[i].part.len) {
            /* the call owns this copy (not the shared staging buffer: the device lock is released
             * between the two phases) */
            cudaError_t e = cudaSetDevice(f[i].dev);
            if (e == cudaSuccess) e = cudaMalloc((void **)&g[i].keep, f[i].part.len + 32);
            if (e != cudaSuccess) { rc = fail(UAES_E_NO_MEMORY, "cudaMalloc(GCM decrypt copy)", (int)e); break; }
            f[i].part.resident = g[i].keep;
        }
    }
    if (decrypt) cudaSetDevice(f[0].dev);
    if (rc) goto cleanup;
    CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
    if ((rc = fan_exec(f, n)) != 0) goto cleanup;
    if ((rc = gcm_fold_tag(c, ks, j0, aad, aadlen, parts, after, total_chunks, len, tag, 16)) != 0) goto cleanup;
    if (!decrypt) {
        /* tag appended at out + len (micro_aes.c:1168, 1178) */
        if (cudaMemcpy((u8 *)out + len, tag, taglen, cudaMemcpyDefault) != cudaSuccess)
            rc = fail(UAES_E_CUDA, "cudaMemcpy(tag)", (int)cudaGetLastError());
        goto cleanup;
    }
    if (cudaMemcpy(rtag, (const u8 *)in + len, taglen, cudaMemcpyDefault) != cudaSuccess) {
        rc = fail(UAES_E_CUDA, "cudaMemcpy(received tag)", (int)cudaGetLastError());
        goto cleanup;
    }
    if (memcmp(tag, rtag, taglen)) { rc = UAES_AUTH_ERROR; goto cleanup; }   /* out untouched, micro_aes.c:1204-1208 */
    /* phase 2: the verified plaintext leaves the devices */
    for (i = 0; i < n; ++i) {
        f[i].part.in = g[i].keep; f[i].part.out = (u8 *)out + (f[i].part.base - whole.base);
        f[i].part.resident = NULL; f[i].part.fn = NULL;
        f[i].before = NULL; f[i].after = gcm_keep_free;
    }
    rc = fan_exec(f, n);
cleanup:
    for (i = 0; i < n; ++i)
        if (g[i].keep) { if (g_burn) cudaMemset(g[i].keep, 0, f[i].part.len); cudaFree(g[i].keep); g[i].keep = NULL; }
done:
    free(parts); free(after);
    if (g_burn) memset(g, 0, sizeof g);
    return rc;
}

/* The whole message on one device: device-resident buffers (zero copies), or a host message of at
 * most one staging chunk. */
static int gcm_common(int keybits, const u8 *key, const u8 *nonce, size_t noncelen, const void *aad, size_t aadlen,
                      const void *in, size_t len, void *out, size_t taglen, int decrypt)
{
    devctx *c;
    uaes_keysched ks;
    int rc, direct, locked = 0;
    const void *din, *daad;
    void *dout;
    u8 *work, *dtag, *dstate, j0[16];
    size_t wbytes;
    scratch *w = NULL;
    cudaStream_t st = NULL;

    if (expand_key(keybits, key, &ks)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    if (taglen < 1 || taglen > 16) return fail(UAES_E_BAD_ARGUMENT, "GCM tag length must be 1..16 bytes", 0);
    if ((rc = get_ctx(&c)) != 0) return rc;
    if ((rc = gcm_j0(c, &ks, nonce, noncelen, j0)) != 0) return rc;

    /* the kernels write ciphertext and tag through `out` and read `in`: both must be device memory
     * for the zero-copy path (an empty message still has a tag to write when encrypting) */
    if (decrypt) direct = len == 0 || (is_direct(in) && is_direct(out));
    else         direct = is_direct(out) && (len == 0 || is_direct(in));
    if (!direct) {
        /* more than one pipeline chunk (64 MiB pinned, 8 MiB pageable): one shard per chunk, copies overlapped */
        pipe_part pp;
        memset(&pp, 0, sizeof pp);
        pp.in = (const u8 *)in; pp.out = (u8 *)out; pp.len = len; pp.unit = 16;
        if (len > pipe_chunk_bytes(&pp)) {
            rc = gcm_staged(c, &ks, j0, aad, aadlen, in, len, out, taglen, decrypt);
            goto wipe;
        }
    }
    if (!direct) { pthread_mutex_lock(&c->lock); locked = 1; }     /* the staged path owns big and st[0] */
    st = direct ? (cudaStream_t)tls_stream : c->st[0];
    wbytes = GCM_WORK_HEAD + uaes_gcm_work_bytes(len > aadlen ? len : aadlen) + aadlen + 64;
    if ((rc = scratch_get(c, wbytes, st, &w)) != 0) goto done;
    dtag = (u8 *)w->p;
    work = (u8 *)w->p + GCM_WORK_HEAD;
    daad = NULL;
    if (aadlen) {
        u8 *a = work + uaes_gcm_work_bytes(len > aadlen ? len : aadlen);
        a += (16 - ((size_t)a & 15)) & 15;
        /* a staged call's AAD copy runs on st[0]: order it after the caller's stream first */
        if (!direct) CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
        CU(cudaMemcpyAsync(a, aad, aadlen, cudaMemcpyDefault, st));
        daad = a;
    }
    if (direct) {
        din = in; dout = out;
    } else {
        if ((rc = grow(&c->big, &c->big_bytes, len + 32, "cudaMalloc(GCM staging)")) != 0) goto done;
        CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
        CU(cudaMemcpyAsync(c->big, in, len + (decrypt ? taglen : 0), cudaMemcpyDefault, st));
        din = c->big; dout = c->big;
    }
    dstate = NULL;
    if (aadlen >= AAD_BULK_MIN) {                             /* xMac over the AAD (micro_aes.c:1134) in bulk */
        dstate = (u8 *)w->p + AAD_STATE_OFF;
        LAUNCH(uaes_launch_gcm(&ks, j0, NULL, 0, NULL, daad, NULL, aadlen, 1, 0, 1, dstate, 16, work, st));
    }

    if (!decrypt) {
        /* one fused pass: CTR + GHASH, tag appended at out + len (micro_aes.c:1168,1178) */
        LAUNCH(uaes_launch_gcm(&ks, j0, daad, aadlen, dstate, din, dout, len, 0, 0, 0, (u8 *)dout + len, (unsigned)taglen, work, st));
        if (!direct) CU(cudaMemcpyAsync(out, c->big, len + taglen, cudaMemcpyDefault, st));
        if (!direct || !tls_async) CU(cudaStreamSynchronize(st));
    } else {
        /* verify first, decrypt only on success; `out` stays untouched otherwise (micro_aes.c:1199-1209) */
        u8 t1[16], t2[16];
        uaes_ctrblock cb;
        LAUNCH(uaes_launch_gcm(&ks, j0, daad, aadlen, dstate, din, NULL, len, 1, 0, 0, dtag, 16, work, st));
        CU(cudaMemcpyAsync(t1, dtag, 16, cudaMemcpyDefault, st));
        CU(cudaMemcpyAsync(t2, (const u8 *)din + len, taglen, cudaMemcpyDefault, st));
        CU(cudaStreamSynchronize(st));
        if (memcmp(t1, t2, taglen)) { rc = UAES_AUTH_ERROR; goto done; }
        if (len) {
            block_to_ctrblock(j0, 1, &cb);                    /* data from J0 + 1 (micro_aes.c:939-941) */
            LAUNCH(uaes_launch_ctr(&ks, &cb, din, dout, len, st));
            if (!direct) CU(cudaMemcpyAsync(out, c->big, len, cudaMemcpyDefault, st));
            if (!direct || !tls_async) CU(cudaStreamSynchronize(st));
        }
    }
done:
    scratch_put(c, w, st);
    if (locked) { big_done(c); pthread_mutex_unlock(&c->lock); }
wipe:
    if (g_burn) { memset(&ks, 0, sizeof ks); memset(j0, 0, sizeof j0); }
    return rc;
}

int uaes_gcm_encrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return gcm_common(keybits, key, nonce, 12, aad, aadlen, in, len, out, 16, 0);
}

int uaes_gcm_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return gcm_common(keybits, key, nonce, 12, aad, aadlen, in, len, out, 16, 1);
}

/* the reference's compile-time GCM_NONCE_LEN / GCM_TAG_LEN (micro_aes.h:107-110) as run-time arguments */
int uaes_gcm_encrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, size_t noncelen,
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen)
{
    return gcm_common(keybits, key, nonce, noncelen, aad, aadlen, in, len, out, taglen, 0);
}

int uaes_gcm_decrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, size_t noncelen,
                        const void *aad, size_t aadlen, const void *in, size_t len, void *out, size_t taglen)
{
    return gcm_common(keybits, key, nonce, noncelen, aad, aadlen, in, len, out, taglen, 1);
}

/* ---- a GCM message sharded over several GPUs / calls ---------------------------
 * Every shard runs the fused CTR+GHASH pass over its own byte range and returns 16 bytes; one
 * caller gathers them (an all-gather of 16 B per rank) and folds them into the tag.  `partial` may
 * be DEVICE memory: the contribution then stays on the GPU (ready for a device-to-device gather)
 * and, in asynchronous mode, the call only enqueues work. */
int uaes_gcm_shard(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, uaes_u64 first_block,
                   const void *in, size_t len, void *out, int decrypt, uaes_u8 *partial)
{
    devctx *c;
    uaes_keysched ks;
    int rc, direct, locked = 0, part_dev;
    const void *din;
    void *dout;
    u8 *work, *dpart, j0[16];
    scratch *w = NULL;
    cudaStream_t st = NULL;

    if (expand_key(keybits, key, &ks)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    if ((rc = get_ctx(&c)) != 0) return rc;
    memcpy(j0, nonce, 12); j0[12] = j0[13] = j0[14] = 0; j0[15] = 1;
    direct = len == 0 || (is_direct(in) && is_direct(out));
    part_dev = ptr_class(partial) == PTR_DEVICE || ptr_class(partial) == PTR_OTHER;
    if (!direct) { pthread_mutex_lock(&c->lock); locked = 1; }
    st = direct ? (cudaStream_t)tls_stream : c->st[0];
    if ((rc = scratch_get(c, GCM_WORK_HEAD + uaes_gcm_work_bytes(len) + 64, st, &w)) != 0) goto done;
    dpart = part_dev ? partial : (u8 *)w->p;
    work = (u8 *)w->p + GCM_WORK_HEAD;
    if (direct) {
        din = in; dout = out;
    } else {
        if ((rc = grow(&c->big, &c->big_bytes, len + 32, "cudaMalloc(GCM staging)")) != 0) goto done;
        CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
        CU(cudaMemcpyAsync(c->big, in, len, cudaMemcpyDefault, st));
        din = c->big; dout = c->big;
    }
    LAUNCH(uaes_launch_gcm(&ks, j0, NULL, 0, NULL, din, dout, len, decrypt ? 2 : 0, first_block, 1, dpart, 16, work, st));
    if (!direct) CU(cudaMemcpyAsync(out, c->big, len, cudaMemcpyDefault, st));
    if (!part_dev) CU(cudaMemcpyAsync(partial, dpart, 16, cudaMemcpyDefault, st));
    if (!part_dev || !direct || !tls_async) CU(cudaStreamSynchronize(st));     /* host results are complete on return */
done:
    scratch_put(c, w, st);
    if (locked) { big_done(c); pthread_mutex_unlock(&c->lock); }
    if (g_burn) memset(&ks, 0, sizeof ks);
    return rc;
}

/* partials (nshards x 16 B), blocks_after (nshards x u64) and tag may each be host or device memory */
int uaes_gcm_combine(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const uaes_u8 *partials, const uaes_u64 *blocks_after, int nshards,
                     uaes_u64 total_len, uaes_u8 *tag)
{
    devctx *c;
    uaes_keysched ks;
    int rc;
    u8 j0[16];
    if (expand_key(keybits, key, &ks)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    if (nshards < 0) return fail(UAES_E_BAD_ARGUMENT, "negative shard count", 0);
    if ((rc = get_ctx(&c)) != 0) return rc;
    memcpy(j0, nonce, 12); j0[12] = j0[13] = j0[14] = 0; j0[15] = 1;
    rc = gcm_fold_tag(c, &ks, j0, aad, aadlen, partials, blocks_after, (size_t)nshards, total_len, tag, 16);
    if (g_burn) memset(&ks, 0, sizeof ks);
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
    int rc, direct, locked = 0;
    const void *din, *daad;
    void *dout;
    u8 *work, *dtag, *dderived, *dstate, derived[48];
    size_t wbytes;
    scratch *w = NULL;
    cudaStream_t st;

    if (expand_key(keybits, key, &master)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    if ((rc = get_ctx(&c)) != 0) return rc;

    if (decrypt) direct = len == 0 || (is_direct(in) && is_direct(out));
    else         direct = is_direct(out) && (len == 0 || is_direct(in));
    if (!direct) { pthread_mutex_lock(&c->lock); locked = 1; }     /* the staged path owns big and st[0] */
    st = direct ? (cudaStream_t)tls_stream : c->st[0];
    wbytes = GCM_WORK_HEAD + uaes_gcm_work_bytes(len > aadlen ? len : aadlen) + aadlen + 64;
    if ((rc = scratch_get(c, wbytes, st, &w)) != 0) goto done;
    dtag = (u8 *)w->p;                    /* [0,16) computed tag, [64,112) derived key material */
    dderived = (u8 *)w->p + 64;
    work = (u8 *)w->p + GCM_WORK_HEAD;

    /* message keys: E_K(LE32(i) || nonce) on the device, KeyExpansion of the result on the host */
    LAUNCH(uaes_launch_gcmsiv_derive(&master, nonce, dderived, st));
    CU(cudaMemcpyAsync(derived, dderived, 48, cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
    expand_key(keybits, derived + 16, &enc);

    daad = NULL;
    if (aadlen) {
        u8 *a = work + uaes_gcm_work_bytes(len > aadlen ? len : aadlen);
        a += (16 - ((size_t)a & 15)) & 15;
        if (!direct) CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
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
        dstate = (u8 *)w->p + AAD_STATE_OFF;
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
    scratch_put(c, w, st);
    if (locked) { big_done(c); pthread_mutex_unlock(&c->lock); }
    if (g_burn) { memset(&master, 0, sizeof master); memset(&enc, 0, sizeof enc); memset(derived, 0, sizeof derived); }
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
static int chain_common(int keybits, const u8 *key, const u8 *iv, const void *in, size_t len, void *out, int cbc, int cts)
{
    devctx *c;
    uaes_keysched enc, dec;
    int rc, locked = 0;
    size_t n = len / 16, r = len % 16;
    cudaStream_t st;

    if (expand_key(keybits, key, &enc)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    if (cbc && cts) {                                         /* CS3 rules of micro_aes.c:751-760 */
        if (n > 1 && !r) { --n; r = 16; }
        if (n == 0) return UAES_DATALENGTH_ERROR;
        n -= r > 0;
        invert_schedule(&enc, &dec);
    } else if (cbc) {                                         /* built with CTS = 0: whole blocks only, micro_aes.c:757-759 */
        if (r) return UAES_DATALENGTH_ERROR;
        if (n == 0) return 0;
        invert_schedule(&enc, &dec);
    } else if (len == 0) {
        return 0;
    }
    if ((rc = get_ctx(&c)) != 0) return rc;
    if (is_direct(in) && is_direct(out) && in != out) {
        LAUNCH(uaes_launch_chain_dec(cbc ? &dec : &enc, &enc, cbc, iv, in, out, n, (unsigned)r, tls_stream));
        rc = finish_direct();
    } else {
        const size_t half = (len + 255) & ~(size_t)255;
        u8 *din, *dout;
        pthread_mutex_lock(&c->lock); locked = 1;
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
    if (locked) { big_done(c); pthread_mutex_unlock(&c->lock); }
    if (g_burn) { memset(&enc, 0, sizeof enc); memset(&dec, 0, sizeof dec); }
    return rc;
}

int uaes_cbc_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *iv, const void *in, size_t len, void *out)
{
    return chain_common(keybits, key, iv, in, len, out, 1, 1);
}

/* cts = 0: the reference built with CTS = 0 -- plain CBC, UAES_DATALENGTH_ERROR unless len % 16 == 0
 * (micro_aes.c:757-759) */
int uaes_cbc_decrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *iv, const void *in, size_t len, void *out, int cts)
{
    return chain_common(keybits, key, iv, in, len, out, 1, cts);
}

int uaes_cfb_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *iv, const void *in, size_t len, void *out)
{
    return chain_common(keybits, key, iv, in, len, out, 0, 0);
}

/* ------------------------------------------------------------------ OCB */

/* micro_aes.c:1779-1813.  One pass; decrypt writes the plaintext and then reports the tag
 * comparison, exactly like the reference (micro_aes.c:1806-1812). */
static int ocb_common(int keybits, const u8 *key, const u8 *nonce, const void *aad, size_t aadlen,
                      const void *in, size_t len, void *out, size_t taglen, int decrypt)
{
    devctx *c;
    uaes_keysched enc, dec;
    int rc, direct, locked = 0;
    const void *din, *daad;
    void *dout;
    u8 *work, *dtag;
    scratch *w = NULL;
    cudaStream_t st;

    if (expand_key(keybits, key, &enc)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    if (taglen < 1 || taglen > 16) return fail(UAES_E_BAD_ARGUMENT, "OCB tag length must be 1..16 bytes", 0);
    if (decrypt) invert_schedule(&enc, &dec);
    if ((rc = get_ctx(&c)) != 0) return rc;
    if (decrypt) direct = len == 0 || (is_direct(in) && is_direct(out));
    else         direct = is_direct(out) && (len == 0 || is_direct(in));
    if (!direct) { pthread_mutex_lock(&c->lock); locked = 1; }
    st = direct ? (cudaStream_t)tls_stream : c->st[0];
    if ((rc = scratch_get(c, GCM_WORK_HEAD + uaes_ocb_work_bytes() + aadlen + 64, st, &w)) != 0) goto done;
    dtag = (u8 *)w->p;
    work = (u8 *)w->p + GCM_WORK_HEAD;
    daad = NULL;
    if (aadlen) {
        u8 *a = work + ((uaes_ocb_work_bytes() + 15) & ~(size_t)15);
        if (!direct) CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
        CU(cudaMemcpyAsync(a, aad, aadlen, cudaMemcpyDefault, st));
        daad = a;
    }
    if (direct) {
        din = in; dout = out;
    } else {
        if ((rc = grow(&c->big, &c->big_bytes, len + 32, "cudaMalloc(OCB staging)")) != 0) goto done;
        CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
        CU(cudaMemcpyAsync(c->big, in, len + (decrypt ? taglen : 0), cudaMemcpyDefault, st));
        din = c->big; dout = c->big;
    }
    if (!decrypt) {
        LAUNCH(uaes_launch_ocb(&enc, &enc, 1, nonce, daad, aadlen, din, dout, len, (u8 *)dout + len, (unsigned)taglen, work, st));
        if (!direct) CU(cudaMemcpyAsync(out, c->big, len + taglen, cudaMemcpyDefault, st));
        if (!direct || !tls_async) CU(cudaStreamSynchronize(st));
    } else {
        u8 t1[16], t2[16];
        CU(cudaMemcpyAsync(t2, (const u8 *)din + len, taglen, cudaMemcpyDefault, st));   /* before it can be overwritten */
        LAUNCH(uaes_launch_ocb(&enc, &dec, 0, nonce, daad, aadlen, din, dout, len, dtag, (unsigned)taglen, work, st));
        if (!direct && len) CU(cudaMemcpyAsync(out, c->big, len, cudaMemcpyDefault, st));
        CU(cudaMemcpyAsync(t1, dtag, taglen, cudaMemcpyDefault, st));
        CU(cudaStreamSynchronize(st));
        if (memcmp(t1, t2, taglen)) rc = UAES_AUTH_ERROR;            /* micro_aes.c:1807 */
    }
done:
    scratch_put(c, w, st);
    if (locked) { big_done(c); pthread_mutex_unlock(&c->lock); }
    if (g_burn) { memset(&enc, 0, sizeof enc); memset(&dec, 0, sizeof dec); }
    return rc;
}

int uaes_ocb_encrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return ocb_common(keybits, key, nonce, aad, aadlen, in, len, out, 16, 0);
}

int uaes_ocb_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return ocb_common(keybits, key, nonce, aad, aadlen, in, len, out, 16, 1);
}

/* the reference built with OCB_TAG_LEN < 16 (micro_aes.h:117): the length enters the nonce block
 * (micro_aes.c:1707) and cuts the tag (micro_aes.c:1783, 1807) */
int uaes_ocb_encrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                        const void *in, size_t len, void *out, size_t taglen)
{
    return ocb_common(keybits, key, nonce, aad, aadlen, in, len, out, taglen, 0);
}

int uaes_ocb_decrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                        const void *in, size_t len, void *out, size_t taglen)
{
    return ocb_common(keybits, key, nonce, aad, aadlen, in, len, out, taglen, 1);
}

/* ------------------------------------------------------------------ CCM, batched */

/* One launch for n messages.  Host-side work is bookkeeping only: where the descriptors and the
 * three byte ranges live, staging whatever is host memory through the grow-only device buffers. */
#define BATCH_CCM 0
#define BATCH_EAX 1
#define BATCH_SIV 2
#define BATCH_GCM 3

static int launch_batch(int mode, const uaes_keysched *ks, const uaes_keysched *ks2, int decrypt, unsigned taglen,
                        void *msgs_dev, u64 n, const void *aad, const void *in, void *out, void *stream)
{
    if (mode == BATCH_CCM) return uaes_launch_ccm_batch(ks, decrypt, taglen, msgs_dev, n, aad, in, out, stream);
    return uaes_launch_mac_batch(mode, ks, ks2, decrypt, taglen, msgs_dev, n, aad, in, out, stream);
}

static int mac_batch(int mode, int keybits, const u8 *key, uaes_msg *msgs, size_t n,
                     const void *aad, const void *in, void *out, int decrypt, size_t taglen)
{
    devctx *c;
    uaes_keysched ks, ks2;
    int rc = 0, msgs_dev, locked = 0;
    size_t i, in_ext = 0, out_ext = 0, aad_ext = 0, off;
    const void *din, *daad;
    void *dout, *dmsgs;
    scratch *w = NULL;
    cudaStream_t st = NULL;
    struct cudaPointerAttributes at;

    if (expand_key(keybits, key, &ks)) return fail(UAES_E_BAD_ARGUMENT, "key size must be 128, 192 or 256", 0);
    ks2 = ks;
    if (mode == BATCH_SIV) expand_key(keybits, key + keybits / 8, &ks2);     /* keys = K1 || K2, micro_aes.c:1378 */
    if (n == 0) return 0;
    if ((rc = get_ctx(&c)) != 0) return rc;
    msgs_dev = cudaPointerGetAttributes(&at, msgs) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    if (msgs_dev) {
        if (!is_direct(out) || !is_direct(in) || (aad && !is_direct(aad))) {
            rc = fail(UAES_E_BAD_ARGUMENT, "device descriptors need device (16-byte aligned) aad/in/out", 0);
            goto done;
        }
        st = (cudaStream_t)tls_stream;
        LAUNCH(launch_batch(mode, &ks, &ks2, decrypt, (unsigned)taglen, msgs, n, aad, in, out, st));
        if (!tls_async) CU(cudaStreamSynchronize(st));
        goto done;                                   /* per-message results stay on the device */
    }
    for (i = 0; i < n; ++i) {
        const size_t tag_in = decrypt ? taglen : 0, tag_out = decrypt ? 0 : taglen;
        if (msgs[i].in_off + msgs[i].len + tag_in > in_ext) in_ext = msgs[i].in_off + msgs[i].len + tag_in;
        if (msgs[i].out_off + msgs[i].len + tag_out > out_ext) out_ext = msgs[i].out_off + msgs[i].len + tag_out;
        if (msgs[i].aad_len && msgs[i].aad_off + msgs[i].aad_len > aad_ext) aad_ext = msgs[i].aad_off + msgs[i].aad_len;
    }
    pthread_mutex_lock(&c->lock); locked = 1;                 /* the staged path owns big and st[0] */
    st = c->st[0];
    CU(cudaStreamSynchronize((cudaStream_t)tls_stream));
    /* work area: descriptors, then (if host) the associated data; big: input and output ranges */
    off = (n * sizeof(uaes_msg) + 255) & ~(size_t)255;
    if ((rc = scratch_get(c, off + aad_ext + 64, st, &w)) != 0) goto done;
    dmsgs = w->p;
    CU(cudaMemcpyAsync(dmsgs, msgs, n * sizeof(uaes_msg), cudaMemcpyDefault, st));
    daad = aad;
    if (aad_ext && !is_direct(aad)) {
        CU(cudaMemcpyAsync((u8 *)w->p + off, aad, aad_ext, cudaMemcpyDefault, st));
        daad = (u8 *)w->p + off;
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
    LAUNCH(launch_batch(mode, &ks, &ks2, decrypt, (unsigned)taglen, dmsgs, n, daad, din, dout, st));
    if (dout != out && out_ext) CU(cudaMemcpyAsync(out, dout, out_ext, cudaMemcpyDefault, st));
    CU(cudaMemcpyAsync(msgs, dmsgs, n * sizeof(uaes_msg), cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
    for (i = 0; i < n; ++i) if (msgs[i].result) rc = UAES_AUTH_ERROR;
done:
    scratch_put(c, w, st);
    if (locked) { big_done(c); pthread_mutex_unlock(&c->lock); }
    if (g_burn) { memset(&ks, 0, sizeof ks); memset(&ks2, 0, sizeof ks2); }
    return rc;
}

#define BATCH_ENTRY(name, mode, dec) \
    int name(int keybits, const uaes_u8 *key, uaes_msg *msgs, size_t n, const void *aad, const void *in, void *out) \
    { return mac_batch(mode, keybits, key, msgs, n, aad, in, out, dec, 16); }
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
                      size_t aadlen, const void *in, size_t len, void *out, int decrypt, size_t taglen)
{
    uaes_msg m;
    if (taglen < 1 || taglen > 16 || (mode == BATCH_CCM && (taglen < 4 || taglen % 2)))
        return fail(UAES_E_BAD_ARGUMENT, "tag length must be 1..16 bytes (CCM: even, 4..16)", 0);
    if (len > 0xFFFFFFFFu - 16 || aadlen > 0xFFFFFFFFu) return fail(UAES_E_BAD_ARGUMENT, "message longer than 4 GiB", 0);
    memset(&m, 0, sizeof m);
    m.len = (unsigned int)len; m.aad_len = (unsigned int)aadlen;
    if (noncelen) memcpy(m.nonce, nonce, noncelen);
    return mac_batch(mode, keybits, key, &m, 1, aad, in, out, decrypt, taglen);
}

int uaes_ccm_encrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return mac_single(BATCH_CCM, keybits, key, nonce, 11, aad, aadlen, in, len, out, 0, 16);
}

int uaes_ccm_encrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                        const void *in, size_t len, void *out, size_t taglen)
{
    return mac_single(BATCH_CCM, keybits, key, nonce, 11, aad, aadlen, in, len, out, 0, taglen);
}

int uaes_ccm_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return mac_single(BATCH_CCM, keybits, key, nonce, 11, aad, aadlen, in, len, out, 1, 16);
}

int uaes_ccm_decrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                        const void *in, size_t len, void *out, size_t taglen)
{
    return mac_single(BATCH_CCM, keybits, key, nonce, 11, aad, aadlen, in, len, out, 1, taglen);
}

int uaes_eax_encrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return mac_single(BATCH_EAX, keybits, key, nonce, 16, aad, aadlen, in, len, out, 0, 16);
}

int uaes_eax_encrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                        const void *in, size_t len, void *out, size_t taglen)
{
    return mac_single(BATCH_EAX, keybits, key, nonce, 16, aad, aadlen, in, len, out, 0, taglen);
}

int uaes_eax_decrypt(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                     const void *in, size_t len, void *out)
{
    return mac_single(BATCH_EAX, keybits, key, nonce, 16, aad, aadlen, in, len, out, 1, 16);
}

int uaes_eax_decrypt_ex(int keybits, const uaes_u8 *key, const uaes_u8 *nonce, const void *aad, size_t aadlen,
                        const void *in, size_t len, void *out, size_t taglen)
{
    return mac_single(BATCH_EAX, keybits, key, nonce, 16, aad, aadlen, in, len, out, 1, taglen);
}

/* The reference keeps the synthetic IV and the ciphertext in separate buffers (micro_aes.c:1372,
 * 1394); the batch layout is IV || ciphertext.  A host temporary bridges the two. */
int uaes_siv_encrypt(int keybits, const uaes_u8 *keys, const void *aad, size_t aadlen,
                     const void *in, size_t len, uaes_u8 *iv, void *out)
{
    int rc;
    u8 *tmp = (u8 *)malloc(len + 16);
    if (!tmp) return fail(UAES_E_NO_MEMORY, "malloc(SIV temporary)", 0);
    rc = mac_single(BATCH_SIV, keybits, keys, NULL, 0, aad, aadlen, in, len, tmp, 0, 16);
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
    rc = mac_single(BATCH_SIV, keybits, keys, NULL, 0, aad, aadlen, tmp, len, out, 1, 16);
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
    scratch *sc = NULL;
    cudaStream_t st = (cudaStream_t)tls_stream;
    const u64 end = s->end_block[s->npending - 1];

    expand_key(s->keybits, s->key, &ks);
    for (i = 0; i < s->npending; ++i) after[i] = end - s->end_block[i];
    if ((rc = get_ctx(&c)) != 0) return rc;
    if ((rc = scratch_get(c, GCM_WORK_HEAD + 32 * 24 + 64, st, &sc)) != 0) return rc;
    w = (u8 *)sc->p;
    CU(cudaMemcpyAsync(w + 64, s->partial, (size_t)s->npending * 16, cudaMemcpyDefault, st));
    CU(cudaMemcpyAsync(w + 64 + 32 * 16, after, (size_t)s->npending * 8, cudaMemcpyDefault, st));
    LAUNCH(uaes_launch_gcm_fold(&ks, w + 64, w + 64 + 32 * 16, (unsigned)s->npending, w, st));
    CU(cudaMemcpyAsync(s->partial[0], w, 16, cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
    s->end_block[0] = end;
    s->npending = 1;
done:
    scratch_put(c, sc, st);
    if (g_burn) memset(&ks, 0, sizeof ks);
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
    if ((rc = get_ctx(&c)) != 0) return rc;
    LAUNCH(uaes_launch_fill(seed, first_word, dst, nwords, tls_stream));
    rc = finish_direct();
done:
    return rc;
}

int uaes_xor_fold64(const void *src, size_t nwords, uaes_u64 *result)
{
    devctx *c;
    int rc;
    scratch *w = NULL;
    cudaStream_t st = (cudaStream_t)tls_stream;
    if ((rc = get_ctx(&c)) != 0) return rc;
    if ((rc = scratch_get(c, 64, st, &w)) != 0) return rc;
    LAUNCH(uaes_launch_xor_fold(src, nwords, w->p, st));
    CU(cudaMemcpyAsync(result, w->p, 8, cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
done:
    scratch_put(c, w, st);
    return rc;
}

/* ------------------------------------------------------------------ library state, lifecycle */

void uaes_shutdown(void);

int uaes_set_devices(int n)
{
    const int have = uaes_device_count();
    if (!g_cfg_ready) cfg_init();
    g_fan_devices = (n <= 0 || (n > have && !g_fan_oversubscribe)) ? have : n;
    if (g_fan_devices > MAX_DEV) g_fan_devices = MAX_DEV;
    if (g_fan_devices < 1) g_fan_devices = 1;
    return g_fan_devices;
}

int uaes_get_devices(void)
{
    if (!g_cfg_ready) cfg_init();
    return g_fan_devices;
}

void uaes_set_fanout_min(size_t bytes_per_device)
{
    if (!g_cfg_ready) cfg_init();
    g_fan_min = bytes_per_device ? bytes_per_device : 1;
}

/* staging geometry: chunk bytes (rounded down to 64 KiB, 64 KiB .. 256 MiB) and number of slots
 * (1 .. 8); 0 keeps a value.  Releases the current chunks first, so no call may be in flight. */
void uaes_set_staging(size_t chunk_bytes, int slots)
{
    if (!g_cfg_ready) cfg_init();
    uaes_shutdown();
    if (chunk_bytes) {
        chunk_bytes &= ~(size_t)65535;
        if (chunk_bytes < 65536) chunk_bytes = 65536;
        if (chunk_bytes > MAX_CHUNK) chunk_bytes = MAX_CHUNK;
        g_chunk = chunk_bytes;
    }
    if (slots >= 1 && slots <= MAX_SLOT) g_nslot = slots;
}

void uaes_set_burn(int enable)
{
    if (!g_cfg_ready) cfg_init();
    g_burn = enable != 0;
}

void uaes_set_copy_threads(int n)
{
    if (!g_cfg_ready) cfg_init();
    if (n >= 1 && n <= 32) g_copy_threads = n;
}

int uaes_host_register(void *p, size_t bytes)
{
    if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess)
        return fail(UAES_E_CUDA, "cudaHostRegister", (int)cudaGetLastError());
    return 0;
}

int uaes_host_unregister(void *p)
{
    if (cudaHostUnregister(p) != cudaSuccess) return fail(UAES_E_CUDA, "cudaHostUnregister", (int)cudaGetLastError());
    return 0;
}

/* frees what the library holds on one device; all = 0 keeps the streams and the staging chunks */
static void dev_release(devctx *c, int all)
{
    int i, cur = 0;
    if (!c->ready) return;
    cudaGetDevice(&cur);
    if (cudaSetDevice(c->dev) != cudaSuccess) { cudaGetLastError(); return; }
    pthread_mutex_lock(&c->lock);
    cudaDeviceSynchronize();
    if (c->big) { if (g_burn) cudaMemset(c->big, 0, c->big_bytes); cudaFree(c->big); c->big = NULL; c->big_bytes = 0; }
    pthread_mutex_lock(&c->plock);
    for (i = 0; i < MAX_SCRATCH; ++i) {
        scratch *s = &c->pool[i];
        if (s->busy) continue;
        if (s->p) { if (g_burn) cudaMemset(s->p, 0, s->bytes); cudaFree(s->p); s->p = NULL; s->bytes = 0; s->used = 0; }
        if (all && s->have_ev) { cudaEventDestroy(s->ev); s->have_ev = 0; }
    }
    pthread_mutex_unlock(&c->plock);
    for (i = 0; i < MAX_SLOT; ++i) {
        if (c->slot[i] && (all || g_burn)) cudaMemset(c->slot[i], 0, CHUNK_BYTES);
        if (c->hslot[i] && (all || g_burn)) memset(c->hslot[i], 0, CHUNK_BYTES);
        if (all) {
            if (c->slot[i]) { cudaFree(c->slot[i]); c->slot[i] = NULL; }
            if (c->hslot[i]) { cudaFreeHost(c->hslot[i]); c->hslot[i] = NULL; }
            cudaStreamDestroy(c->st[i]);
            cudaEventDestroy(c->ev[i]);
        }
    }
    if (all)
        for (i = 0; i < 16; ++i)
            if (c->rev_ok[i]) { cudaEventDestroy(c->rev[i]); c->rev_ok[i] = 0; }
    if (all) c->ready = 0;
    pthread_mutex_unlock(&c->lock);
    cudaSetDevice(cur);
    cudaGetLastError();
}

/* gives back the grow-on-demand device memory (full-size staging, work areas) of every device;
 * staging chunks and streams stay for the next call */
void uaes_trim(void)
{
    int i;
    pthread_once(&g_dev_once, dev_locks_init);
    for (i = 0; i < MAX_DEV; ++i) dev_release(&g_dev[i], 0);
}

/* releases everything, wiping staging memory first: the library is back in its initial state and
 * initialises again on the next call.  No call may be in flight. */
void uaes_shutdown(void)
{
    int i;
    pthread_once(&g_dev_once, dev_locks_init);
    for (i = 0; i < MAX_DEV; ++i) dev_release(&g_dev[i], 1);
}
