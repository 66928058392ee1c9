// cornetto_b200/csrc/context.cu -- context, host/device batches, scratch, timing.
#include <pthread.h>
#include <stdarg.h>

#include <thread>

#include "corn_internal.cuh"

// --------------------------------------------------------------------------------------------
// errors
// --------------------------------------------------------------------------------------------
int corn_set_err(corn_ctx *ctx, int code, const char *fmt, ...)
{
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof ctx->err, fmt, ap);
        va_end(ap);
    }
    return code;
}

extern "C" const char *corn_gpu_strerror(int status)
{
    switch (status) {
    case CORN_OK: return "ok";
    case CORN_E_NOGPU: return "no usable CUDA device (this build has no CPU fallback)";
    case CORN_E_CUDA: return "CUDA runtime error";
    case CORN_E_ARG: return "invalid argument";
    case CORN_E_NOMEM: return "out of memory";
    case CORN_E_LAYOUT: return "batch violates the CORN_ALIGN / zero-padding layout";
    case CORN_E_TOOBIG: return "batch or record too large";
    case CORN_E_STATE: return "call sequence error";
    case CORN_E_INTERNAL: return "internal consistency check failed";
    default: return "unknown status";
    }
}

extern "C" const char *corn_gpu_last_error(const corn_ctx_t *ctx) { return ctx ? ctx->err : ""; }

// --------------------------------------------------------------------------------------------
// context
// --------------------------------------------------------------------------------------------
extern "C" int corn_gpu_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return CORN_E_NOGPU; }
    return n;
}

extern "C" int corn_gpu_init(int device, corn_ctx_t **out)
{
    if (!out) return CORN_E_ARG;
    *out = NULL;
    int n = corn_gpu_device_count();
    if (n <= 0) return CORN_E_NOGPU;
    if (device < 0) {
        const char *e = getenv("CORNETTO_GPU");
        device = e ? atoi(e) : 0;
    }
    if (device >= n) return CORN_E_ARG;
    // (attribute queries, not cudaGetDeviceProperties: that one reads every property and costs tens of ms)
    int cc_major = 0, sm_count = 0;
    if (cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return CORN_E_NOGPU;
    if (cc_major < 10) return CORN_E_NOGPU;     // kernels are built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return CORN_E_NOGPU;

    corn_ctx *ctx = (corn_ctx *)calloc(1, sizeof(corn_ctx));
    if (!ctx) return CORN_E_NOMEM;
    ctx->device = device;
    ctx->sm_count = sm_count;
    int fail = CORN_OK;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { free(ctx); return CORN_E_CUDA; }
    ctx->stream = ctx->own_stream;
    for (int i = 0; i < 16 && fail == CORN_OK; ++i)
        if (cudaEventCreate(&ctx->ev[i]) != cudaSuccess) { ctx->ev[i] = NULL; fail = CORN_E_CUDA; }
    if (fail == CORN_OK && cudaMallocHost(&ctx->h_pinned_small, 4096) != cudaSuccess) fail = CORN_E_NOMEM;
    if (fail != CORN_OK) {                               // give back what was created so far
        for (int i = 0; i < 16; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
        cudaStreamDestroy(ctx->own_stream);
        cudaGetLastError();
        free(ctx);
        return fail;
    }
    *out = ctx;
    return CORN_OK;
}

static void dbuf_free(corn_dbuf *b) { if (b->p) cudaFree(b->p); b->p = NULL; b->cap = 0; }

extern "C" void corn_gpu_destroy(corn_ctx_t *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    corn_ctx_adopt(ctx, NULL);
    if (ctx->spare_base) cudaFree(ctx->spare_base);
    corn_dbuf *bufs[] = { &ctx->cand, &ctx->tile_tab, &ctx->events, &ctx->runs, &ctx->misc, &ctx->scan_tmp,
                          &ctx->bins, &ctx->bitmap, &ctx->wins, &ctx->sd_slots, &ctx->sd_out, &ctx->sd_tab,
                          &ctx->ing_text, &ctx->ing_tab, &ctx->ing_lines, &ctx->ing_rec, &ctx->ranks, &ctx->hot };
    for (size_t i = 0; i < sizeof bufs / sizeof bufs[0]; ++i) dbuf_free(bufs[i]);
    for (int i = 0; i < 16; ++i) cudaEventDestroy(ctx->ev[i]);
    if (ctx->aux_stream) { cudaStreamDestroy(ctx->aux_stream); cudaEventDestroy(ctx->aux_ev[0]); cudaEventDestroy(ctx->aux_ev[1]); }
    cudaStreamDestroy(ctx->own_stream);
    cudaFreeHost(ctx->h_pinned_small);
    if (ctx->stage) {
        cudaFreeHost(ctx->stage);
        for (int i = 0; i < ctx->stage_threads; ++i) cudaStreamDestroy(ctx->stage_stream[i]);
        for (int i = 0; i < ctx->stage_threads * CORN_STAGE_SLOTS; ++i) cudaEventDestroy(ctx->stage_ev[i]);
    }
    free(ctx);
}

extern "C" int corn_gpu_set_stream(corn_ctx_t *ctx, void *s)
{
    if (!ctx) return CORN_E_ARG;
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return CORN_OK;
}

// --------------------------------------------------------------------------------------------
// host -> device copies
// --------------------------------------------------------------------------------------------
static void stage_worker(corn_ctx *ctx, int t, uint8_t *d_dst, const uint8_t *h_src, size_t bytes, cudaError_t *err)
{
    cudaSetDevice(ctx->device);
    const size_t n_piece = (bytes + CORN_STAGE_BYTES - 1) / CORN_STAGE_BYTES;
    int use = 0;
    for (size_t p = (size_t)t; p < n_piece; p += (size_t)ctx->stage_threads, ++use) {
        const int slot = t * CORN_STAGE_SLOTS + (use % CORN_STAGE_SLOTS);
        uint8_t *buf = ctx->stage + (size_t)slot * CORN_STAGE_BYTES;
        const size_t off = p * CORN_STAGE_BYTES, len = bytes - off < CORN_STAGE_BYTES ? bytes - off : CORN_STAGE_BYTES;
        // slot drained?  (also covers a copy left in flight by a previous call; an event never recorded is "complete")
        cudaError_t e = cudaEventSynchronize(ctx->stage_ev[slot]);
        if (e == cudaSuccess) {
            memcpy(buf, h_src + off, len);
            e = cudaMemcpyAsync(d_dst + off, buf, len, cudaMemcpyHostToDevice, ctx->stage_stream[t]);
        }
        if (e == cudaSuccess) e = cudaEventRecord(ctx->stage_ev[slot], ctx->stage_stream[t]);
        if (e != cudaSuccess) { *err = e; return; }
    }
}

int corn_h2d(corn_ctx *ctx, void *d_dst, const void *h_src, size_t bytes)
{
    if (!bytes) return CORN_OK;
    cudaPointerAttributes at;
    bool pinned = false;
    if (cudaPointerGetAttributes(&at, h_src) == cudaSuccess) pinned = at.type == cudaMemoryTypeHost;
    else cudaGetLastError();
    if (pinned || bytes < 4 * (size_t)CORN_STAGE_BYTES) {
        CORN_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return CORN_OK;
    }
    if (!ctx->stage) {
        int nt = 4;      // measured on the 32-vCPU VM, 3 GB: 4 threads 116-132 ms, 8: 144-158, 12: 144-192, 16: 177-212 (memory-bound)
        if (const char *e = getenv("CORNETTO_STAGE_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= CORN_STAGE_THREADS) nt = v; }
        CORN_CUDA(ctx, cudaMallocHost((void **)&ctx->stage, (size_t)nt * CORN_STAGE_SLOTS * CORN_STAGE_BYTES));
        ctx->stage_threads = nt;
        for (int i = 0; i < nt; ++i) CORN_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stage_stream[i], cudaStreamNonBlocking));
        for (int i = 0; i < nt * CORN_STAGE_SLOTS; ++i) CORN_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
    }
    const int NT = ctx->stage_threads;
    // the copy streams start after whatever ctx->stream has queued so far (e.g. a kernel still reading d_dst)
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[15], ctx->stream));
    for (int i = 0; i < NT; ++i) CORN_CUDA(ctx, cudaStreamWaitEvent(ctx->stage_stream[i], ctx->ev[15], 0));
    cudaError_t err[CORN_STAGE_THREADS];
    std::thread th[CORN_STAGE_THREADS - 1];
    for (int t = 0; t < NT; ++t) err[t] = cudaSuccess;
    for (int t = 1; t < NT; ++t)
        th[t - 1] = std::thread(stage_worker, ctx, t, (uint8_t *)d_dst, (const uint8_t *)h_src, bytes, &err[t]);
    stage_worker(ctx, 0, (uint8_t *)d_dst, (const uint8_t *)h_src, bytes, &err[0]);
    for (int t = 1; t < NT; ++t) th[t - 1].join();
    for (int t = 0; t < NT; ++t)
        if (err[t] != cudaSuccess) return corn_set_err(ctx, CORN_E_CUDA, "staged host->device copy: %s", cudaGetErrorString(err[t]));
    // ... and ctx->stream continues after the last piece of every copy stream
    for (int t = 0; t < NT; ++t) {
        CORN_CUDA(ctx, cudaEventRecord(ctx->ev[15], ctx->stage_stream[t]));
        CORN_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev[15], 0));
    }
    return CORN_OK;
}

int corn_dbuf_reserve(corn_ctx *ctx, corn_dbuf *b, size_t bytes)
{
    if (bytes <= b->cap) return CORN_OK;
    CORN_CUDA(ctx, cudaSetDevice(ctx->device));
    if (b->p) {
        CORN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(b->p);
        b->p = NULL; b->cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;   // a little slack so steady-state calls do not reallocate
    CORN_CUDA(ctx, cudaMalloc(&b->p, want));
    b->cap = want;
    return CORN_OK;
}

int corn_read_small(corn_ctx *ctx, void *h_dst, const void *d_src, size_t bytes)
{
    if (bytes > 4096) return corn_set_err(ctx, CORN_E_INTERNAL, "corn_read_small: %zu bytes", bytes);
    CORN_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned_small, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CORN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(h_dst, ctx->h_pinned_small, bytes);
    return CORN_OK;
}

// Pinned result blocks are recycled through a small process-wide pool: cudaMallocHost /
// cudaFreeHost cost far more than the copies they serve when a caller loops over batches.
namespace {
struct HostBlock { void *p; size_t bytes; };
constexpr int POOL_SLOTS = 8;
HostBlock g_pool[POOL_SLOTS];        // free blocks
HostBlock g_live[64];                // handed out (size lookup on free)
pthread_mutex_t g_pool_mu = PTHREAD_MUTEX_INITIALIZER;
}

void *corn_host_alloc(size_t bytes)
{
    if (bytes == 0) bytes = 1;
    void *p = NULL;
    size_t got = 0;
    pthread_mutex_lock(&g_pool_mu);
    int best = -1;
    for (int i = 0; i < POOL_SLOTS; ++i)
        if (g_pool[i].p && g_pool[i].bytes >= bytes && (best < 0 || g_pool[i].bytes < g_pool[best].bytes)) best = i;
    if (best >= 0) { p = g_pool[best].p; got = g_pool[best].bytes; g_pool[best].p = NULL; }
    pthread_mutex_unlock(&g_pool_mu);
    if (!p) {
        got = bytes + bytes / 4;
        if (cudaMallocHost(&p, got) != cudaSuccess) { cudaGetLastError(); return NULL; }
    }
    pthread_mutex_lock(&g_pool_mu);
    for (int i = 0; i < 64; ++i) if (!g_live[i].p) { g_live[i].p = p; g_live[i].bytes = got; break; }
    pthread_mutex_unlock(&g_pool_mu);
    return p;
}

void corn_host_free(void *p)
{
    if (!p) return;
    size_t bytes = 0;
    pthread_mutex_lock(&g_pool_mu);
    for (int i = 0; i < 64; ++i) if (g_live[i].p == p) { bytes = g_live[i].bytes; g_live[i].p = NULL; break; }
    if (bytes) {
        int slot = -1;
        for (int i = 0; i < POOL_SLOTS; ++i) if (!g_pool[i].p) { slot = i; break; }
        if (slot < 0) {                      // pool full: evict the smallest block if this one is bigger
            int small = 0;
            for (int i = 1; i < POOL_SLOTS; ++i) if (g_pool[i].bytes < g_pool[small].bytes) small = i;
            if (g_pool[small].bytes < bytes) { void *old = g_pool[small].p; g_pool[small].p = p; g_pool[small].bytes = bytes; p = old; }
        } else { g_pool[slot].p = p; g_pool[slot].bytes = bytes; p = NULL; }
    }
    pthread_mutex_unlock(&g_pool_mu);
    if (p) cudaFreeHost(p);
}

extern "C" int corn_gpu_last_timing(const corn_ctx_t *ctx_c, corn_timing_t *t)
{
    if (!ctx_c || !t) return CORN_E_ARG;
    corn_ctx *ctx = (corn_ctx *)ctx_c;
    int r = CORN_OK;
    if (ctx->pending) {                       // an un-synced telofind_dev(out == NULL): finish it first
        cudaSetDevice(ctx->device);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return CORN_E_CUDA;
        r = corn_telofind_resolve(ctx);
    }
    *t = ctx->timing;
    return r;
}
extern "C" uint64_t corn_gpu_total_launches(const corn_ctx_t *ctx) { return ctx ? ctx->total_launches : 0; }

// --------------------------------------------------------------------------------------------
// host batch builder (pinned)
// --------------------------------------------------------------------------------------------
struct corn_hbatch {
    uint8_t  *seq;
    uint64_t  cap, used;       // used is always a multiple of CORN_ALIGN
    uint64_t *offset;
    uint32_t *length;
    uint32_t  n_rec, max_rec;
    int       pinned;
};

static inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

extern "C" int corn_hbatch_create(uint64_t capacity_bytes, uint32_t max_records, corn_hbatch_t **out)
{
    if (!out || max_records == 0) return CORN_E_ARG;
    capacity_bytes = align_up(capacity_bytes < CORN_ALIGN ? CORN_ALIGN : capacity_bytes, CORN_ALIGN);
    if (capacity_bytes > CORN_MAX_BATCH_BYTES) return CORN_E_TOOBIG;
    corn_hbatch *hb = (corn_hbatch *)calloc(1, sizeof *hb);
    if (!hb) return CORN_E_NOMEM;
    // plain page-aligned memory: no CUDA call here (driver start-up takes seconds and must not
    // delay the parser); corn_hbatch_pin() page-locks it later, on the thread that owns the GPU
    void *p = NULL;
    if (posix_memalign(&p, 4096, capacity_bytes) != 0) p = NULL;
    hb->seq = (uint8_t *)p;
    hb->offset = (uint64_t *)malloc(sizeof(uint64_t) * max_records);
    hb->length = (uint32_t *)malloc(sizeof(uint32_t) * max_records);
    if (!hb->seq || !hb->offset || !hb->length) { corn_hbatch_destroy(hb); return CORN_E_NOMEM; }
    hb->cap = capacity_bytes;
    hb->max_rec = max_records;
    *out = hb;
    return CORN_OK;
}

extern "C" void corn_hbatch_destroy(corn_hbatch_t *hb)
{
    if (!hb) return;
    if (hb->seq) { if (hb->pinned) cudaHostUnregister(hb->seq); free(hb->seq); }
    free(hb->offset); free(hb->length); free(hb);
}

extern "C" int corn_hbatch_pin(corn_hbatch_t *hb)
{
    if (!hb) return CORN_E_ARG;
    if (hb->pinned) return CORN_OK;
    cudaError_t e = cudaHostRegister(hb->seq, hb->cap, cudaHostRegisterDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorMemoryAllocation ? CORN_E_NOMEM : CORN_E_CUDA; }
    hb->pinned = 1;
    return CORN_OK;
}

extern "C" void corn_hbatch_reset(corn_hbatch_t *hb) { hb->used = 0; hb->n_rec = 0; }

extern "C" uint64_t corn_hbatch_room(const corn_hbatch_t *hb)
{
    if (hb->n_rec >= hb->max_rec || hb->used + CORN_ALIGN > hb->cap) return 0;
    // a record of length L occupies align_up(L + 1, CORN_ALIGN) bytes
    uint64_t left = hb->cap - hb->used;
    uint64_t room = left - 1;
    if (room > 0x7FFFFFFFull) room = 0x7FFFFFFFull;   // kseq records are < 2^31 (src/kseq.h:84,185)
    return room;
}

extern "C" uint8_t *corn_hbatch_cursor(corn_hbatch_t *hb) { return hb->seq + hb->used; }

extern "C" int corn_hbatch_commit(corn_hbatch_t *hb, uint64_t length)
{
    if (hb->n_rec >= hb->max_rec) return CORN_E_TOOBIG;
    if (length > 0x7FFFFFFFull) return CORN_E_TOOBIG;
    uint64_t span = align_up(length + 1, CORN_ALIGN);
    if (hb->used + span > hb->cap) return CORN_E_TOOBIG;
    memset(hb->seq + hb->used + length, 0, span - length);
    hb->offset[hb->n_rec] = hb->used;
    hb->length[hb->n_rec] = (uint32_t)length;
    hb->n_rec++;
    hb->used += span;
    return CORN_OK;
}

extern "C" int corn_hbatch_add(corn_hbatch_t *hb, const void *bases, uint64_t length)
{
    if (length > corn_hbatch_room(hb) || (length == 0 && corn_hbatch_room(hb) == 0)) return CORN_E_TOOBIG;
    if (length) memcpy(hb->seq + hb->used, bases, length);
    return corn_hbatch_commit(hb, length);
}

extern "C" void corn_hbatch_view(const corn_hbatch_t *hb, corn_batch_t *v)
{
    v->seq = hb->seq; v->offset = hb->offset; v->length = hb->length;
    v->n_rec = hb->n_rec; v->total_bytes = hb->used;
}

// --------------------------------------------------------------------------------------------
// device batches
// --------------------------------------------------------------------------------------------
static int dbatch_new(corn_ctx *ctx, const uint64_t *offset, const uint32_t *length, uint32_t n_rec,
                      uint64_t total_bytes, corn_dbatch **out)
{
    if (total_bytes % CORN_ALIGN) return corn_set_err(ctx, CORN_E_LAYOUT, "total_bytes %% %u != 0", CORN_ALIGN);
    if (total_bytes > CORN_MAX_BATCH_BYTES) return corn_set_err(ctx, CORN_E_TOOBIG, "batch of %llu bytes", (unsigned long long)total_bytes);
    corn_dbatch *db = (corn_dbatch *)calloc(1, sizeof *db);
    if (!db) return CORN_E_NOMEM;
    db->n_rec = n_rec;
    db->total_bytes = total_bytes;
    db->h_rec_off = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)n_rec + 1));
    db->h_rec_len = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)n_rec + 1));
    db->h_bin_base = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)n_rec + 1));
    if (!db->h_rec_off || !db->h_rec_len || !db->h_bin_base) { free(db->h_rec_off); free(db->h_rec_len); free(db->h_bin_base); free(db); return CORN_E_NOMEM; }
    uint64_t prev_end = 0;
    for (uint32_t i = 0; i < n_rec; ++i) {
        uint64_t o = offset[i], l = length[i];
        int bad = (o % CORN_ALIGN) || o < prev_end || o + l + 1 > total_bytes || l > 0x7FFFFFFFull;
        if (bad) {
            free(db->h_rec_off); free(db->h_rec_len); free(db->h_bin_base); free(db);
            return corn_set_err(ctx, CORN_E_LAYOUT, "record %u: offset %llu length %llu", i, (unsigned long long)o, (unsigned long long)l);
        }
        prev_end = o + l + 1;
        db->h_rec_off[i] = (uint32_t)o;
        db->h_rec_len[i] = (uint32_t)l;
        db->h_bin_base[i] = (uint32_t)(db->n_bins_total > 0xFFFFFF00ull ? 0xFFFFFF00ull : db->n_bins_total);
        db->n_bins_total += corn_nbins_of((uint32_t)l);
        db->n_bases += l;
    }
    db->h_rec_off[n_rec] = (uint32_t)total_bytes;
    db->h_rec_len[n_rec] = 0;
    db->h_bin_base[n_rec] = (uint32_t)(db->n_bins_total > 0xFFFFFF00ull ? 0xFFFFFF00ull : db->n_bins_total);

    cudaError_t e;
    // round the data area up to whole tiles so the scan kernel never needs a bounds check
    uint64_t tiles = (total_bytes + CORN_TILE_BYTES - 1) / CORN_TILE_BYTES;
    db->alloc_bytes = CORN_GUARD_BYTES + tiles * CORN_TILE_BYTES + CORN_TAIL_BYTES;
    db->span_bytes = db->alloc_bytes;
    if (ctx->spare_base && ctx->spare_bytes >= db->alloc_bytes) {
        db->d_base = ctx->spare_base; db->alloc_bytes = ctx->spare_bytes;
        ctx->spare_base = NULL; ctx->spare_bytes = 0;
        e = cudaSuccess;
    } else {
        e = cudaMalloc((void **)&db->d_base, db->alloc_bytes);
    }
    if (e == cudaSuccess) e = cudaMalloc((void **)&db->d_rec_off, sizeof(uint32_t) * ((size_t)n_rec + 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&db->d_rec_len, sizeof(uint32_t) * ((size_t)n_rec + 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&db->d_bin_base, sizeof(uint32_t) * ((size_t)n_rec + 1));
    if (e != cudaSuccess) {
        cudaGetLastError();
        const unsigned long long want = db->alloc_bytes;
        if (db->d_base) cudaFree(db->d_base);
        if (db->d_rec_off) cudaFree(db->d_rec_off);
        if (db->d_rec_len) cudaFree(db->d_rec_len);
        if (db->d_bin_base) cudaFree(db->d_bin_base);
        free(db->h_rec_off); free(db->h_rec_len); free(db->h_bin_base); free(db);
        return corn_set_err(ctx, CORN_E_NOMEM, "cudaMalloc of %llu bytes: %s", want, cudaGetErrorString(e));
    }
    db->d_seq = db->d_base + CORN_GUARD_BYTES;
    *out = db;
    return CORN_OK;
}

extern "C" void corn_gpu_dbatch_free(corn_ctx_t *ctx, corn_dbatch_t *db)
{
    if (!db) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->last_db == db) { ctx->last_db = NULL; ctx->pending = 0; }
        if (db->alloc_bytes > ctx->spare_bytes) {            // keep the larger buffer for the next upload
            uint8_t *old = ctx->spare_base;
            ctx->spare_base = db->d_base; ctx->spare_bytes = db->alloc_bytes;
            db->d_base = old;
        }
    }
    if (db->d_base) cudaFree(db->d_base);
    cudaFree(db->d_rec_off); cudaFree(db->d_rec_len); cudaFree(db->d_bin_base);
    free(db->h_rec_off); free(db->h_rec_len); free(db->h_bin_base); free(db);
}

void corn_ctx_adopt(corn_ctx *ctx, corn_dbatch *db)
{
    if (ctx->owned_db && ctx->owned_db != db) corn_gpu_dbatch_free(ctx, ctx->owned_db);
    ctx->owned_db = db;
}

extern "C" void *corn_gpu_dbatch_seq_ptr(const corn_dbatch_t *db) { return db ? db->d_seq : NULL; }
extern "C" uint64_t corn_gpu_dbatch_bytes(const corn_dbatch_t *db) { return db ? db->total_bytes : 0; }

static int dbatch_tables_to_device(corn_ctx *ctx, corn_dbatch *db)
{
    CORN_CUDA(ctx, cudaMemcpyAsync(db->d_rec_off, db->h_rec_off, sizeof(uint32_t) * ((size_t)db->n_rec + 1), cudaMemcpyHostToDevice, ctx->stream));
    CORN_CUDA(ctx, cudaMemcpyAsync(db->d_rec_len, db->h_rec_len, sizeof(uint32_t) * ((size_t)db->n_rec + 1), cudaMemcpyHostToDevice, ctx->stream));
    CORN_CUDA(ctx, cudaMemcpyAsync(db->d_bin_base, db->h_bin_base, sizeof(uint32_t) * ((size_t)db->n_rec + 1), cudaMemcpyHostToDevice, ctx->stream));
    return CORN_OK;
}

extern "C" int corn_gpu_upload(corn_ctx_t *ctx, const corn_batch_t *b, corn_dbatch_t **out)
{
    if (!ctx || !b || !out || (b->n_rec && (!b->offset || !b->length)) || (b->total_bytes && !b->seq)) return CORN_E_ARG;
    CORN_CUDA(ctx, cudaSetDevice(ctx->device));
    corn_dbatch *db = NULL;
    CORN_TRY(dbatch_new(ctx, b->offset, b->length, b->n_rec, b->total_bytes, &db));
    int r = CORN_OK;
    cudaError_t e = cudaEventRecord(ctx->ev[0], ctx->stream);
    // guard + tail (and the round-up-to-tile area) are zero; the record area is copied as is:
    // the layout contract makes the caller responsible for the zero padding between records.
    if (e == cudaSuccess) e = cudaMemsetAsync(db->d_base, 0, CORN_GUARD_BYTES, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(db->d_seq + b->total_bytes, 0, db->span_bytes - CORN_GUARD_BYTES - b->total_bytes, ctx->stream);
    if (e == cudaSuccess && b->total_bytes) r = corn_h2d(ctx, db->d_seq, b->seq, b->total_bytes);
    if (e == cudaSuccess && r == CORN_OK) r = dbatch_tables_to_device(ctx, db);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[1], ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess || r != CORN_OK) {
        corn_gpu_dbatch_free(ctx, db);
        return r != CORN_OK ? r : corn_set_err(ctx, CORN_E_CUDA, "upload: %s", cudaGetErrorString(e));
    }
    memset(&ctx->timing, 0, sizeof ctx->timing);
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[0], ctx->ev[1]);
    *out = db;
    return CORN_OK;
}

extern "C" int corn_gpu_dbatch_alloc(corn_ctx_t *ctx, const uint32_t *length, uint32_t n_rec, corn_dbatch_t **out)
{
    if (!ctx || !out || (n_rec && !length)) return CORN_E_ARG;
    CORN_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t *off = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)n_rec + 1));
    if (!off) return CORN_E_NOMEM;
    uint64_t used = 0;
    for (uint32_t i = 0; i < n_rec; ++i) { off[i] = used; used += align_up((uint64_t)length[i] + 1, CORN_ALIGN); }
    corn_dbatch *db = NULL;
    int r = dbatch_new(ctx, off, length, n_rec, used, &db);
    free(off);
    if (r != CORN_OK) return r;
    cudaError_t e = cudaMemsetAsync(db->d_base, 0, db->span_bytes, ctx->stream);
    if (e == cudaSuccess) r = dbatch_tables_to_device(ctx, db);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess || r != CORN_OK) {
        corn_gpu_dbatch_free(ctx, db);
        return r != CORN_OK ? r : corn_set_err(ctx, CORN_E_CUDA, "dbatch_alloc: %s", cudaGetErrorString(e));
    }
    *out = db;
    return CORN_OK;
}

// ingest.cu: a batch whose record bytes (and padding inside [0, total_bytes)) a kernel is about to
// write; only the guard and the area behind total_bytes are zeroed here.  No host sync.
int corn_dbatch_from_lengths(corn_ctx *ctx, const uint32_t *length, uint32_t n_rec, corn_dbatch **out)
{
    uint64_t *off = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)n_rec + 1));
    if (!off) return CORN_E_NOMEM;
    uint64_t used = 0;
    for (uint32_t i = 0; i < n_rec; ++i) { off[i] = used; used += align_up((uint64_t)length[i] + 1, CORN_ALIGN); }
    corn_dbatch *db = NULL;
    int r = dbatch_new(ctx, off, length, n_rec, used, &db);
    free(off);
    if (r != CORN_OK) return r;
    cudaError_t e = cudaMemsetAsync(db->d_base, 0, CORN_GUARD_BYTES, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(db->d_seq + used, 0, db->span_bytes - CORN_GUARD_BYTES - used, ctx->stream);
    if (e == cudaSuccess) r = dbatch_tables_to_device(ctx, db);
    if (e != cudaSuccess || r != CORN_OK) {
        corn_gpu_dbatch_free(ctx, db);
        return r != CORN_OK ? r : corn_set_err(ctx, CORN_E_CUDA, "dbatch_from_lengths: %s", cudaGetErrorString(e));
    }
    *out = db;
    return CORN_OK;
}

extern "C" int corn_gpu_dbatch_download(corn_ctx_t *ctx, const corn_dbatch_t *db, uint32_t rec, uint8_t *dst)
{
    if (!ctx || !db || !dst || rec >= db->n_rec) return CORN_E_ARG;
    CORN_CUDA(ctx, cudaSetDevice(ctx->device));
    CORN_CUDA(ctx, cudaMemcpyAsync(dst, db->d_seq + db->h_rec_off[rec], db->h_rec_len[rec], cudaMemcpyDeviceToHost, ctx->stream));
    CORN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CORN_OK;
}
