/* cornetto_b200/host/ingest.c -- feeder for plain FASTA/FASTQ files that parses on the GPU.
 *
 * The reference reads every input through kseq (src/kseq.h:184-224), one byte loop on one thread;
 * cornetto_b200/host/fastx.c is that loop, restated, and stays the reader for stdin, gzip and
 * irregular text.  For plain files this feeder skips it: the calling thread only read(2)s the file
 * in large blocks; a worker thread per GPU context hands each block to corn_gpu_ingest(), which
 * ships the raw bytes over PCIe, finds the records on the device and leaves them resident, and then
 * runs the scan (telofind / sdust callback) on that resident batch.  Blocks are cut where the device
 * says the last complete record ended (`consumed`); the unconsumed tail is carried into the next
 * block.  Output is written strictly in block order, so stdout is byte-identical to the sequential
 * reference loop.
 *
 * If a block is outside the regular subset (include/corn_gpu.h: corn_gpu_ingest), or holds a record
 * larger than the block, run_ingest_pipeline() stops, reports the file offset of that block and the
 * caller continues from there with the serial reader. */
#include <ctype.h>
#include <fcntl.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "cornetto.h"

size_t cornetto_name_len(const uint8_t *p, size_t max)
{
    size_t k = 0;
    while (k < max && !isspace(p[k])) ++k;
    return k;
}

typedef struct {
    pthread_t       th;
    pthread_mutex_t mu;
    pthread_cond_t  cv;
    int             state;       /* 0 idle (output, if any, ready), 1 has a block, 2 quit */
    int             ingested;    /* the current block's consumed / irregular are valid */
    int             device;
    corn_ctx_t     *ctx;
    uint8_t        *text;        /* block buffer: [reserve for a carried tail | bytes read from the file] */
    uint8_t        *blk;         /* where the current block starts inside it */
    uint64_t        cap, n;
    int             final, teardown;
    uint64_t        seq;         /* index of the current block in input order */
    uint64_t        consumed;
    int             irregular;
    outbuf_t        out;
    batch_fn        fn;
    void           *arg;
} iworker_t;

static void build_names(rec_batch_t *b, const uint8_t *text, uint64_t n, const corn_ingest_t *ing)
{
    size_t total = 0;
    b->name = (char **)malloc(sizeof(char *) * ((size_t)ing->n_rec + 1));
    CORN_MALLOC_CHK(b->name);
    size_t *len = (size_t *)malloc(sizeof(size_t) * ((size_t)ing->n_rec + 1));
    CORN_MALLOC_CHK(len);
    for (uint32_t r = 0; r < ing->n_rec; ++r) {
        const uint64_t at = ing->hdr_off[r] + 1;
        len[r] = at < n ? cornetto_name_len(text + at, (size_t)(n - at)) : 0;
        total += len[r] + 1;
    }
    b->name_arena = (char *)malloc(total + 1);
    CORN_MALLOC_CHK(b->name_arena);
    char *w = b->name_arena;
    for (uint32_t r = 0; r < ing->n_rec; ++r) {
        memcpy(w, text + ing->hdr_off[r] + 1, len[r]);
        w[len[r]] = 0;
        b->name[r] = w;
        w += len[r] + 1;
    }
    free(len);
}

static int g_trace;     /* $CORNETTO_TRACE=1: phase times on stderr */
static double g_t0;
#define TRACE(...) do { if (g_trace) { fprintf(stderr, "[%7.3f] ", realtime() - g_t0); fprintf(stderr, __VA_ARGS__); } } while (0)

static void *iworker_main(void *p)
{
    iworker_t *w = (iworker_t *)p;
    double t0 = realtime();
    int r = corn_gpu_init(w->device, &w->ctx);
    TRACE("[ingest] worker dev %d: corn_gpu_init %.3f s\n", w->device, realtime() - t0);
    if (r != CORN_OK) {
        CORN_ERROR("cannot initialise the GPU (device %d): %s", w->device, corn_gpu_strerror(r));
        exit(EXIT_FAILURE);
    }
    for (;;) {
        pthread_mutex_lock(&w->mu);
        while (w->state == 0) pthread_cond_wait(&w->cv, &w->mu);
        const int st = w->state;
        pthread_mutex_unlock(&w->mu);
        if (st == 2) break;
        /* (the block buffer stays pageable: the library stages the copy through its own small page-locked
         *  ring, which is faster than page-locking and releasing GBs for a buffer that is used once or twice) */
        t0 = realtime();
        corn_ingest_t ing;
        r = corn_gpu_ingest(w->ctx, w->blk, w->n, w->final, &ing);
        if (g_trace) {
            corn_timing_t tm;
            corn_gpu_last_timing(w->ctx, &tm);
            TRACE("[ingest] corn_gpu_ingest %.3f s (h2d %.1f ms, tables %.1f ms, copy %.1f ms) %u records, irregular %d\n",
                  realtime() - t0, tm.h2d_ms, tm.post_ms, tm.scan_ms, ing.n_rec, ing.irregular);
        }
        if (r != CORN_OK) {
            CORN_ERROR("ingest: %s (%s)", corn_gpu_strerror(r), corn_gpu_last_error(w->ctx));
            exit(EXIT_FAILURE);
        }
        pthread_mutex_lock(&w->mu);
        w->consumed = ing.consumed; w->irregular = ing.irregular; w->ingested = 1;
        pthread_cond_broadcast(&w->cv);
        pthread_mutex_unlock(&w->mu);
        if (!ing.irregular && ing.n_rec) {
            rec_batch_t b;
            memset(&b, 0, sizeof b);
            b.n = ing.n_rec; b.db = ing.db; b.length = ing.length; b.seq = w->seq;
            t0 = realtime();
            build_names(&b, w->blk, w->n, &ing);
            w->fn(w->ctx, &b, &w->out, w->arg);
            TRACE("[ingest] scan + format %.3f s\n", realtime() - t0);
            free(b.name); free(b.name_arena);
            corn_gpu_dbatch_free(w->ctx, ing.db);
        }
        corn_gpu_ingest_free(&ing);
        pthread_mutex_lock(&w->mu);
        w->state = 0;
        pthread_cond_broadcast(&w->cv);
        pthread_mutex_unlock(&w->mu);
    }
    if (w->teardown) corn_gpu_destroy(w->ctx);
    return NULL;
}

static void wait_idle(iworker_t *w)
{
    pthread_mutex_lock(&w->mu);
    while (w->state != 0) pthread_cond_wait(&w->cv, &w->mu);
    pthread_mutex_unlock(&w->mu);
}

/* ---- file reading: the page cache delivers ~3 GB/s to one thread, so large blocks are read by four -- */
typedef struct { int fd; uint8_t *dst; uint64_t off, n, got; pthread_t th; } slice_t;

/* A read that fails, or stops before the size fstat() reported (the file shrank under us), is an ERROR: the
 * caller must not mistake it for the end of the input and print a truncated result with exit status 0. */
static int g_read_errno;     /* first read error seen by any slice (0 = none); -1 = short read before st_size */

static uint64_t pread_full(int fd, uint8_t *dst, uint64_t off, uint64_t n)
{
    uint64_t got = 0;
    while (got < n) {
        const size_t want = n - got > (1u << 30) ? (1u << 30) : (size_t)(n - got);
        const ssize_t k = pread(fd, dst + got, want, (off_t)(off + got));
        if (k < 0) { if (errno == EINTR) continue; if (!g_read_errno) g_read_errno = errno ? errno : EIO; break; }
        if (k == 0) { if (!g_read_errno) g_read_errno = -1; break; }      /* callers never ask beyond st_size */
        got += (uint64_t)k;
    }
    return got;
}

static void die_if_read_failed(const char *path)
{
    if (!g_read_errno) return;
    CORN_ERROR("reading %s failed: %s", path, g_read_errno > 0 ? strerror(g_read_errno) : "the file is shorter than its size at open time");
    exit(EXIT_FAILURE);
}

static void *slice_main(void *p)
{
    slice_t *s = (slice_t *)p;
    s->got = pread_full(s->fd, s->dst, s->off, s->n);
    return NULL;
}

static uint64_t read_block(int fd, uint8_t *dst, uint64_t off, uint64_t n)
{
    enum { PARTS = 4 };
    if (n < (64u << 20)) return pread_full(fd, dst, off, n);
    slice_t s[PARTS];
    const uint64_t per = (n / PARTS + 4095) & ~4095ull;
    int started = 0;
    for (int i = 0; i < PARTS; ++i) {
        const uint64_t a = per * (uint64_t)i;
        if (a >= n) break;
        s[i].fd = fd; s[i].dst = dst + a; s[i].off = off + a; s[i].n = a + per > n ? n - a : per; s[i].got = 0;
        if (i == PARTS - 1 || pthread_create(&s[i].th, NULL, slice_main, &s[i]) != 0) { slice_main(&s[i]); s[i].th = 0; }
        ++started;
    }
    uint64_t got = 0;
    int short_read = 0;
    for (int i = 0; i < started; ++i) {
        if (s[i].th) pthread_join(s[i].th, NULL);
        if (!short_read) got += s[i].got;
        if (s[i].got < s[i].n) short_read = 1;
    }
    return got;
}

int run_ingest_pipeline(const char *path, batch_fn fn, void *arg, uint64_t *resume)
{
    *resume = 0;
    g_trace = getenv("CORNETTO_TRACE") && atoi(getenv("CORNETTO_TRACE")) > 0;
    g_t0 = realtime();
    const char *off = getenv("CORNETTO_INGEST");
    if (off && atoi(off) == 0) return 0;
    if (strcmp(path, "-") == 0) return 0;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return 0;
    struct stat sb;
    unsigned char magic[2] = { 0, 0 };
    if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode) || sb.st_size < 2 || pread(fd, magic, 2, 0) != 2) { close(fd); return 0; }
    /* gzip input: the text comes from gzsrc.c (BGZF members inflated in parallel, other gzip files by one stream that
     * runs beside the GPU work); its length is not known in advance.  $CORNETTO_GZ_INGEST=0: serial reader as before. */
    gzsrc_t *gz = NULL;
    if (magic[0] == 0x1f && magic[1] == 0x8b) {
        const char *eg = getenv("CORNETTO_GZ_INGEST");
        if (!(eg && atoi(eg) == 0)) gz = gzsrc_open(path);
        if (!gz) { close(fd); return 0; }
    } else if (magic[0] != '>' && magic[0] != '@') {
        close(fd);
        return 0;
    }
    const uint64_t size = gz ? (1ull << 62) : (uint64_t)sb.st_size;
    int src_eof = 0;

    int n_gpus = 1;                        /* requested; clipped to the devices present once the driver is up */
    const char *e = getenv("CORNETTO_GPUS");
    if (e && atoi(e) > 0) n_gpus = atoi(e);
    const int n_req = n_gpus;
    int n_workers = n_gpus > 1 ? n_gpus : 2;
    const char *dev0 = getenv("CORNETTO_GPU");
    const int base_dev = (n_gpus == 1 && dev0) ? atoi(dev0) : 0;

    /* block size: the whole file on one GPU, an n-th plus an eighth (room for the carried tail) on n */
    const uint64_t max_block = 0xE0000000ull;
    uint64_t block = n_gpus > 1 ? size / (uint64_t)n_gpus + size / (8ull * (uint64_t)n_gpus) + (1u << 16) : size;
    const char *eb = getenv("CORNETTO_BATCH_BYTES"), *em = getenv("CORNETTO_BATCH_MB");
    if (eb && atoll(eb) > 0) block = (uint64_t)atoll(eb);
    else if (em && atoll(em) > 0) block = (uint64_t)atoll(em) << 20;
    if (gz && !(eb && atoll(eb) > 0) && !(em && atoll(em) > 0)) block = 1ull << 30;      /* text blocks of 1 GiB */
    if (block > size) block = size;
    if (block > max_block) block = max_block;
    if (block < 64) block = 64;

    /* Worker 0 starts the driver at once; the others are created while it works on the first block (asking
     * for the device count first would put a second of CUDA start-up in front of everything). */
    iworker_t *w = (iworker_t *)calloc((size_t)n_workers, sizeof(iworker_t));
    CORN_MALLOC_CHK(w);
    int n_spawned = 0;
#define SPAWN_WORKERS(upto)                                                                              \
    for (; n_spawned < (upto); ++n_spawned) {                                                            \
        iworker_t *nw = &w[n_spawned];                                                                   \
        pthread_mutex_init(&nw->mu, NULL);                                                               \
        pthread_cond_init(&nw->cv, NULL);                                                                \
        nw->device = base_dev + (n_spawned % n_gpus);                                                    \
        nw->fn = fn; nw->arg = arg;                                                                      \
        outbuf_init(&nw->out, NULL);                                                                     \
        if (pthread_create(&nw->th, NULL, iworker_main, nw) != 0) { CORN_ERROR("%s", "pthread_create failed"); exit(EXIT_FAILURE); } \
    }
    SPAWN_WORKERS(1);

    /* A block is cut where the device says its last complete record ends, so block i+1 begins with a tail
     * carried over from block i.  To keep the file read going while block i is being shipped and parsed, the
     * bytes that follow block i in the file are read ahead into the NEXT worker's buffer behind a reserve of
     * `reserve` bytes; once the tail is known it is copied in front of them.  (A tail larger than the reserve
     * -- a record of hundreds of MB cut by a block end -- drops the read-ahead and re-reads.) */
    const uint64_t reserve = block >= size ? 0 : (block / 4 < (256ull << 20) ? block / 4 : (256ull << 20));
    int dispatched = 0, written = 0, complete = 0;
    uint64_t file_pos = 0;                 /* next byte to read from the file */
    const uint8_t *carry = NULL;           /* unconsumed tail of the previous block (in that worker's buffer) */
    uint64_t carry_len = 0;
    int ahead_for = -1;                    /* block index whose fresh bytes are already in its worker's buffer */
    uint64_t ahead_bytes = 0;
    for (;;) {
        iworker_t *x = &w[dispatched % n_workers];
        wait_idle(x);
        if (dispatched - n_workers >= written) { outbuf_write(&x->out, cornetto_pipeline_out()); written = dispatched - n_workers + 1; }
        if (!x->text) {
            void *p = NULL;
            if (posix_memalign(&p, 2u << 20, reserve + block + 64) != 0) p = NULL;
            CORN_MALLOC_CHK(p);
            madvise(p, reserve + block + 64, MADV_HUGEPAGE);       /* fewer faults while reading, cheaper to release */
            x->text = (uint8_t *)p; x->cap = reserve + block + 64;
        }
        if (carry_len >= block) { *resume = file_pos - carry_len; break; }      /* (cannot happen: a block with no record ends the loop below) */
        uint64_t fresh, want;
        if (ahead_for == dispatched && carry_len <= reserve) {                   /* read ahead: only the tail is missing */
            x->blk = x->text + reserve - carry_len;
            if (carry_len) memcpy(x->blk, carry, carry_len);
            fresh = want = ahead_bytes;
        } else {
            if (ahead_for == dispatched) {                                       /* tail too large: read again behind it */
                if (gz) { *resume = file_pos - ahead_bytes - carry_len; break; } /* (a stream cannot go back: the serial reader takes over) */
                file_pos -= ahead_bytes;
            }
            x->blk = x->text;
            if (carry_len) memcpy(x->blk, carry, carry_len);
            want = size - file_pos < block - carry_len ? size - file_pos : block - carry_len;
            const double t_rd = realtime();
            if (gz) {
                const int64_t k = gzsrc_read(gz, x->blk + carry_len, want, &src_eof);
                if (k < 0) { *resume = file_pos - carry_len; break; }             /* damaged: the serial reader reports it as gzread does */
                fresh = (uint64_t)k;
            } else fresh = read_block(fd, x->blk + carry_len, file_pos, want);
            die_if_read_failed(path);
            TRACE("[ingest] read %.1f MB in %.3f s\n", (double)fresh / 1e6, realtime() - t_rd);
            file_pos += fresh;
        }
        ahead_for = -1;
        const uint64_t block_off = file_pos - fresh - carry_len;
        x->n = carry_len + fresh;
        x->final = gz ? src_eof : (fresh < want || file_pos >= size);
        if (gz && x->final && x->n == 0) { complete = 1; break; }                /* the stream ended exactly at a block end */
        x->ingested = 0;
        x->seq = cornetto_next_batch_seq();
        pthread_mutex_lock(&x->mu);
        x->state = 1;
        pthread_cond_broadcast(&x->cv);
        pthread_mutex_unlock(&x->mu);
        if (n_spawned < n_workers && !x->final) {        /* first block is on its way and there is more: bring up the rest */
            if (n_req > 1) {
                const int avail = corn_gpu_device_count();
                if (avail <= 0) {
                    CORN_ERROR("cannot initialise the GPU: %s", corn_gpu_strerror(avail < 0 ? avail : CORN_E_NOGPU));
                    exit(EXIT_FAILURE);
                }
                if (n_gpus > avail) { n_gpus = avail; n_workers = n_gpus > 1 ? n_gpus : 2; }
            }
            SPAWN_WORKERS(n_workers);
        }
        if (!x->final && reserve) {                      /* read ahead for the next block while this one is shipped and parsed */
            iworker_t *y = &w[(dispatched + 1) % n_workers];
            wait_idle(y);
            if (dispatched + 1 - n_workers >= written) { outbuf_write(&y->out, cornetto_pipeline_out()); written = dispatched + 1 - n_workers + 1; }
            if (!y->text) {
                void *p = NULL;
                if (posix_memalign(&p, 2u << 20, reserve + block + 64) != 0) p = NULL;
                CORN_MALLOC_CHK(p);
                madvise(p, reserve + block + 64, MADV_HUGEPAGE);
                y->text = (uint8_t *)p; y->cap = reserve + block + 64;
            }
            const uint64_t w2 = size - file_pos < block - reserve ? size - file_pos : block - reserve;
            const double t_rd = realtime();
            if (gz) {
                const int64_t k = gzsrc_read(gz, y->text + reserve, w2, &src_eof);
                ahead_bytes = k < 0 ? 0 : (uint64_t)k;                            /* (a damaged stream shows up again at the next read) */
            } else ahead_bytes = read_block(fd, y->text + reserve, file_pos, w2);
            die_if_read_failed(path);
            TRACE("[ingest] read ahead %.1f MB in %.3f s\n", (double)ahead_bytes / 1e6, realtime() - t_rd);
            file_pos += ahead_bytes;
            ahead_for = dispatched + 1;
        }
        pthread_mutex_lock(&x->mu);
        while (!x->ingested) pthread_cond_wait(&x->cv, &x->mu);
        pthread_mutex_unlock(&x->mu);
        ++dispatched;
        if (x->irregular || (x->consumed == 0 && !x->final)) { *resume = block_off; break; }
        carry = x->blk + x->consumed;
        carry_len = x->n - x->consumed;
        if (x->final) { complete = 1; break; }
    }
    for (; written < dispatched; ++written) {
        iworker_t *y = &w[written % n_workers];
        wait_idle(y);
        const double t_wr = realtime();
        outbuf_write(&y->out, cornetto_pipeline_out());
        TRACE("[ingest] wrote output in %.3f s\n", realtime() - t_wr);
    }
    for (int i = 0; i < n_spawned; ++i) {
        pthread_mutex_lock(&w[i].mu);
        w[i].teardown = !complete || !cornetto_fast_exit();      /* the serial reader continues: give the device memory back */
        w[i].state = 2;
        pthread_cond_broadcast(&w[i].cv);
        pthread_mutex_unlock(&w[i].mu);
        pthread_join(w[i].th, NULL);
        outbuf_free(&w[i].out);
        if (!complete || !cornetto_fast_exit()) free(w[i].text);
    }
    fflush(cornetto_pipeline_out());
    free(w);
    close(fd);
    if (gz) gzsrc_close(gz);
    TRACE("[ingest] pipeline done (complete %d)\n", complete);
    return complete;
}
