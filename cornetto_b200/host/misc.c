/* cornetto_b200/host/misc.c -- timing helpers for the stderr footer (reference: src/misc.c:48-70),
 * the shared GPU context, and buffered text output. */
#include <pthread.h>
#include <sys/resource.h>
#include <sys/stat.h>
#include <sys/time.h>

#include "cornetto.h"

double realtime(void)
{
    struct timeval tp;
    gettimeofday(&tp, NULL);
    return (double)tp.tv_sec + (double)tp.tv_usec * 1e-6;
}

double cputime(void)
{
    struct rusage r;
    getrusage(RUSAGE_SELF, &r);
    return (double)(r.ru_utime.tv_sec + r.ru_stime.tv_sec) + 1e-6 * (double)(r.ru_utime.tv_usec + r.ru_stime.tv_usec);
}

long peakrss(void)
{
    struct rusage r;
    getrusage(RUSAGE_SELF, &r);
    return r.ru_maxrss * 1024;
}

static corn_ctx_t *g_ctx;

void cornetto_gpu_die(const char *what, int status)
{
    CORN_ERROR("%s: %s (%s)", what, corn_gpu_strerror(status), g_ctx ? corn_gpu_last_error(g_ctx) : "");
    exit(EXIT_FAILURE);
}

/* CUDA start-up takes 0.4-2.5 s: commands that first parse text start it on a helper thread */
static pthread_t g_init_thread;
static int g_init_started, g_init_status;

static void *gpu_init_main(void *arg)
{
    (void)arg;
    g_init_status = corn_gpu_init(-1, &g_ctx);
    return NULL;
}

void cornetto_gpu_prefetch(void)
{
    if (g_ctx || g_init_started) return;
    if (pthread_create(&g_init_thread, NULL, gpu_init_main, NULL) == 0) g_init_started = 1;
}

corn_ctx_t *cornetto_gpu(void)
{
    if (g_init_started) {
        pthread_join(g_init_thread, NULL);
        g_init_started = 0;
        if (g_init_status != CORN_OK) { g_ctx = NULL; cornetto_gpu_die("cannot initialise the GPU", g_init_status); }
    }
    if (!g_ctx) {
        int r = corn_gpu_init(-1, &g_ctx);
        if (r != CORN_OK) cornetto_gpu_die("cannot initialise the GPU", r);
    }
    return g_ctx;
}

int cornetto_fast_exit(void)
{
    const char *e = getenv("CORNETTO_FAST_EXIT");
    return !(e && atoi(e) == 0);
}

void cornetto_gpu_release(void)
{
    if (g_init_started) { pthread_join(g_init_thread, NULL); g_init_started = 0; }
    if (g_ctx) corn_gpu_destroy(g_ctx);
    g_ctx = NULL;
}

uint64_t cornetto_batch_capacity(const char *path, int n_parts)
{
    uint64_t cap = 1ull << 30;
    const char *e = getenv("CORNETTO_BATCH_MB"), *eb = getenv("CORNETTO_BATCH_BYTES");
    if (eb && atoll(eb) > 0) cap = (uint64_t)atoll(eb);
    else if (e && atoll(e) > 0) cap = (uint64_t)atoll(e) << 20;
    else {
        struct stat st;
        size_t n = strlen(path);
        int gz = n > 3 && strcmp(path + n - 3, ".gz") == 0;
        if (!gz && strcmp(path, "-") != 0 && stat(path, &st) == 0 && S_ISREG(st.st_mode)) {
            /* plain file: the whole input fits one batch per part (records + padding <= file size + slack) */
            uint64_t want = (uint64_t)st.st_size / (uint64_t)(n_parts > 0 ? n_parts : 1) + (1u << 16);
            if (want < cap) cap = want;
        }
    }
    if (cap > CORN_MAX_BATCH_BYTES) cap = CORN_MAX_BATCH_BYTES;
    return cap;
}

/* ---- output ------------------------------------------------------------------------------------ */
void outbuf_init(outbuf_t *o, FILE *fp)
{
    o->cap = 1 << 20; o->n = 0; o->fp = fp;
    o->buf = (char *)malloc(o->cap);
    CORN_MALLOC_CHK(o->buf);
}
void outbuf_flush(outbuf_t *o) { if (o->fp && o->n) { fwrite(o->buf, 1, o->n, o->fp); o->n = 0; } }
void outbuf_free(outbuf_t *o) { outbuf_flush(o); if (o->fp) fflush(o->fp); free(o->buf); o->buf = NULL; }
static FILE *g_pipeline_out;
void  cornetto_set_pipeline_out(FILE *fp) { g_pipeline_out = fp; }
FILE *cornetto_pipeline_out(void) { return g_pipeline_out ? g_pipeline_out : stdout; }
static uint64_t g_batch_seq;
uint64_t cornetto_next_batch_seq(void) { return g_batch_seq++; }      /* called by the one dispatching thread */

void outbuf_write(outbuf_t *o, FILE *fp) { if (o->n) fwrite(o->buf, 1, o->n, fp); o->n = 0; }
static inline void need(outbuf_t *o, size_t k)
{
    if (o->n + k <= o->cap) return;
    if (o->fp) { outbuf_flush(o); if (k <= o->cap) return; }
    while (o->n + k > o->cap) o->cap *= 2;
    o->buf = (char *)realloc(o->buf, o->cap);
    CORN_MALLOC_CHK(o->buf);
}
void outbuf_str(outbuf_t *o, const char *s, size_t len)
{
    if (o->fp && len > o->cap / 2) { outbuf_flush(o); fwrite(s, 1, len, o->fp); return; }
    need(o, len);
    memcpy(o->buf + o->n, s, len);
    o->n += len;
}
void outbuf_chr(outbuf_t *o, char c) { need(o, 1); o->buf[o->n++] = c; }
void outbuf_u64(outbuf_t *o, uint64_t v)
{
    char t[24]; int k = 0;
    do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    need(o, (size_t)k);
    while (k) o->buf[o->n++] = t[--k];
}
void outbuf_i32(outbuf_t *o, int32_t v)
{
    if (v < 0) { outbuf_chr(o, '-'); outbuf_u64(o, (uint64_t)(-(int64_t)v)); }
    else outbuf_u64(o, (uint64_t)v);
}

/* ---- parallel formatting ------------------------------------------------------------------------
 * Millions of output lines (telofind on a genome: 1.5 M) take longer to format than the GPU needs to
 * find them.  The items are split into contiguous ranges, each formatted into its own memory buffer by
 * a helper thread, and the buffers are appended in range order: the bytes are the same as the serial
 * loop's. */
typedef struct { outbuf_t ob; uint64_t begin, end; format_range_fn fn; void *arg; pthread_t th; int started; } fmt_job_t;

static void *fmt_main(void *p)
{
    fmt_job_t *j = (fmt_job_t *)p;
    j->fn(&j->ob, j->begin, j->end, j->arg);
    return NULL;
}

void outbuf_format_parallel(outbuf_t *ob, uint64_t n_items, format_range_fn fn, void *arg)
{
    enum { MAX_T = 6 };
    int T = n_items < 200000 ? 1 : MAX_T;
    if (T == 1) { fn(ob, 0, n_items, arg); return; }
    fmt_job_t job[MAX_T];
    const uint64_t per = (n_items + T - 1) / T;
    for (int t = 0; t < T; ++t) {
        job[t].begin = per * (uint64_t)t < n_items ? per * (uint64_t)t : n_items;
        job[t].end = job[t].begin + per < n_items ? job[t].begin + per : n_items;
        job[t].fn = fn; job[t].arg = arg;
        outbuf_init(&job[t].ob, NULL);
        job[t].started = t > 0 && pthread_create(&job[t].th, NULL, fmt_main, &job[t]) == 0;
    }
    for (int t = 0; t < T; ++t) if (!job[t].started) fmt_main(&job[t]);      /* range 0, and any thread that could not start */
    for (int t = 0; t < T; ++t) {
        if (job[t].started) pthread_join(job[t].th, NULL);
        outbuf_str(ob, job[t].ob.buf, job[t].ob.n);
        job[t].ob.n = 0;
        outbuf_free(&job[t].ob);
    }
}
