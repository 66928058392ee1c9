/* cornetto_b200/host/depthtxt.c -- the two per-base depth tables of `noboringbits` / `boringbits` into memory.
 *
 * Replaces get_depths() (src/boringbits_main.c:179-293): the reference reads both files record by record with
 * fscanf("%s\t%d\t%d\t%d\n") on its one thread.  The files hold one line per BASE -- tens of GB each for a human
 * assembly -- so the text, not the window scan, is where the command spends its time.  Two readers:
 *
 *   depthtxt_load_serial    one pass over both memory-mapped files, token by token exactly as fscanf's format string
 *                           consumes them (whitespace-separated tokens whatever the line structure), every check,
 *                           message and exit of the reference in the reference's order.
 *   depthtxt_load_parallel  the files cut into blocks at line ends, the blocks parsed on several threads, the
 *                           per-block contig runs stitched, the two files compared run by run, the values copied into
 *                           place.  It only accepts input the serial reader would accept WITHOUT a message: every line
 *                           exactly four tokens, plain digit strings, end = start + 1, positions rising by one from 0
 *                           within a contig, depths <= 65535, the same contig runs in both files.  Anything else --
 *                           including every input the reference rejects or warns about -- makes it return -1 before
 *                           anything was printed, and the caller runs the serial reader, which then behaves as the
 *                           reference does.
 * The sums the mean depths are made of are integers below 2^53 either way: a double adds them exactly in any order. */
#include <fcntl.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "cornetto.h"

void depthtxt_open(depth_text_t *t, const char *path)
{
    const int fd = open(path, O_RDONLY);
    struct stat sb;
    if (fd < 0 || fstat(fd, &sb) != 0) { CORN_ERROR("Could not to open file %s: %s", path, strerror(errno)); exit(EXIT_FAILURE); }
    t->size = (size_t)sb.st_size;
    t->map = t->size ? mmap(NULL, t->size, PROT_READ, MAP_PRIVATE, fd, 0) : NULL;
    if (t->size && t->map == MAP_FAILED) { CORN_ERROR("Could not to open file %s: %s", path, strerror(errno)); exit(EXIT_FAILURE); }
    if (t->size) madvise(t->map, t->size, MADV_SEQUENTIAL);
    close(fd);
    t->p = (const char *)t->map; t->e = t->p + t->size;
}

void depthtxt_close(depth_text_t *t)
{
    if (t->size) munmap(t->map, t->size);
    t->map = NULL; t->size = 0;
}

void depth_table_free(depth_table_t *T)
{
    for (size_t i = 0; i < T->n_ctg; ++i) free(T->ctg[i].name);
    free(T->ctg); free(T->depth); free(T->mq);
    memset(T, 0, sizeof *T);
}

static int is_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

/* ---- serial reader ------------------------------------------------------------------------------------------------ */
/* one record: 4 = all fields converted, EOF = nothing left, else the number of fields converted (what fscanf returns) */
static int tok_record(depth_text_t *t, const char **name, size_t *name_len, int *st, int *end, int *depth)
{
    while (t->p < t->e && is_ws(*t->p)) ++t->p;
    if (t->p >= t->e) return EOF;
    *name = t->p;
    while (t->p < t->e && !is_ws(*t->p)) ++t->p;
    *name_len = (size_t)(t->p - *name);
    int *dst[3] = { st, end, depth };
    for (int k = 0; k < 3; ++k) {
        while (t->p < t->e && is_ws(*t->p)) ++t->p;
        if (t->p >= t->e) return 1 + k;                       /* (fscanf: input failure after k + 1 conversions) */
        const char *q = t->p;
        int neg = 0;
        if (*q == '-' || *q == '+') { neg = *q == '-'; ++q; }
        if (q >= t->e || *q < '0' || *q > '9') return 1 + k;  /* matching failure */
        long long v = 0;
        while (q < t->e && *q >= '0' && *q <= '9') { v = v * 10 + (*q - '0'); if (v > 0x7fffffffLL) v = 0x7fffffffLL; ++q; }
        *dst[k] = (int)(neg ? -v : v);
        t->p = q;
    }
    return 4;
}

void depthtxt_load_serial(depth_text_t *t1, depth_text_t *t2, depth_table_t *T)
{
    memset(T, 0, sizeof *T);
    t1->p = (const char *)t1->map; t2->p = (const char *)t2->map;
    size_t m_ctg = 0;
    uint64_t cap = 1u << 20;
    T->depth = (uint16_t *)malloc(cap * sizeof(uint16_t)); T->mq = (uint16_t *)malloc(cap * sizeof(uint16_t));
    CORN_MALLOC_CHK(T->depth); CORN_MALLOC_CHK(T->mq);
    const char *prev = NULL;
    size_t prev_len = 0;
    int prev_pos = 0;
    for (;;) {
        const char *n1, *n2;
        size_t l1, l2;
        int st1, st2, e1, e2, d1, d2;
        int ret = tok_record(t1, &n1, &l1, &st1, &e1, &d1);
        if (ret == EOF) break;
        if (ret != 4) { CORN_ERROR("The depth files should have 4 columns. Had %d.", ret); exit(EXIT_FAILURE); }
        ret = tok_record(t2, &n2, &l2, &st2, &e2, &d2);
        if (ret == EOF) { CORN_ERROR("%s", "The two files are not in the same order"); exit(EXIT_FAILURE); }
        if (ret != 4) { CORN_ERROR("The depth files should have 4 columns. Had %d.", ret); exit(EXIT_FAILURE); }
        if (l1 != l2 || memcmp(n1, n2, l1) != 0 || st1 != st2 || e1 != e2) { CORN_ERROR("%s", "The two files are not in the same order"); exit(EXIT_FAILURE); }
        if (!prev || l1 != prev_len || memcmp(n1, prev, l1) != 0) {
            prev = n1; prev_len = l1;
            if (T->n_ctg == m_ctg) { m_ctg = m_ctg ? m_ctg * 2 : 16; T->ctg = (depth_ctg_t *)realloc(T->ctg, m_ctg * sizeof(depth_ctg_t)); CORN_MALLOC_CHK(T->ctg); }
            T->ctg[T->n_ctg].name = strndup(n1, l1); CORN_MALLOC_CHK(T->ctg[T->n_ctg].name);
            T->ctg[T->n_ctg].off = T->n_tot; T->ctg[T->n_ctg].len = 0;
            ++T->n_ctg;
            prev_pos = 0;
        } else {
            if (prev_pos + 1 != st1) { CORN_ERROR("The depth files should be incremantal at one base resolution. Found %d to %d", prev_pos, st1); exit(EXIT_FAILURE); }
            ++prev_pos;
        }
        if (st1 + 1 != e1) { CORN_ERROR("The depth files should have end=start+1. Found %d to %d", st1, e1); exit(EXIT_FAILURE); }
        if (d1 > 65535) {
            fprintf(stderr, "[%s::WARNING]\033[1;33m The depth at %.*s:%d-%d was truncated to 65535. Found %d\033[0m At %s:%d\n", "get_depths", (int)l1, n1, st1, e1, d1, __FILE__, __LINE__);
            d1 = 65535;
        }
        if (d2 > 65535) {
            fprintf(stderr, "[%s::WARNING]\033[1;33m The depth at %.*s:%d-%d was truncated to 65535. Found %d\033[0m At %s:%d\n", "get_depths", (int)l2, n2, st2, e2, d2, __FILE__, __LINE__);
            d2 = 65535;
        }
        if (T->n_tot == cap) {
            cap *= 2;
            T->depth = (uint16_t *)realloc(T->depth, cap * sizeof(uint16_t)); T->mq = (uint16_t *)realloc(T->mq, cap * sizeof(uint16_t));
            CORN_MALLOC_CHK(T->depth); CORN_MALLOC_CHK(T->mq);
        }
        if (T->ctg[T->n_ctg - 1].len == 0x7fffffffu) { CORN_ERROR("contig %s is too long", T->ctg[T->n_ctg - 1].name); exit(EXIT_FAILURE); }
        T->depth[T->n_tot] = (uint16_t)d1; T->mq[T->n_tot] = (uint16_t)d2;
        ++T->n_tot; ++T->ctg[T->n_ctg - 1].len;
        T->tot_depth += d1; T->tot_mq += d2;
    }
}

/* ---- parallel reader ---------------------------------------------------------------------------------------------- */
typedef struct { const char *name; uint32_t name_len; uint32_t st0; uint64_t n; uint64_t at; } run_t;   /* at: first value, in the file's value order */
typedef struct {
    const char *file, *file_end, *b, *e;      /* the file and this block's byte range: lines that START in [b, e) */
    run_t *run; size_t n_run, m_run;
    uint16_t *val; uint64_t n_val, sum;
    int bad;
} blk_t;

static int is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

static void parse_block(blk_t *k)
{
    const char *p = k->b, *fe = k->file_end;
    if (p > k->file) {                        /* the line that straddles the block start belongs to the block before */
        p = (const char *)memchr(p - 1, '\n', (size_t)(fe - (p - 1)));
        if (!p) return;
        ++p;
    }
    const uint64_t cap = (uint64_t)(k->e - k->b) / 8 + 2;        /* a line is at least 8 bytes: "a 0 1 0\n" */
    k->val = (uint16_t *)malloc(cap * sizeof(uint16_t));
    if (!k->val) { k->bad = 1; return; }
    run_t *cur = NULL;
    uint64_t n = 0, sum = 0;
    while (p < k->e && p < fe) {
        while (p < fe && is_blank(*p)) ++p;
        if (p >= fe) break;
        if (*p == '\n') { ++p; continue; }    /* an empty line is only whitespace to fscanf */
        const char *name = p;
        while (p < fe && !is_ws(*p)) ++p;
        const size_t nl = (size_t)(p - name);
        uint32_t v[3];
        for (int f = 0; f < 3; ++f) {
            const char *q = p;
            while (p < fe && is_blank(*p)) ++p;
            if (p == q || p >= fe || *p < '0' || *p > '9') { k->bad = 1; return; }
            uint64_t x = 0;
            const char *d0 = p;
            while (p < fe && *p >= '0' && *p <= '9') { x = x * 10 + (uint64_t)(*p - '0'); ++p; }
            if (p - d0 > 10 || x > 0x7fffffffu) { k->bad = 1; return; }
            v[f] = (uint32_t)x;
        }
        while (p < fe && is_blank(*p)) ++p;
        if (p < fe) { if (*p != '\n') { k->bad = 1; return; } ++p; }
        if (nl >= 10000 || v[1] != v[0] + 1 || v[2] > 65535u || n >= cap) { k->bad = 1; return; }   /* (the reference's name buffer holds 10000 bytes) */
        if (!cur || cur->name_len != nl || memcmp(cur->name, name, nl) != 0) {
            if (cur && v[0] != 0) { k->bad = 1; return; }        /* a contig that does not start at 0: the serial reader decides */
            if (k->n_run == k->m_run) {
                k->m_run = k->m_run ? k->m_run * 2 : 16;
                run_t *r = (run_t *)realloc(k->run, k->m_run * sizeof(run_t));
                if (!r) { k->bad = 1; return; }
                k->run = r;
            }
            cur = &k->run[k->n_run++];
            cur->name = name; cur->name_len = (uint32_t)nl; cur->st0 = v[0]; cur->n = 0; cur->at = n;
        } else if ((uint64_t)v[0] != (uint64_t)cur->st0 + cur->n) { k->bad = 1; return; }
        ++cur->n;
        k->val[n++] = (uint16_t)v[2];
        sum += v[2];
    }
    k->n_val = n; k->sum = sum;
    if (n) { uint16_t *v2 = (uint16_t *)realloc(k->val, n * sizeof(uint16_t)); if (v2) k->val = v2; }   /* give the unused tail back */
    else { free(k->val); k->val = NULL; }
}

typedef struct { blk_t *blk; size_t n_blk; size_t next; pthread_mutex_t mu; int phase; uint16_t *dst[2]; size_t n_blk_file[2]; uint64_t *blk_at; } pool_t;

static void *pool_worker(void *arg)
{
    pool_t *P = (pool_t *)arg;
    for (;;) {
        pthread_mutex_lock(&P->mu);
        const size_t i = P->next++;
        pthread_mutex_unlock(&P->mu);
        if (i >= P->n_blk) return NULL;
        blk_t *k = &P->blk[i];
        if (P->phase == 0) parse_block(k);
        else if (k->n_val) memcpy(P->dst[i >= P->n_blk_file[0]] + P->blk_at[i], k->val, k->n_val * sizeof(uint16_t));
    }
}

static void pool_run(pool_t *P, int threads, int phase)
{
    pthread_t th[64];
    if (threads > 64) threads = 64;
    P->phase = phase; P->next = 0;
    int started = 0;
    for (int t = 0; t + 1 < threads; ++t) if (pthread_create(&th[started], NULL, pool_worker, P) == 0) ++started;
    pool_worker(P);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

/* the runs of one file's blocks, in order, stitched across block seams; -1 = leave it to the serial reader */
static int stitch(blk_t *blk, size_t n_blk, run_t **out, size_t *n_out)
{
    size_t total = 0;
    for (size_t i = 0; i < n_blk; ++i) total += blk[i].n_run;
    run_t *r = (run_t *)malloc((total + 1) * sizeof(run_t));
    if (!r) return -1;
    size_t n = 0;
    uint64_t at = 0;
    for (size_t i = 0; i < n_blk; ++i) {
        for (size_t j = 0; j < blk[i].n_run; ++j) {
            const run_t *s = &blk[i].run[j];
            if (n && r[n - 1].name_len == s->name_len && memcmp(r[n - 1].name, s->name, s->name_len) == 0) {
                if ((uint64_t)s->st0 != (uint64_t)r[n - 1].st0 + r[n - 1].n) { free(r); return -1; }
                r[n - 1].n += s->n;
            } else {
                if (n && s->st0 != 0) { free(r); return -1; }
                r[n] = *s; r[n].at = at + s->at;
                ++n;
            }
            if (r[n - 1].n > 0x7fffffffu) { free(r); return -1; }
        }
        at += blk[i].n_val;
    }
    *out = r; *n_out = n;
    return 0;
}

int depthtxt_load_parallel(const depth_text_t *t1, const depth_text_t *t2, int threads, size_t block_bytes, depth_table_t *T)
{
    memset(T, 0, sizeof *T);
    if (threads < 1) threads = 1;
    if (block_bytes < 64) block_bytes = 64;
    const depth_text_t *tt[2] = { t1, t2 };
    size_t nb[2];
    for (int f = 0; f < 2; ++f) nb[f] = (tt[f]->size + block_bytes - 1) / block_bytes;
    pool_t P;
    memset(&P, 0, sizeof P);
    P.n_blk = nb[0] + nb[1];
    P.n_blk_file[0] = nb[0]; P.n_blk_file[1] = nb[1];
    P.blk = (blk_t *)calloc(P.n_blk + 1, sizeof(blk_t));
    P.blk_at = (uint64_t *)calloc(P.n_blk + 1, sizeof(uint64_t));
    if (!P.blk || !P.blk_at) { free(P.blk); free(P.blk_at); return -1; }
    pthread_mutex_init(&P.mu, NULL);
    /* blocks of the two files alternate in the queue so both are read front to back at the same pace */
    for (int f = 0; f < 2; ++f)
        for (size_t i = 0; i < nb[f]; ++i) {
            blk_t *k = &P.blk[(f ? nb[0] : 0) + i];
            k->file = (const char *)tt[f]->map; k->file_end = k->file + tt[f]->size;
            k->b = k->file + i * block_bytes;
            k->e = (i + 1 == nb[f]) ? k->file_end : k->b + block_bytes;
        }
    pool_run(&P, threads, 0);
    int rc = 0;
    run_t *r[2] = { NULL, NULL };
    size_t nr[2] = { 0, 0 };
    uint64_t nv[2] = { 0, 0 }, sum[2] = { 0, 0 };
    for (size_t i = 0; i < P.n_blk; ++i) {
        const int f = i >= nb[0];
        if (P.blk[i].bad) rc = -1;
        P.blk_at[i] = nv[f];
        nv[f] += P.blk[i].n_val; sum[f] += P.blk[i].sum;
    }
    if (rc == 0) rc = stitch(P.blk, nb[0], &r[0], &nr[0]);
    if (rc == 0) rc = stitch(P.blk + nb[0], nb[1], &r[1], &nr[1]);
    if (rc == 0 && (nr[0] != nr[1] || nv[0] != nv[1])) rc = -1;
    /* the two files must hold the same contig runs (names, first positions, lengths): what the reference checks record by record */
    for (size_t i = 0; rc == 0 && i < nr[0]; ++i)
        if (r[0][i].name_len != r[1][i].name_len || memcmp(r[0][i].name, r[1][i].name, r[0][i].name_len) != 0 ||
            r[0][i].st0 != r[1][i].st0 || r[0][i].n != r[1][i].n) rc = -1;
    /* a contig of more than one record has to start at 0 (the reference compares its second position with 1) */
    for (size_t i = 0; rc == 0 && i < nr[0]; ++i) if (r[0][i].st0 != 0 && r[0][i].n > 1) rc = -1;
    if (rc == 0 && nv[0]) {
        T->depth = (uint16_t *)malloc(nv[0] * sizeof(uint16_t)); T->mq = (uint16_t *)malloc(nv[0] * sizeof(uint16_t));
        T->ctg = (depth_ctg_t *)calloc(nr[0], sizeof(depth_ctg_t));
        if (!T->depth || !T->mq || !T->ctg) { free(T->depth); free(T->mq); free(T->ctg); memset(T, 0, sizeof *T); rc = -1; }
    }
    if (rc == 0 && nv[0]) {
        P.dst[0] = T->depth; P.dst[1] = T->mq;
        pool_run(&P, threads, 1);
        for (size_t i = 0; i < nr[0]; ++i) {
            T->ctg[i].name = strndup(r[0][i].name, r[0][i].name_len); CORN_MALLOC_CHK(T->ctg[i].name);
            T->ctg[i].off = r[0][i].at; T->ctg[i].len = (uint32_t)r[0][i].n;
        }
        T->n_ctg = nr[0]; T->n_tot = nv[0];
        T->tot_depth = (double)sum[0]; T->tot_mq = (double)sum[1];
    }
    for (size_t i = 0; i < P.n_blk; ++i) { free(P.blk[i].run); free(P.blk[i].val); }
    free(r[0]); free(r[1]); free(P.blk); free(P.blk_at);
    pthread_mutex_destroy(&P.mu);
    return rc;
}
