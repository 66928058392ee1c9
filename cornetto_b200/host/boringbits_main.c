/* cornetto_b200/host/boringbits_main.c -- `cornetto noboringbits cov-total.bg -q cov-mq20.bg` (and `boringbits`).
 *
 * Same contract as boringbits_main(), src/boringbits_main.c:561-660: the same getopt_long option table (-q -w -i -L -H
 * -Q -m -e -v -h -V and the ignored batch/thread options), help text, input checks and their ERROR texts, the parameter
 * block on stderr, and on stdout  name \t st \t end \t depth \t mq_depth  per selected window after the per-contig
 * lines of print_fun_bits() (:425-446) / the selection of print_boring_bits() (:465-485).
 * The two depth files are read into one uint16 value per base as get_depths() does (:179-293); what the reference
 * then does per window -- re-summing 2500 values of both arrays every 50 bases (:340-363) and testing them -- is one
 * corn_gpu_depthwin() call (csrc/depthwin.c: one pass over the two arrays). */
#include <fcntl.h>
#include <getopt.h>
#include <math.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "cornetto.h"

static const struct option bits_options[] = {
    { "threads", required_argument, 0, 't' }, { "batchsize", required_argument, 0, 'K' }, { "max-bytes", required_argument, 0, 'B' },
    { "verbose", required_argument, 0, 'v' }, { "help", no_argument, 0, 'h' }, { "version", no_argument, 0, 'V' },
    { "output", required_argument, 0, 'o' }, { "debug-break", required_argument, 0, 0 }, { "profile-cpu", required_argument, 0, 0 },
    { "accel", required_argument, 0, 0 }, { "qual", required_argument, 0, 'q' }, { "window-size", required_argument, 0, 'w' },
    { "window-inc", required_argument, 0, 'i' }, { "low-thresh", required_argument, 0, 'L' }, { "high-thresh", required_argument, 0, 'H' },
    { "low-mq-thresh", required_argument, 0, 'Q' }, { "min-ctg-len", required_argument, 0, 'm' }, { "edge-len", required_argument, 0, 'e' },
    { 0, 0, 0, 0 } };

typedef struct {
    int window_size, window_inc;
    float low_cov_thresh, high_cov_thresh, low_mq_cov_thresh;
    int min_ctg_len, edge_len, verbose;
} bits_opt_t;

static void bits_help(FILE *fp, const bits_opt_t *o)
{   /* print_help_msg(), :86-113 */
    fprintf(fp, "Usage: cornetto boringbits cov-total.bg -q cov-mq20.bg\n");
    fprintf(fp, "\nbasic options:\n");
    fprintf(fp, "   -q FILE                    depth file with high mapq read coverage\n");
    fprintf(fp, "   -w INT                     window size [%d]\n", o->window_size);
    fprintf(fp, "   -i INT                     window increment [%d]\n", o->window_inc);
    fprintf(fp, "   -L FLOAT                   low coverage threshold factor [%.1f]\n", o->low_cov_thresh);
    fprintf(fp, "   -H FLOAT                   high coverage threshold factor [%.1f]\n", o->high_cov_thresh);
    fprintf(fp, "   -Q FLOAT                   mapq low coverage threshold factor [%.1f]\n", o->low_mq_cov_thresh);
    fprintf(fp, "   -m INT                     minimum contig length [%d]\n", o->min_ctg_len);
    fprintf(fp, "   -e INT                     edge length to ignore [%d]\n", o->edge_len);
    fprintf(fp, "   -h                         help\n");
    fprintf(fp, "   --verbose INT              verbosity level [%d]\n", o->verbose);
}

/* ---- the two files, token by token: fscanf("%s\t%d\t%d\t%d\n") reads whitespace-separated tokens, whatever the line
 * structure (:202,211).  A memory map and a hand-written integer reader instead of fscanf: the files hold one line
 * per BASE (tens of GB for a human assembly). */
typedef struct { const char *p, *e; void *map; size_t size; } tok_t;

static void tok_open(tok_t *t, const char *path)
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

static int is_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

/* one record: 4 = all fields converted, EOF = nothing left, else the number of fields converted (what fscanf returns) */
static int tok_record(tok_t *t, const char **name, size_t *name_len, int *st, int *end, int *depth)
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

typedef struct { char *name; uint64_t off; uint32_t len; } ctg_t;

int boringbits_main(int argc, char *argv[], int boring)
{
    bits_opt_t o;
    o.window_size = 2500; o.window_inc = 50;                 /* init_optp(), :543-559 */
    o.low_cov_thresh = 0.4f; o.high_cov_thresh = 2.5f; o.low_mq_cov_thresh = 0.4f;
    o.min_ctg_len = 1000000; o.edge_len = 100000; o.verbose = 4;
    const char *covmq = NULL;
    FILE *fp_help = stderr;
    int c, longindex = 0;
    optind = 1;
    while ((c = getopt_long(argc, argv, "t:B:K:v:o:q:Q:H:L:w:i:e:m:hV", bits_options, &longindex)) >= 0) {
        if (c == 'V') { fprintf(stdout, "cornetto %s\n", CORNETTO_VERSION); exit(EXIT_SUCCESS); }
        else if (c == 'h') fp_help = stdout;
        else if (c == 'v') o.verbose = atoi(optarg);
        else if (c == 'q') covmq = optarg;
        else if (c == 'w') o.window_size = atoi(optarg);
        else if (c == 'i') o.window_inc = atoi(optarg);
        else if (c == 'L') o.low_cov_thresh = (float)atof(optarg);
        else if (c == 'H') o.high_cov_thresh = (float)atof(optarg);
        else if (c == 'Q') o.low_mq_cov_thresh = (float)atof(optarg);
        else if (c == 'm') o.min_ctg_len = atoi(optarg);
        else if (c == 'e') o.edge_len = atoi(optarg);
        else if (c == 'B' && atof(optarg) <= 0) { CORN_ERROR("%s", "Maximum number of bytes should be larger than 0."); exit(EXIT_FAILURE); }
        else if (c == 'K' && atoi(optarg) < 1) { CORN_ERROR("Batch size should larger than 0. You entered %d", atoi(optarg)); exit(EXIT_FAILURE); }
        else if (c == 't' && atoi(optarg) < 1) { CORN_ERROR("Number of threads should larger than 0. You entered %d", atoi(optarg)); exit(EXIT_FAILURE); }
    }
    if (argc - optind != 1 || fp_help == stdout || covmq == NULL) {
        bits_help(fp_help, &o);
        exit(fp_help == stdout ? EXIT_SUCCESS : EXIT_FAILURE);
    }
    const char *covtotal = argv[optind];
    if (o.window_size < 1 || o.window_inc < 1) {             /* (the reference divides by zero / trips its asserts here) */
        CORN_ERROR("%s", "window size and window increment must be positive");
        exit(EXIT_FAILURE);
    }
    cornetto_gpu_prefetch();                                 /* the driver starts while the text is parsed */

    /* ---- get_depths(), :179-293 ---- */
    tok_t t1, t2;
    tok_open(&t1, covtotal);
    tok_open(&t2, covmq);
    ctg_t *ctg = NULL;
    size_t n_ctg = 0, m_ctg = 0;
    uint64_t n_tot = 0, cap = 1u << 20;
    uint16_t *depth = (uint16_t *)malloc(cap * sizeof(uint16_t)), *mq = (uint16_t *)malloc(cap * sizeof(uint16_t));
    CORN_MALLOC_CHK(depth); CORN_MALLOC_CHK(mq);
    const char *prev = NULL;
    size_t prev_len = 0;
    int prev_pos = 0;
    double tot_depth = 0, tot_mq = 0, tot_len = 0;
    for (;;) {
        const char *n1, *n2;
        size_t l1, l2;
        int st1, st2, e1, e2, d1, d2;
        int ret = tok_record(&t1, &n1, &l1, &st1, &e1, &d1);
        if (ret == EOF) break;
        if (ret != 4) { CORN_ERROR("The depth files should have 4 columns. Had %d.", ret); exit(EXIT_FAILURE); }
        ret = tok_record(&t2, &n2, &l2, &st2, &e2, &d2);
        if (ret == EOF) { CORN_ERROR("%s", "The two files are not in the same order"); exit(EXIT_FAILURE); }
        if (ret != 4) { CORN_ERROR("The depth files should have 4 columns. Had %d.", ret); exit(EXIT_FAILURE); }
        if (l1 != l2 || memcmp(n1, n2, l1) != 0 || st1 != st2 || e1 != e2) { CORN_ERROR("%s", "The two files are not in the same order"); exit(EXIT_FAILURE); }
        if (!prev || l1 != prev_len || memcmp(n1, prev, l1) != 0) {
            prev = n1; prev_len = l1;
            if (n_ctg == m_ctg) { m_ctg = m_ctg ? m_ctg * 2 : 16; ctg = (ctg_t *)realloc(ctg, m_ctg * sizeof(ctg_t)); CORN_MALLOC_CHK(ctg); }
            ctg[n_ctg].name = strndup(n1, l1); CORN_MALLOC_CHK(ctg[n_ctg].name);
            ctg[n_ctg].off = n_tot; ctg[n_ctg].len = 0;
            ++n_ctg;
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
        if (n_tot == cap) {
            cap *= 2;
            depth = (uint16_t *)realloc(depth, cap * sizeof(uint16_t)); mq = (uint16_t *)realloc(mq, cap * sizeof(uint16_t));
            CORN_MALLOC_CHK(depth); CORN_MALLOC_CHK(mq);
        }
        if (ctg[n_ctg - 1].len == 0x7fffffffu) { CORN_ERROR("contig %s is too long", ctg[n_ctg - 1].name); exit(EXIT_FAILURE); }
        depth[n_tot] = (uint16_t)d1; mq[n_tot] = (uint16_t)d2;
        ++n_tot; ++ctg[n_ctg - 1].len;
        tot_depth += d1; tot_mq += d2; tot_len++;
    }
    const int mean_depth = (int)round(tot_depth / tot_len), mean_mq_depth = (int)round(tot_mq / tot_len);

    /* the_boring_bits(), :504-513 */
    fprintf(stderr, "Number of contigs: %d\n", (int)n_ctg);
    fprintf(stderr, "Average depth: %d\n", mean_depth);
    fprintf(stderr, "Average mq depth: %d\n", mean_mq_depth);
    fprintf(stderr, "Window size: %d\n", o.window_size);
    fprintf(stderr, "Window increment: %d\n", o.window_inc);
    fprintf(stderr, "Low coverage threshold: %.1fx%d\n", o.low_cov_thresh, mean_depth);
    fprintf(stderr, "High coverage threshold: %.1fx%d\n", o.high_cov_thresh, mean_depth);
    fprintf(stderr, "Low mapq coverage threshold: %.1f\n", o.low_mq_cov_thresh);
    fprintf(stderr, "Min contig length: %d\n", o.min_ctg_len);
    fprintf(stderr, "Edge length: %d\n", o.edge_len);

    corn_depth_windows_t w;
    w.win = NULL; w.n_win = 0; w._owner = NULL;
    if (n_ctg) {
        uint64_t *off = (uint64_t *)malloc(n_ctg * sizeof(uint64_t));
        uint32_t *len = (uint32_t *)malloc(n_ctg * sizeof(uint32_t));
        CORN_MALLOC_CHK(off); CORN_MALLOC_CHK(len);
        for (size_t i = 0; i < n_ctg; ++i) { off[i] = ctg[i].off; len[i] = ctg[i].len; }
        corn_depth_batch_t b;
        b.depth = depth; b.mq_depth = mq; b.offset = off; b.length = len; b.n_ctg = (uint32_t)n_ctg; b.n_total = n_tot;
        corn_depth_params_t prm;
        prm.window_size = o.window_size; prm.window_inc = o.window_inc;
        prm.thresh_low_depth = (int)round(o.low_cov_thresh * mean_depth);       /* :524-525 (float x int, rounded as a double) */
        prm.thresh_high_depth = (int)round(o.high_cov_thresh * mean_depth);
        prm.low_mq_cov_thresh = o.low_mq_cov_thresh;
        prm.edge_len = o.edge_len; prm.min_ctg_len = o.min_ctg_len; prm.boring = boring;
        corn_ctx_t *ctx = cornetto_gpu();
        const int r = corn_gpu_depthwin(ctx, &b, &prm, &w);
        if (r != CORN_OK) cornetto_gpu_die(boring ? "boringbits" : "noboringbits", r);
        free(off); free(len);
    }
    /* print_fun_bits() / print_boring_bits(): the windows come back in (contig, start) order */
    outbuf_t ob;
    outbuf_init(&ob, stdout);
    uint64_t k = 0;
    for (size_t i = 0; i < n_ctg; ++i) {
        const char *name = ctg[i].name;
        const size_t nl = strlen(name);
        if (!boring) {
            if ((int)ctg[i].len < o.min_ctg_len) {                                /* small contigs are always fun (:429-430) */
                outbuf_str(&ob, name, nl); outbuf_chr(&ob, '\t'); outbuf_i32(&ob, 0); outbuf_chr(&ob, '\t'); outbuf_i32(&ob, o.min_ctg_len); outbuf_str(&ob, "\t.\t.\n", 5);
            } else {
                outbuf_str(&ob, name, nl); outbuf_chr(&ob, '\t'); outbuf_i32(&ob, 0); outbuf_chr(&ob, '\t'); outbuf_i32(&ob, o.edge_len); outbuf_str(&ob, "\t.\t.\n", 5);
                outbuf_str(&ob, name, nl); outbuf_chr(&ob, '\t'); outbuf_i32(&ob, (int)ctg[i].len - o.edge_len); outbuf_chr(&ob, '\t'); outbuf_i32(&ob, (int)ctg[i].len); outbuf_str(&ob, "\t.\t.\n", 5);
            }
        }
        for (; k < w.n_win && w.win[k].ctg == i; ++k) {
            outbuf_str(&ob, name, nl);
            outbuf_chr(&ob, '\t'); outbuf_i32(&ob, (int)w.win[k].st);
            outbuf_chr(&ob, '\t'); outbuf_i32(&ob, (int)w.win[k].end);
            outbuf_chr(&ob, '\t'); outbuf_i32(&ob, w.win[k].depth);
            outbuf_chr(&ob, '\t'); outbuf_i32(&ob, w.win[k].mq_depth);
            outbuf_chr(&ob, '\n');
        }
    }
    outbuf_flush(&ob);
    outbuf_free(&ob);
    corn_gpu_depth_windows_free(&w);
    if (!cornetto_fast_exit()) {
        for (size_t i = 0; i < n_ctg; ++i) free(ctg[i].name);
        free(ctg); free(depth); free(mq);
        if (t1.size) munmap(t1.map, t1.size);
        if (t2.size) munmap(t2.map, t2.size);
    }
    return 0;
}
