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

/* the selected windows of one contig as text: name \t st \t end \t depth \t mq_depth */
typedef struct { const corn_depth_window_t *win; const char *name; size_t nl; } fmt_bits_t;

static void fmt_bits(outbuf_t *ob, uint64_t begin, uint64_t end, void *arg)
{
    const fmt_bits_t *a = (const fmt_bits_t *)arg;
    for (uint64_t k = begin; k < end; ++k) {
        outbuf_str(ob, a->name, a->nl);
        outbuf_chr(ob, '\t'); outbuf_i32(ob, (int)a->win[k].st);
        outbuf_chr(ob, '\t'); outbuf_i32(ob, (int)a->win[k].end);
        outbuf_chr(ob, '\t'); outbuf_i32(ob, a->win[k].depth);
        outbuf_chr(ob, '\t'); outbuf_i32(ob, a->win[k].mq_depth);
        outbuf_chr(ob, '\n');
    }
}

int boringbits_main(int argc, char *argv[], int boring)
{
    bits_opt_t o;
    o.window_size = 2500; o.window_inc = 50;                 /* init_optp(), :543-559 */
    o.low_cov_thresh = 0.4f; o.high_cov_thresh = 2.5f; o.low_mq_cov_thresh = 0.4f;
    o.min_ctg_len = 1000000; o.edge_len = 100000; o.verbose = 4;
    const char *covmq = NULL;
    int threads = 0;                                         /* -t: the reference parses it and never uses it; here: text reader threads */
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
        else if (c == 't') {
            threads = atoi(optarg);
            if (threads < 1) { CORN_ERROR("Number of threads should larger than 0. You entered %d", threads); exit(EXIT_FAILURE); }
        }
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

    /* ---- get_depths(), :179-293: blocks of both files on several threads when the text is plain (depthtxt.c), else --
     * and for small files -- the one-pass reader that reports what the reference reports ---- */
    depth_text_t t1, t2;
    depthtxt_open(&t1, covtotal);
    depthtxt_open(&t2, covmq);
    depth_table_t T;
    {
        const char *e_min = getenv("CORNETTO_DEPTH_PAR_MIN"), *e_blk = getenv("CORNETTO_DEPTH_BLOCK");
        const size_t par_min = e_min ? (size_t)strtoull(e_min, NULL, 10) : ((size_t)8 << 20);
        const size_t block = e_blk ? (size_t)strtoull(e_blk, NULL, 10) : ((size_t)16 << 20);
        if (threads == 0) { const long nc = sysconf(_SC_NPROCESSORS_ONLN); threads = nc < 1 ? 1 : (nc > 16 ? 16 : (int)nc); }
        if (t1.size < par_min || depthtxt_load_parallel(&t1, &t2, threads, block, &T) != 0) depthtxt_load_serial(&t1, &t2, &T);
    }
    const size_t n_ctg = T.n_ctg;
    const depth_ctg_t *ctg = T.ctg;
    const uint64_t n_tot = T.n_tot;
    uint16_t *depth = T.depth, *mq = T.mq;
    const double tot_depth = T.tot_depth, tot_mq = T.tot_mq, tot_len = (double)T.n_tot;
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
        uint64_t k1 = k;
        while (k1 < w.n_win && w.win[k1].ctg == i) ++k1;
        fmt_bits_t fa = { w.win + k, name, nl };
        outbuf_format_parallel(&ob, k1 - k, fmt_bits, &fa);                       /* (threads from 200 k lines on) */
        k = k1;
    }
    outbuf_flush(&ob);
    outbuf_free(&ob);
    corn_gpu_depth_windows_free(&w);
    if (!cornetto_fast_exit()) {
        depth_table_free(&T);
        depthtxt_close(&t1); depthtxt_close(&t2);
    }
    return 0;
}
