/* cornetto_b200/host/sdust_main.c -- `cornetto sdust [-w 64] [-t 20] <in.fa|->`.
 *
 * Same contract as sdust_main(), src/sdust/sdust.c:179-207: ketopt-style short options "w:t:"
 * that may also follow the file name (permutation), "-" reads stdin, usage + exit(1) when no
 * input is given, one line  name \t start \t end  per masked interval.  sdust_core() is
 * replaced by corn_gpu_sdust() on batches of records. */
#include "cornetto.h"

typedef struct { int T, W; } sdust_arg_t;

typedef struct { const corn_intervals_t *iv; const rec_batch_t *b; } sdust_fmt_t;

/* items are intervals; the record of an interval k is the r with rec_first[r] <= k < rec_first[r + 1] */
static void sdust_format(outbuf_t *ob, uint64_t begin, uint64_t end, void *arg)
{
    const sdust_fmt_t *f = (const sdust_fmt_t *)arg;
    const uint64_t *first = f->iv->rec_first;
    if (begin >= end) return;
    uint32_t lo = 0, hi = f->iv->n_rec;              /* last r with first[r] <= begin */
    while (hi - lo > 1) { const uint32_t mid = lo + (hi - lo) / 2; if (first[mid] <= begin) lo = mid; else hi = mid; }
    uint32_t rec = lo;
    const char *name = NULL;
    size_t nl = 0;
    for (uint64_t k = begin; k < end; ++k) {
        while (k >= first[rec + 1]) { ++rec; name = NULL; }
        if (!name) { name = f->b->name[rec]; nl = strlen(name); }
        outbuf_str(ob, name, nl);
        outbuf_chr(ob, '\t'); outbuf_i32(ob, (int32_t)(f->iv->iv[k] >> 32));
        outbuf_chr(ob, '\t'); outbuf_i32(ob, (int32_t)f->iv->iv[k]);
        outbuf_chr(ob, '\n');
    }
}

static void sdust_batch(corn_ctx_t *ctx, rec_batch_t *b, outbuf_t *ob, void *arg)
{
    const sdust_arg_t *a = (const sdust_arg_t *)arg;
    corn_intervals_t iv;
    int r;
    if (b->db) r = corn_gpu_sdust_dev(ctx, b->db, a->T, a->W, &iv);       /* parsed on the device: already resident */
    else {
        corn_batch_t view;
        corn_hbatch_view(b->hb, &view);
        r = corn_gpu_sdust(ctx, &view, a->T, a->W, &iv);
    }
    if (r != CORN_OK) {
        CORN_ERROR("sdust: %s (%s)", corn_gpu_strerror(r), corn_gpu_last_error(ctx));
        exit(EXIT_FAILURE);
    }
    sdust_fmt_t f;
    f.iv = &iv; f.b = b;
    outbuf_format_parallel(ob, iv.n_iv, sdust_format, &f);
    corn_gpu_intervals_free(&iv);
}

int sdust_main(int argc, char *argv[])
{
    int W = 64, T = 20;
    const char *file = NULL;
    /* option scan equivalent to ketopt(&o, argc, argv, 1, "w:t:", 0) (src/sdust/ketopt.h:57-118):
     * arguments not starting with '-' (and a bare "-") are operands wherever they stand */
    int i = 1;
    while (i < argc) {
        const char *a = argv[i];
        if (a[0] != '-' || a[1] == 0) { if (!file) file = a; ++i; continue; }
        if (a[1] == '-') {                 /* "--" ends the options; "--long" is unknown ('?') and skipped */
            if (a[2] == 0) { for (++i; i < argc; ++i) if (!file) file = argv[i]; break; }
            ++i; continue;
        }
        int pos = 1, next = i + 1;
        while (a[pos]) {
            const int c = a[pos++];
            if (c == 'w' || c == 't') {
                const char *arg = NULL;
                if (a[pos]) arg = a + pos;
                else if (i < argc - 1) { arg = argv[i + 1]; next = i + 2; }
                if (arg) { if (c == 'w') W = atoi(arg); else T = atoi(arg); }
                break;
            }
        }
        i = next;
    }
    if (!file) {
        fprintf(stderr, "Usage: sdust [-w %d] [-t %d] <in.fa>\n", W, T);
        exit(1);
    }
    fastx_t *fx = fastx_open(file);
    if (!fx) return 0;     /* the reference reads nothing from an unopenable file and returns 0 */
    sdust_arg_t sa;
    sa.T = T; sa.W = W;
    uint64_t resume = 0;
    if (!run_ingest_pipeline(file, sdust_batch, &sa, &resume)) {
        if (resume) { fastx_close(fx); fx = fastx_open_at(file, resume); if (!fx) return 0; }
        run_batch_pipeline(fx, file, sdust_batch, &sa);
    }
    fastx_close(fx);
    return 0;
}
