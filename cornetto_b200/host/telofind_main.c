/* cornetto_b200/host/telofind_main.c -- `cornetto telofind <fasta|fastq[.gz]> [MOTIF]`.
 *
 * Same contract as find_telomere_main(), src/find_telomere.c:83-111: usage text and exit(1)
 * when the file argument is missing, motif = argv[2] or "TTAGGG" (used as given, never folded),
 * one line  name \t len \t strand \t start \t end \t end-start  per maximal tandem run, strand 0
 * runs of a record before its strand 1 runs.  The per-record toupper()/strstr() scan is replaced
 * by corn_gpu_telofind() on batches of records. */
#include "cornetto.h"

typedef struct { const corn_hits_t *hits; const rec_batch_t *b; const uint32_t *length; } telofind_fmt_t;

static void telofind_format(outbuf_t *ob, uint64_t begin, uint64_t end, void *arg)
{
    const telofind_fmt_t *f = (const telofind_fmt_t *)arg;
    for (uint64_t i = begin; i < end; ++i) {
        const corn_run_t *h = &f->hits->run[i];
        const char *name = f->b->name[h->rec];
        outbuf_str(ob, name, strlen(name));
        outbuf_chr(ob, '\t'); outbuf_u64(ob, f->length[h->rec]);
        outbuf_chr(ob, '\t'); outbuf_u64(ob, h->strand);
        outbuf_chr(ob, '\t'); outbuf_u64(ob, h->start);
        outbuf_chr(ob, '\t'); outbuf_u64(ob, h->end);
        outbuf_chr(ob, '\t'); outbuf_u64(ob, h->end - h->start);
        outbuf_chr(ob, '\n');
    }
}

static void telofind_batch(corn_ctx_t *ctx, rec_batch_t *b, outbuf_t *ob, void *arg)
{
    const char *query = (const char *)arg;
    const uint32_t *length = rec_batch_lengths(b);
    corn_hits_t hits;
    int r;
    if (b->db) r = corn_gpu_telofind_dev(ctx, b->db, query, &hits);     /* parsed on the device: already resident */
    else {
        corn_batch_t view;
        corn_hbatch_view(b->hb, &view);
        r = corn_gpu_telofind(ctx, &view, query, &hits);
    }
    if (r != CORN_OK) {
        CORN_ERROR("telofind: %s (%s)", corn_gpu_strerror(r), corn_gpu_last_error(ctx));
        exit(EXIT_FAILURE);
    }
    telofind_fmt_t f;
    f.hits = &hits; f.b = b; f.length = length;
    outbuf_format_parallel(ob, hits.n_run, telofind_format, &f);
    corn_gpu_hits_free(&hits);
}

int find_telomere_main(int argc, char *argv[])
{
    if (argc < 2) {
        fprintf(stderr, "Error: invalid number of parameters\n");
        fprintf(stderr, "Usage: find <input fasta> [optional sequence to search for, default is vertebrate TTAGGG]\n");
        exit(EXIT_FAILURE);
    }
    const char *fasta = argv[1];
    const char *query = (argc >= 3 ? argv[2] : "TTAGGG");

    fastx_t *fx = fastx_open(fasta);
    CORN_F_CHK(fx, fasta);
    if (query[0] == 0) {
        /* the reference spins forever on an empty motif (strstr always matches); refuse instead */
        CORN_ERROR("%s", "empty motif");
        exit(EXIT_FAILURE);
    }
    uint64_t resume = 0;
    if (!run_ingest_pipeline(fasta, telofind_batch, (void *)query, &resume)) {
        if (resume) { fastx_close(fx); fx = fastx_open_at(fasta, resume); CORN_F_CHK(fx, fasta); }
        run_batch_pipeline(fx, fasta, telofind_batch, (void *)query);
    }
    fastx_close(fx);
    return EXIT_SUCCESS;
}

/* `cornetto fa2bed <fasta>`: name \t 0 \t len per record (src/assbed.c:97-100); a by-product of
 * the reader, needed by scripts/telostats.sh:36 for the .lens file. */
int assbed_main(int argc, char *argv[])
{
    if (argc != 2 || strcmp(argv[1], "-h") == 0 || strcmp(argv[1], "--help") == 0) {
        FILE *fp = (argc == 2) ? stdout : stderr;
        fprintf(fp, "Usage: cornetto fa2bed <assembly.fa>\n");
        exit(fp == stdout ? EXIT_SUCCESS : EXIT_FAILURE);
    }
    fastx_t *fx = fastx_open(argv[1]);
    CORN_F_CHK(fx, argv[1]);
    uint8_t *scratch = (uint8_t *)malloc(1 << 20);
    CORN_MALLOC_CHK(scratch);
    while (fastx_next(fx)) {
        char *name = strdup(fastx_name(fx));
        uint64_t len = 0;
        int done = 0;
        while (!done) len += fastx_seq(fx, scratch, 1 << 20, &done);
        if (fastx_finish(fx) != 0) { free(name); break; }
        fprintf(stdout, "%s\t%d\t%d\n", name, 0, (int)len);
        free(name);
    }
    free(scratch);
    fastx_close(fx);
    return 0;
}
