/* cornetto_b200/host/cornetto.h -- shared declarations of the drop-in `cornetto` host program.
 *
 * The host side keeps the reference's operator interface for the hot path: the same four
 * `int xxx_main(int argc, char *argv[])` entry points (src/main.c:45-48,111-122), the same
 * arguments, stdout text, stderr messages and exit codes -- with the inner loops replaced by
 * calls into the CUDA library (include/corn_gpu.h). */
#ifndef CORNETTO_HOST_H
#define CORNETTO_HOST_H

#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "corn_gpu.h"

#define CORNETTO_VERSION "0.2.0"

/* message style of the reference (src/error.h:54-119) */
#define CORN_ERROR(msg, ...) \
    fprintf(stderr, "[%s::ERROR]\033[1;31m " msg "\033[0m At %s:%d\n", __func__, __VA_ARGS__, __FILE__, __LINE__ - 1)
#define CORN_F_CHK(ret, file) \
    do { if ((ret) == NULL) { CORN_ERROR("Could not to open file %s: %s", file, strerror(errno)); exit(EXIT_FAILURE); } } while (0)
#define CORN_MALLOC_CHK(ret) \
    do { if ((ret) == NULL) { CORN_ERROR("Failed to allocate memory: %s", strerror(errno)); exit(EXIT_FAILURE); } } while (0)

int find_telomere_main(int argc, char *argv[]);
int telomere_windows_main(int argc, char *argv[]);
int telomere_breaks_main(int argc, char *argv[]);
int sdust_main(int argc, char *argv[]);
int assbed_main(int argc, char *argv[]);
int nx_main(int argc, char *argv[]);
int report_main(int argc, char *argv[]);
int seq_main(int argc, char *argv[]);
int telostats_main(int argc, char *argv[]);
int boringbits_main(int argc, char *argv[], int boring);

/* misc.c */
uint64_t cornetto_batch_capacity(const char *path, int n_parts);
double realtime(void);
double cputime(void);
long   peakrss(void);

/* GPU context shared by the sub-commands: created on first use; exits with the reference's
 * ERROR style when no device is usable (there is no CPU fallback). */
corn_ctx_t *cornetto_gpu(void);
/* starts creating that context on a helper thread (cornetto_gpu() then waits for it); failures surface in cornetto_gpu() */
void        cornetto_gpu_prefetch(void);
void        cornetto_gpu_release(void);
/* prints the library error and exits */
void        cornetto_gpu_die(const char *what, int status);
/* 1 (default): the process ends with _exit() once its output is flushed and skips freeing GPU
 * contexts and large buffers; $CORNETTO_FAST_EXIT=0 restores the orderly teardown (leak checkers). */
int         cornetto_fast_exit(void);

/* ---- FASTA/FASTQ reader with kseq_read() semantics (src/kseq.h:184-224) -------------------- */
typedef struct fastx fastx_t;
fastx_t *fastx_open(const char *path);        /* "-" = stdin; plain or gzip; NULL if it cannot be opened */
/* plain (uncompressed, seekable) files only: starts reading at a byte offset that is a record boundary */
fastx_t *fastx_open_at(const char *path, uint64_t offset);
void     fastx_close(fastx_t *fx);
/* Advances to the next record header.  1 = a record begins (fastx_name() valid), 0 = end of input. */
int      fastx_next(fastx_t *fx);
const char *fastx_name(const fastx_t *fx);
/* Appends sequence bytes of the current record to dst (at most cap).  Returns the number of
 * bytes written; *done becomes 1 once the sequence part of the record is complete. */
size_t   fastx_seq(fastx_t *fx, uint8_t *dst, size_t cap, int *done);
/* After the sequence is complete: consumes a FASTQ quality block if there is one.
 * 0 = record valid; -2 = truncated/mismatching quality (the reference stops reading there). */
int      fastx_finish(fastx_t *fx);

/* ---- record batches: names + pinned sequence bytes in the CORN_ALIGN layout ------------------ */
typedef struct {
    corn_hbatch_t *hb;    /* host batch filled by the serial reader (NULL for device-parsed batches) */
    char   **name;        /* [n] strdup'ed (serial reader) or pointers into name_arena (device-parsed) */
    uint32_t n, max_rec;
    int      eof;         /* input exhausted (or stopped at a malformed FASTQ record) */
    uint64_t seq;         /* 0, 1, 2, ... in input order (set by the pipelines before the callback runs) */
    /* device-parsed batches (ingest.c): the records are already resident */
    corn_dbatch_t  *db;
    const uint32_t *length;      /* [n] */
    char           *name_arena;
} rec_batch_t;

/* record lengths of a batch, whichever way it was built */
static inline const uint32_t *rec_batch_lengths(const rec_batch_t *b)
{
    if (b->db) return b->length;
    corn_batch_t v;
    corn_hbatch_view(b->hb, &v);
    return v.length;
}

rec_batch_t *rec_batch_create(uint64_t capacity_bytes, uint32_t max_rec);
void         rec_batch_destroy(rec_batch_t *b);
/* Fills the batch with as many whole records as fit.  A record that does not fit into an EMPTY
 * batch makes the batch grow.  Returns the number of records now in the batch. */
uint32_t     rec_batch_fill(rec_batch_t *b, fastx_t *fx);

/* ---- buffered text output (fp == NULL: grows in memory, written out later with outbuf_write) -- */
typedef struct { char *buf; size_t n, cap; FILE *fp; } outbuf_t;
void outbuf_init(outbuf_t *o, FILE *fp);
void outbuf_write(outbuf_t *o, FILE *fp);     /* memory buffer -> fp, then empty it */
void outbuf_flush(outbuf_t *o);
void outbuf_free(outbuf_t *o);
void outbuf_str(outbuf_t *o, const char *s, size_t len);
void outbuf_u64(outbuf_t *o, uint64_t v);
void outbuf_i32(outbuf_t *o, int32_t v);
void outbuf_chr(outbuf_t *o, char c);
/* formats items [0, n) with fn(ob, begin, end, arg) -- on several threads when there are many -- and appends
 * the text to ob in item order */
typedef void (*format_range_fn)(outbuf_t *ob, uint64_t begin, uint64_t end, void *arg);
void outbuf_format_parallel(outbuf_t *ob, uint64_t n_items, format_range_fn fn, void *arg);

/* ---- batch pipeline: one parser thread, one worker thread per GPU context --------------------
 * Records are parsed into pinned batches by the calling thread while worker threads run the GPU
 * call and format the text of earlier batches; output is written strictly in batch order, so the
 * result is byte-identical to the sequential loop of the reference (src/find_telomere.c:101-105).
 * $CORNETTO_GPUS = number of devices to use (default 1; batches go round-robin over them). */
typedef void (*batch_fn)(corn_ctx_t *ctx, rec_batch_t *b, outbuf_t *out, void *arg);
void run_batch_pipeline(fastx_t *fx, const char *path, batch_fn fn, void *arg);
/* where the pipelines write the batches' text, in batch order (stdout unless set; telostats writes a file) */
void  cornetto_set_pipeline_out(FILE *fp);
FILE *cornetto_pipeline_out(void);
/* next batch number of this process (batches are numbered in input order across both pipelines) */
uint64_t cornetto_next_batch_seq(void);

/* ---- device-side parsing for plain files (ingest.c) --------------------------------------------
 * Reads the file in large blocks, hands each block to corn_gpu_ingest() and calls fn with a
 * device-parsed batch (b->db set).  Returns 1 when the whole input was handled this way.  Returns 0
 * when the input is not eligible (stdin, gzip, $CORNETTO_INGEST=0) or turned out to be irregular
 * text: everything before *resume has been processed and written, and the caller continues with the
 * serial reader from byte offset *resume. */
int run_ingest_pipeline(const char *path, batch_fn fn, void *arg, uint64_t *resume);
/* ---- decompressed text of a .gz input in large blocks (gzsrc.c): BGZF members inflated on several threads, any other
 * gzip file by one stream beside the GPU work.  gzsrc_read: up to n bytes at dst, *eof once the input has ended; -1 if
 * the file is damaged. */
typedef struct gzsrc gzsrc_t;
gzsrc_t *gzsrc_open(const char *path);
void     gzsrc_close(gzsrc_t *g);
int      gzsrc_is_bgzf(const gzsrc_t *g);
int64_t  gzsrc_read(gzsrc_t *g, uint8_t *dst, uint64_t n, int *eof);

/* ---- the two per-base depth tables of noboringbits / boringbits (depthtxt.c; get_depths(), src/boringbits_main.c:179-293).
 * depthtxt_load_parallel returns -1, with nothing printed, for any input the serial reader would not accept silently;
 * the caller then runs depthtxt_load_serial, which reports and exits as the reference does. */
typedef struct { const char *p, *e; void *map; size_t size; } depth_text_t;
typedef struct { char *name; uint64_t off; uint32_t len; } depth_ctg_t;
typedef struct { depth_ctg_t *ctg; size_t n_ctg; uint16_t *depth, *mq; uint64_t n_tot; double tot_depth, tot_mq; } depth_table_t;
void depthtxt_open(depth_text_t *t, const char *path);
void depthtxt_close(depth_text_t *t);
void depthtxt_load_serial(depth_text_t *t1, depth_text_t *t2, depth_table_t *T);
int  depthtxt_load_parallel(const depth_text_t *t1, const depth_text_t *t2, int threads, size_t block_bytes, depth_table_t *T);
void depth_table_free(depth_table_t *T);

/* length of a record name starting at p: up to the first isspace() byte or max (src/kseq.h:195) */
size_t cornetto_name_len(const uint8_t *p, size_t max);

/* ---- khash iteration order (src/khash.h:230-348,395-400), for telobreaks' output order -------- */
size_t khash_str_order(const char *const *names, size_t n, size_t *order);

#endif
