/* include/corn_gpu.h -- C ABI of the B200-native cornetto sequence-scan path.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no C++/torch types.
 * The reference (hasindu2008/cornetto, C99) has no device boundary at all -- its four hot
 * sub-commands do their arithmetic inline in
 *     find_telomere_main      src/find_telomere.c:83-111   (disambiguate :76-81, find :44-74, rc :24-42)
 *     telomere_windows_main   src/telomere_windows.c:45-86 (process_scaffold :28-43)
 *     telomere_breaks_main    src/telomere_breaks.c:47-172
 *     sdust_main / sdust()    src/sdust/sdust.c:162-207    (sdust_core :130-160), API src/sdust/sdust.h:16-21
 * A maintainer replaces those inline loops by the calls below (INTEGRATION.md shows the patch);
 * cornetto_b200/host/ is that host side written out in C with the CLI and text formats unchanged.
 *
 * Conventions
 *   - every function returns CORN_OK (0) or a negative corn_status; nothing here prints or exits
 *     (the host wrappers own stdout/stderr/exit so the reference's text stays byte-identical);
 *   - there is NO CPU fallback: without a usable sm_100 device corn_gpu_init() fails with
 *     CORN_E_NOGPU and every other entry point needs a context;
 *   - result arrays are owned by the library (pinned host memory) until the matching *_free;
 *   - one context per GPU; a context may be used by one host thread at a time.
 */
#ifndef CORN_GPU_H
#define CORN_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CORN_GPU_ABI_VERSION 1

/* HBM sequence layout: records are concatenated; every record starts at a multiple of
 * CORN_ALIGN bytes, is followed by at least one 0x00 byte, and all bytes outside records are
 * 0x00.  (0x00 never matches a motif byte, so no occurrence can span two records.) */
#define CORN_ALIGN 32u
/* Largest batch (bytes incl. padding): hit positions are 32-bit inside the kernels. */
#define CORN_MAX_BATCH_BYTES 0xFFF00000ull

typedef enum corn_status {
    CORN_OK          =  0,
    CORN_E_NOGPU     = -1,  /* no CUDA device / not sm_100 / driver missing */
    CORN_E_CUDA      = -2,  /* a CUDA runtime call failed; see corn_gpu_last_error() */
    CORN_E_ARG       = -3,  /* invalid argument (NULL, empty motif, W out of range, ...) */
    CORN_E_NOMEM     = -4,  /* host or device allocation failed */
    CORN_E_LAYOUT    = -5,  /* batch violates the layout contract above */
    CORN_E_TOOBIG    = -6,  /* batch larger than CORN_MAX_BATCH_BYTES or record >= 2^31 */
    CORN_E_STATE     = -7,  /* call sequence error (e.g. telowin(NULL hits) without a prior telofind) */
    CORN_E_INTERNAL  = -8   /* device-side consistency check failed */
} corn_status;

typedef struct corn_ctx corn_ctx_t;

/* ---- context ------------------------------------------------------------------------------ */
int         corn_gpu_device_count(void);                 /* >=0, or negative corn_status */
/* device < 0: take $CORNETTO_GPU if set, else device 0. */
int         corn_gpu_init(int device, corn_ctx_t **ctx);
void        corn_gpu_destroy(corn_ctx_t *ctx);
const char *corn_gpu_strerror(int status);
const char *corn_gpu_last_error(const corn_ctx_t *ctx);  /* detail of the last failure ("" if none) */
/* Run all work of this context on an existing CUDA stream (cudaStream_t passed as void*).
 * NULL restores the context's own stream. */
int         corn_gpu_set_stream(corn_ctx_t *ctx, void *cuda_stream);

/* ---- host batches ------------------------------------------------------------------------- */
typedef struct corn_batch {
    const uint8_t  *seq;          /* total_bytes bytes in the layout described above */
    const uint64_t *offset;       /* [n_rec] start of each record in seq (multiple of CORN_ALIGN) */
    const uint32_t *length;       /* [n_rec] record length in bytes (kseq's seq.l) */
    uint32_t        n_rec;
    uint64_t        total_bytes;  /* multiple of CORN_ALIGN, <= CORN_MAX_BATCH_BYTES */
} corn_batch_t;

/* Builder for a host batch (what the kseq-style reader fills directly).  Replaces the per-record
 * kstring_t buffer that kseq_read() grows with realloc (src/kseq.h:201-211).
 * corn_hbatch_create() touches no CUDA API, so parsing can start while the driver is still
 * initialising on another thread; corn_hbatch_pin() page-locks the buffer (cudaHostRegister, once;
 * needs a device) so that the H2D copy runs at full PCIe rate.  Without it the library copies
 * through its own small page-locked ring (~25 GB/s), which is the better deal for a buffer that is
 * filled only once or twice: page-locking costs 0.1-0.4 s per GB and as much again to release. */
typedef struct corn_hbatch corn_hbatch_t;
int      corn_hbatch_create(uint64_t capacity_bytes, uint32_t max_records, corn_hbatch_t **hb);
int      corn_hbatch_pin(corn_hbatch_t *hb);
void     corn_hbatch_destroy(corn_hbatch_t *hb);
void     corn_hbatch_reset(corn_hbatch_t *hb);
/* Space left for ONE more record (bytes of sequence), 0 if the record table is full. */
uint64_t corn_hbatch_room(const corn_hbatch_t *hb);
/* Zero-copy fill: returns where the next record's bytes go (up to corn_hbatch_room()). */
uint8_t *corn_hbatch_cursor(corn_hbatch_t *hb);
/* Seals the record written at the cursor (length bytes), zero-fills its padding. */
int      corn_hbatch_commit(corn_hbatch_t *hb, uint64_t length);
/* Convenience: copy a record in.  CORN_E_TOOBIG if it does not fit. */
int      corn_hbatch_add(corn_hbatch_t *hb, const void *bases, uint64_t length);
void     corn_hbatch_view(const corn_hbatch_t *hb, corn_batch_t *view);

/* ---- device-resident batches -------------------------------------------------------------- */
typedef struct corn_dbatch corn_dbatch_t;
int  corn_gpu_upload(corn_ctx_t *ctx, const corn_batch_t *batch, corn_dbatch_t **db);
void corn_gpu_dbatch_free(corn_ctx_t *ctx, corn_dbatch_t *db);
/* Device address of the sequence bytes (for in-place synthetic generation in bench.py). */
void *corn_gpu_dbatch_seq_ptr(const corn_dbatch_t *db);
uint64_t corn_gpu_dbatch_bytes(const corn_dbatch_t *db);
/* Allocates a device batch with the given record lengths WITHOUT uploading bytes: the
 * sequence area is zeroed and the caller (bench synthetic generator) fills the records. */
int  corn_gpu_dbatch_alloc(corn_ctx_t *ctx, const uint32_t *length, uint32_t n_rec, corn_dbatch_t **db);
int  corn_gpu_dbatch_download(corn_ctx_t *ctx, const corn_dbatch_t *db, uint32_t rec, uint8_t *dst);

/* ---- telofind: replaces disambiguate()+find(), src/find_telomere.c:44-81 ------------------- */
typedef struct corn_run {       /* one printed line: name \t len \t strand \t start \t end \t end-start */
    uint32_t rec;               /* record index inside the batch */
    uint32_t strand;            /* 0 = motif, 1 = rc(motif)  (src/find_telomere.c:51,65) */
    uint32_t start, end;        /* [start,end) on the record */
} corn_run_t;

typedef struct corn_hits {
    corn_run_t *run;            /* in the reference's print order: per record, strand 0 runs then strand 1 */
    uint64_t    n_run;
    void       *_owner;
} corn_hits_t;

/* Host-buffer entry point (H2D + kernels + D2H). */
/* motif: 1..255 bytes, used as given (never case-folded, src/find_telomere.c:90-91); longer motifs are refused with
 * CORN_E_ARG (the reference has no bound; nothing biological comes near it). */
int  corn_gpu_telofind(corn_ctx_t *ctx, const corn_batch_t *batch, const char *motif, corn_hits_t *out);
/* Same on a resident batch.  out may be NULL: the runs then stay on the device only (they are
 * always kept there for a following corn_gpu_telowin(hits == NULL)). */
int  corn_gpu_telofind_dev(corn_ctx_t *ctx, const corn_dbatch_t *db, const char *motif, corn_hits_t *out);
void corn_gpu_hits_free(corn_hits_t *hits);

/* ---- telowin: replaces the paint loop + process_scaffold(), src/telomere_windows.c:28-43,75-79 */
typedef struct corn_contigs {
    const uint32_t *length;     /* [n] contig lengths (atoi(col2), src/telomere_windows.c:72) */
    uint32_t        n;
} corn_contigs_t;

typedef struct corn_window {    /* "Window\tname\tlen\tstart\tend\t%.3g(car/(end-start))" */
    uint32_t rec;
    uint32_t start, end;        /* i, i+den */
    uint32_t car;               /* marked bases in [start,end) */
} corn_window_t;

typedef struct corn_windows {
    corn_window_t *win;
    uint64_t       n_win;
    void          *_owner;
} corn_windows_t;

/* threshold_adj = thr * pow(identity/100, 6), computed by the caller in double exactly as
 * src/telomere_windows.c:53-54 does.
 * hits == NULL : use the runs left on the device by the last corn_gpu_telofind*() of this
 *                context (fused telofind+telowin; contigs may be NULL = the batch's records).
 * hits != NULL : paint these runs (any order, overlaps allowed; rec indexes contigs[]). */
int  corn_gpu_telowin(corn_ctx_t *ctx, const corn_hits_t *hits, const corn_contigs_t *contigs,
                      double threshold_adj, corn_windows_t *out);
void corn_gpu_windows_free(corn_windows_t *w);

/* ---- sdust: replaces sdust_core(), src/sdust/sdust.c:130-160 ------------------------------- */
typedef struct corn_intervals {
    uint64_t *iv;               /* start<<32 | finish, exactly the reference's encoding (sdust.c:99) */
    uint64_t *rec_first;        /* [n_rec+1] iv[rec_first[r] .. rec_first[r+1]) belong to record r */
    uint64_t  n_iv;
    uint32_t  n_rec;
    void     *_owner;
} corn_intervals_t;

/* Any T (default 20).  W (default 64) as the reference takes it (src/sdust/sdust.c:186-189, W = atoi()):
 *   3 <= W <= 128    the tuned kernels (csrc/sdust.cu);
 *   128 < W <= 1024  a generic instance (csrc/sdust_wide.cu), same results as the reference, not tuned;
 *   W < 3            no interval, CORN_OK (the reference dereferences an empty deque there and crashes
 *                    on the first triplet without printing anything);
 *   W > 1024         CORN_E_ARG: the reference's own 32-bit score products (r * l, :113-117) overflow inside
 *                    long low-complexity runs, so there is no defined result to reproduce. */
int  corn_gpu_sdust(corn_ctx_t *ctx, const corn_batch_t *batch, int T, int W, corn_intervals_t *out);
int  corn_gpu_sdust_dev(corn_ctx_t *ctx, const corn_dbatch_t *db, int T, int W, corn_intervals_t *out);
void corn_gpu_intervals_free(corn_intervals_t *iv);

/* The reference's own library interface for sdust, src/sdust/sdust.h:16-21 -- same names, same signatures,
 * same ownership (csrc/sdust_api.cu): a program that links the reference's sdust.o can link this library
 * instead.  sdust(): result malloc()ed, caller free()s, *n intervals, l_seq < 0 means strlen(seq), `km` is
 * ignored (as by the reference's kalloc.h).  sdust_core(): result owned by `buf`, valid until the next call on
 * it.  Uses one process-wide context on device $CORNETTO_GPU (default 0); on failure NULL and *n = -1. */
struct sdust_buf_s;
typedef struct sdust_buf_s sdust_buf_t;
uint64_t       *sdust(void *km, const uint8_t *seq, int l_seq, int T, int W, int *n);
sdust_buf_t    *sdust_buf_init(void *km);
void            sdust_buf_destroy(sdust_buf_t *buf);
const uint64_t *sdust_core(const uint8_t *seq, int l_seq, int T, int W, int *n, sdust_buf_t *buf);

/* ---- ingest: replaces the byte loop of kseq_read()/ks_getuntil2(), src/kseq.h:102-141,184-224 ---
 * Parses plain FASTA/FASTQ TEXT on the device: ships the raw file bytes over PCIe once, finds the
 * line structure, builds the record table and compacts the sequence bytes into the resident
 * CORN_ALIGN layout that the *_dev entry points take.  `text` must start at a record boundary:
 * its first byte is the '>' or '@' of a header (the start of the file, or where the previous call
 * stopped: text + consumed).  `final` != 0 says the text ends at the end of the input, so the last
 * record is complete; otherwise the trailing incomplete record is left to the next call.
 *
 * Only REGULAR text is parsed on the device -- the subset on which the result provably equals
 * kseq_read()'s (cornetto_b200/csrc/ingest_core.cuh states the rules and why):
 *   FASTA : every record is a '>' line followed by sequence lines, none of which starts with '@' or
 *           '+' or consists of a lone '\r';
 *   FASTQ : four lines per record, '@...' / sequence / '+...' / quality of the same length.
 * Anything else (multi-line FASTQ, junk before the first header, NUL bytes, truncated quality, ...)
 * sets out->irregular = 1 and nothing else; the caller then parses that text with its serial
 * kseq-equivalent reader (cornetto_b200/host/fastx.c) and uploads batches as before.  That is a
 * choice of FEEDER only: the scans themselves always run on the GPU. */
typedef struct corn_ingest {
    corn_dbatch_t *db;          /* resident batch of the n_rec complete records (NULL when n_rec == 0);
                                   release with corn_gpu_dbatch_free() */
    uint32_t   n_rec;
    uint64_t  *hdr_off;         /* [n_rec] offset in text of each record's '>' / '@' (name starts one byte later,
                                   ends before the first isspace() byte, src/kseq.h:195) */
    uint32_t  *length;          /* [n_rec] sequence length (kseq's seq.l) */
    uint64_t   consumed;        /* bytes of text covered by these records; the next call starts at text + consumed */
    int        irregular;
    void      *_owner;
} corn_ingest_t;

int  corn_gpu_ingest(corn_ctx_t *ctx, const uint8_t *text, uint64_t n_text, int final, corn_ingest_t *out);
/* frees hdr_off / length (NOT out->db) */
void corn_gpu_ingest_free(corn_ingest_t *ing);
/* page-locks a caller-owned text buffer so that corn_gpu_ingest() copies it at full PCIe rate */
int  corn_gpu_host_register(void *p, uint64_t bytes);
void corn_gpu_host_unregister(void *p);

/* ---- depth windows: replaces get_regs() and the selection loops of print_fun_bits() / print_boring_bits(),
 *      src/boringbits_main.c:315-372,425-485 (`cornetto noboringbits`, `boringbits`) ---------------------------
 * depth / mq_depth: one uint16 per base (get_depths(), :179-293), contigs concatenated; offset[c] = first element of
 * contig c.  Windows of window_size every window_inc bases, the last ones clipped to the contig (:340-347); per window
 * the integer means of both arrays (C division, :357-358).  A window is returned when the reference prints it:
 *   boring == 0  contigs of at least min_ctg_len:  depth < thresh_low_depth || depth > thresh_high_depth ||
 *                mq_depth / (double)depth < low_mq_cov_thresh                                   (:436-441)
 *   boring != 0  contigs longer than min_ctg_len, st > edge_len && end < length - edge_len, and none of the above  (:468-481)
 * in (contig, start) order.  thresh_*_depth are round(factor * mean depth), computed by the caller as :524-525 does. */
typedef struct corn_depth_batch {
    const uint16_t *depth, *mq_depth;
    const uint64_t *offset;     /* [n_ctg] */
    const uint32_t *length;     /* [n_ctg], 1 .. 2^31-1 */
    uint32_t        n_ctg;
    uint64_t        n_total;    /* elements in depth / mq_depth */
} corn_depth_batch_t;

typedef struct corn_depth_params {
    int   window_size, window_inc;              /* both >= 1 */
    int   thresh_low_depth, thresh_high_depth;
    float low_mq_cov_thresh;
    int   edge_len, min_ctg_len;
    int   boring;
} corn_depth_params_t;

typedef struct corn_depth_window { uint32_t ctg, st, end; int32_t depth, mq_depth; } corn_depth_window_t;

typedef struct corn_depth_windows {
    corn_depth_window_t *win;
    uint64_t             n_win;
    void                *_owner;
} corn_depth_windows_t;

int  corn_gpu_depthwin(corn_ctx_t *ctx, const corn_depth_batch_t *batch, const corn_depth_params_t *params, corn_depth_windows_t *out);
void corn_gpu_depth_windows_free(corn_depth_windows_t *w);

/* ---- several GPUs: record sharding (csrc/shard.cu; host arithmetic only, usable without a device) ----------
 * The reference is one thread in one address space (records scanned and printed in file order,
 * src/find_telomere.c:101-105).  Here the unit of distribution is the record: no scan looks across a record
 * boundary, so shards are independent -- no halo, no collective -- and results only have to be put back into
 * file order.  Records are never cut (a single-contig input stays on one GPU).
 *   corn_shard_plan         shard_of[r] for every record: longest first onto the least loaded shard
 *                           (deterministic: every process of a job derives the same plan from the lengths);
 *   corn_shard_local_index  local[r] = index of record r inside its shard's batch (file order is kept inside
 *                           a shard), count[s] = records of shard s;
 *   corn_shard_merge_runs / corn_shard_merge_intervals
 *                           per-shard results (rec = index inside the shard) -> one list in file order, i.e.
 *                           exactly what a single corn_gpu_telofind / corn_gpu_sdust over all records returns.
 *                           out must hold the sum of the per-shard counts (out_first: n_rec + 1). */
int corn_shard_plan(const uint32_t *length, uint32_t n_rec, uint32_t n_shards, uint32_t *shard_of);
int corn_shard_local_index(const uint32_t *shard_of, uint32_t n_rec, uint32_t n_shards, uint32_t *local, uint32_t *count);
int corn_shard_merge_runs(const corn_run_t *const *runs, const uint64_t *n_run, const uint32_t *shard_of,
                          uint32_t n_rec, uint32_t n_shards, corn_run_t *out);
int corn_shard_merge_intervals(const uint64_t *const *iv, const uint64_t *const *rec_first, const uint32_t *shard_of,
                               uint32_t n_rec, uint32_t n_shards, uint64_t *out_iv, uint64_t *out_first);

/* ---- measurement hooks (CUDA events on the context's stream; bench.py reads them) ---------- */
typedef struct corn_timing {
    float h2d_ms;       /* host->device copies of the last call */
    float scan_ms;      /* dominant kernel: telofind_scan / sdust_scan */
    float post_ms;      /* every other kernel of the call (ordering, run assembly, bins, windows) */
    float d2h_ms;       /* device->host copies of the results */
    uint32_t launches;  /* kernels launched by the last call */
    uint64_t out_bytes; /* result bytes produced on the device */
} corn_timing_t;
int  corn_gpu_last_timing(const corn_ctx_t *ctx, corn_timing_t *t);
uint64_t corn_gpu_total_launches(const corn_ctx_t *ctx);   /* since corn_gpu_init */

#ifdef __cplusplus
}
#endif
#endif /* CORN_GPU_H */
