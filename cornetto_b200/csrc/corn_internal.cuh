// cornetto_b200/csrc/corn_internal.cuh -- context, device buffers and small device utilities
// shared by the telofind / telowin / sdust translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/corn_gpu.h"

// --------------------------------------------------------------------------------------------
// HBM layout of a resident batch
//   [ GUARD bytes of 0x00 ][ total_bytes of records+padding ][ TAIL bytes of 0x00 ]
// GUARD lets the kernels look m bytes behind position 0, TAIL lets them read one tile row and
// the sdust right halo past the end without bounds checks.
// --------------------------------------------------------------------------------------------
#define CORN_GUARD_BYTES 1024u
#define CORN_TAIL_BYTES  4096u

// telofind scan tiling: a chunk is the 32 bytes one lane owns per warp iteration, a row is the
// 1 KiB a warp consumes per iteration, a tile is what a warp scans between two flushes of its
// candidate list.
#define CORN_CHUNK_BYTES  32u
#define CORN_ROW_BYTES    1024u
#define CORN_TILE_ROWS    32u
#define CORN_TILE_CHUNKS  (CORN_TILE_ROWS * 32u)            // 1024 candidate slots per tile
#define CORN_TILE_BYTES   (CORN_TILE_ROWS * CORN_ROW_BYTES) // 32 KiB

struct corn_dbatch {
    uint8_t  *d_base;      // allocation start (guard included)
    uint8_t  *d_seq;       // d_base + CORN_GUARD_BYTES
    uint64_t  total_bytes; // multiple of CORN_ALIGN
    uint64_t  alloc_bytes; // size of the allocation (may exceed span_bytes when a retired buffer is reused)
    uint64_t  span_bytes;  // guard + whole tiles + tail actually used by this batch
    uint32_t  n_rec;
    uint32_t *d_rec_off;   // [n_rec+1] start of each record (buffer position); [n_rec] = total_bytes
    uint32_t *d_rec_len;   // [n_rec]
    uint32_t *h_rec_off;   // host copies
    uint32_t *h_rec_len;
    uint64_t  n_bases;     // sum of lengths
    uint32_t *d_bin_base;  // [n_rec+1] telowin: first 200-bp bin of each record (corn_nbins_of bins per record)
    uint32_t *h_bin_base;
    uint64_t  n_bins_total;
};

// telowin geometry: a record owns ceil(len/200) bins plus four zero bins so that a window can always read five
static inline __host__ __device__ uint32_t corn_nbins_of(uint32_t len) { return (len + 199u) / 200u + 4u; }
// A full window (1000 bases = five bins) that holds car marked bases has a bin with at least car / 5 of them.  The run
// assembly lists the bins that reach CORN_HOT_BIN; whenever the threshold asks for car >= 5 * CORN_HOT_BIN (any
// threshold_adj >= 0.2: the default is 0.4 * 0.994), every passing full window contains a listed bin, and telowin only
// has to look at the five windows around each of them (a few thousand per genome) instead of at every window.
#define CORN_HOT_BIN 40u
#define CORN_HOT_CAP (1u << 20)
// layout of the first 64 bytes of ctx->misc (one small readback fetches them all):
//   [0,16) telofind totals  [16] tile ticket counter  [20] telofind error counter  [24] dense-tile queue length
//   [32,48) telowin counters ([1] bitmap words, [2] windows written, [3] hot blocks)  [48] hot bins listed
#define CORN_MISC_HOT_COUNT 48

struct corn_dbuf {         // grow-only device scratch
    void  *p;
    size_t cap;
};

#define CORN_STAGE_THREADS 16     /* upper bound; ctx->stage_threads are used */
#define CORN_STAGE_SLOTS   2
#define CORN_STAGE_BYTES   (8u << 20)

struct corn_ctx {
    int          device;
    int          sm_count;
    cudaStream_t own_stream, stream;
    cudaStream_t aux_stream;   // second stream for kernels that may share the GPU with the main one (sdust: dense and sparse items), created on first use
    cudaEvent_t  aux_ev[2];
    cudaEvent_t  ev[16];  // 0-1 upload, 2-5 telofind, 8-11 telowin, 2-6 sdust
    char         err[512];
    corn_timing_t timing;
    uint64_t     total_launches;

    // scratch (grow-only, reused across calls)
    corn_dbuf cand;        // telofind tile candidate lists
    corn_dbuf tile_tab;    // per-tile counts / offsets
    corn_dbuf events;      // ordered start/end position lists
    corn_dbuf runs;        // final corn_run_t list of the last telofind
    corn_dbuf misc;        // small per-call things (counters, totals, scan temporaries)
    corn_dbuf scan_tmp;
    corn_dbuf bins;        // telowin bin counters
    corn_dbuf bitmap;      // telowin general path
    corn_dbuf wins;        // telowin output
    corn_dbuf sd_slots;    // sdust per-chunk interval slots
    corn_dbuf sd_out;      // sdust compacted output
    corn_dbuf sd_tab;      // sdust chunk tables
    corn_dbuf ing_text;    // ingest: raw text block
    corn_dbuf ing_tab;     // ingest: per-tile newline counts / bases
    corn_dbuf ing_lines;   // ingest: line tables
    corn_dbuf ing_rec;     // ingest: record table
    corn_dbuf ranks;       // telofind: per-record run ranks
    corn_dbuf hot;         // telofind -> telowin: bins that reached CORN_HOT_BIN marked bases

    // one retired sequence buffer kept for the next upload (cudaMalloc/cudaFree of GBs cost milliseconds)
    uint8_t *spare_base;
    uint64_t spare_bytes;

    // batch uploaded by a host-buffer entry point; kept resident until the next such call so a
    // fused telowin(hits == NULL) can still reach its record table
    corn_dbatch *owned_db;

    // state left by the last telofind for a fused telowin(hits == NULL)
    // (telofind_dev(out == NULL) returns without a host sync: `pending` marks results whose totals
    // and consistency flags have not been looked at yet; the next sync point resolves them)
    int       pending;
    char      cached_motif[256];   // motif whose byte patterns are already in ctx->misc (skips a tiny H2D per call)
    char      pending_motif[256];
    uint32_t  pending_ev_cap, pending_run_cap;
    const corn_dbatch *last_db;
    uint64_t  last_n_run;
    uint32_t  last_n_win;          // windows returned by the previous telowin: sizes the speculative D2H copy
    int       last_runs_disjoint;  // runs cannot overlap (border-free motif, no fwd/rev overlap)
    int       last_motif_len;
    // 200-bp bin counts filled by the run assembly of the last telofind (disjoint runs only), with the list of bins
    // that reached CORN_HOT_BIN: a fused telowin(hits == NULL) then only looks at windows around those
    const corn_dbatch *bins_for_db;    // NULL: ctx->bins does not hold the counts of last_db's runs
    int       counters_clean;          // the tile / dense-queue counters in ctx->misc were left zeroed by the last telofind

    void     *h_pinned_small;      // 4 KiB pinned scratch for small readbacks

    // staging ring for host->device copies from pageable memory (corn_h2d)
    uint8_t     *stage;            // stage_threads * CORN_STAGE_SLOTS slots of CORN_STAGE_BYTES, page-locked
    int          stage_threads;    // host threads filling the ring (4; $CORNETTO_STAGE_THREADS)
    cudaStream_t stage_stream[CORN_STAGE_THREADS];
    cudaEvent_t  stage_ev[CORN_STAGE_THREADS * CORN_STAGE_SLOTS];
};


// Host->device copy ordered on ctx->stream.  Page-locked sources are copied directly; pageable ones
// go through a small page-locked ring filled by ctx->stage_threads host threads (page-locking a
// multi-GB buffer costs 0.1-0.4 s/GB and as much again to release: more than the copy itself).
int corn_h2d(corn_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);

// --------------------------------------------------------------------------------------------
// error plumbing
// --------------------------------------------------------------------------------------------
int corn_set_err(corn_ctx *ctx, int code, const char *fmt, ...);

#define CORN_CUDA(ctx, call)                                                                    \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return corn_set_err((ctx), e_ == cudaErrorMemoryAllocation ? CORN_E_NOMEM : CORN_E_CUDA, \
                                "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

#define CORN_TRY(expr)                       \
    do {                                     \
        int r_ = (expr);                     \
        if (r_ != CORN_OK) return r_;        \
    } while (0)

// replace the context-owned resident batch (frees the previous one)
void corn_ctx_adopt(corn_ctx *ctx, corn_dbatch *db);

// grow-only scratch; contents are NOT preserved on growth
int corn_dbuf_reserve(corn_ctx *ctx, corn_dbuf *b, size_t bytes);

// kernel launch bookkeeping (gpu_launches in bench.py comes from here)
static inline void corn_count_launch(corn_ctx *ctx, unsigned n = 1)
{
    ctx->timing.launches += n;
    ctx->total_launches += n;
}

#define CORN_LAUNCH_CHECK(ctx)                                                                  \
    do {                                                                                        \
        cudaError_t e_ = cudaGetLastError();                                                    \
        if (e_ != cudaSuccess)                                                                  \
            return corn_set_err((ctx), CORN_E_CUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__, \
                                cudaGetErrorString(e_));                                        \
    } while (0)

// --------------------------------------------------------------------------------------------
// exclusive scans (scan.cu).  n may be 0.  total (device pointer, may be NULL) receives the sum.
// --------------------------------------------------------------------------------------------
int corn_scan_u32(corn_ctx *ctx, const uint32_t *d_in, uint32_t *d_out, size_t n, uint32_t *d_total);
int corn_scan_u32x4(corn_ctx *ctx, const uint4 *d_in, uint4 *d_out, size_t n, uint4 *d_total);

// small synchronous readback through pinned scratch (<= 4 KiB)
int corn_read_small(corn_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);

// telofind.cu: look at the totals / flags of an un-synced telofind_dev(out == NULL); repeats the sparse
// phase synchronously if a speculative buffer was too small.  Stream must be idle (after a sync).
int corn_telofind_resolve(corn_ctx *ctx);
// same, with the first 32 bytes of ctx->misc (totals, counters) already read back by the caller
int corn_telofind_resolve_with(corn_ctx *ctx, const uint32_t tot[8]);

// sdust_wide.cu: the generic sdust instance for 128 < W <= CORN_SDUST_MAX_W (one thread per chunk, serial routines)
#define CORN_SDUST_MAX_W 1024
struct corn_sdust_wide_params {
    const uint8_t  *seq;
    const uint32_t *rec_off, *rec_len, *chunk_base;
    uint32_t n_rec, n_chunks;
    int T, W, C;
    uint32_t cap;
    uint64_t *slots;          // n_chunks * cap interval slots
    uint32_t *cnt, *err;
    uint8_t  *state;          // n_chunks rows of state_stride bytes (window ring, counters, perfect-interval slots)
    size_t    state_stride;
};
size_t corn_sdust_wide_state_stride(int W);
int corn_sdust_wide_scan(corn_ctx *ctx, const corn_sdust_wide_params &P);

// pinned host result blocks handed to the caller (freed by corn_gpu_*_free)
void *corn_host_alloc(size_t bytes);
void  corn_host_free(void *p);

// --------------------------------------------------------------------------------------------
// device helpers
// --------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t corn_lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ uint32_t corn_warp_sum(uint32_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// inclusive warp scan
__device__ __forceinline__ uint32_t corn_warp_iscan(uint32_t v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// first index i in [0,n) with a[i] > x   (upper bound); a ascending
__device__ __forceinline__ uint32_t corn_upper_bound(const uint32_t *__restrict__ a, uint32_t n, uint32_t x)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// first index i in [0,n) with a[i] >= x  (lower bound)
__device__ __forceinline__ uint32_t corn_lower_bound(const uint32_t *__restrict__ a, uint32_t n, uint32_t x)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
#endif
