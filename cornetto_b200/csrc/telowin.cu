// cornetto_b200/csrc/telowin.cu -- windowed telomere density on the GPU.
//
// Replaces the paint loop and process_scaffold() of src/telomere_windows.c:28-43,75-79.
// The reference paints one byte per base and re-reads 1000 bytes for each 200-bp step; here
// the marked bases are counted once per 200-bp bin (a byte per bin, <= 200), a window is the
// sum of five consecutive bins, and the double-precision test
//        (double)car / den >= threshold_adj                       (src/telomere_windows.c:37)
// is evaluated with the same IEEE division on the device.
//
// Two ways to fill the bins:
//   disjoint runs  (runs left on the device by telofind for a border-free motif whose strands
//                   cannot overlap): every run adds its overlap with each bin it touches;
//   general        (runs given by the caller, any order, overlaps allowed -- the text-driven
//                   `cornetto telowin` path): paint a 1-bit-per-base map with atomicOr (union
//                   semantics, exactly the reference's byte map), then popcount per bin.
#include <algorithm>
#include <vector>

#include "corn_internal.cuh"

namespace {

// per-record tables: bin_base[r] = first bin of record r in the bin array (each record owns
// nbins(r) + 4 zero bins so a window can always read five), n_win[r] = number of windows the
// reference evaluates: i = 0,200,... until i + 1000 >= len  (src/telomere_windows.c:31,40-41).
struct WinTables {
    const uint32_t *rec_len;
    uint32_t       *bin_base;   // [n_rec+1]
    uint32_t        n_rec;
};

__host__ __device__ inline uint32_t nwin_of(uint32_t len)
{
    if (len == 0) return 0;                       // den == 0 -> NaN compare, never printed
    if (len <= 1000) return 1;
    return (len - 1000u + 199u) / 200u + 1u;      // last i is the first multiple of 200 with i+1000 >= len
}

__device__ __forceinline__ void bin_add(uint8_t *bins, uint32_t bin, uint32_t amount)
{
    // bins are bytes; total per bin <= 200, so adding into the containing word never carries
    uint32_t *w = (uint32_t *)(bins + (bin & ~3u));
    atomicAdd(w, amount << (8u * (bin & 3u)));
}

// disjoint runs -> bins.  One thread per run.
// (the run count comes from device memory when the producing telofind has not been synced yet)
__global__ void __launch_bounds__(256) k_bins_from_runs(const corn_run_t *__restrict__ runs, uint64_t n_run_host,
                                                        const uint4 *__restrict__ totals, uint32_t run_capacity,
                                                        const uint32_t *__restrict__ bin_base, uint8_t *bins)
{
    uint64_t n_run = n_run_host;
    if (totals) { const uint4 t4 = *totals; n_run = (uint64_t)t4.x + t4.z; if (n_run > run_capacity) return; }
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_run; t += (uint64_t)gridDim.x * blockDim.x) {
        const corn_run_t r = runs[t];
        const uint32_t b0 = bin_base[r.rec];
        uint32_t s = r.start;
        while (s < r.end) {
            const uint32_t bin = s / 200u;
            const uint32_t lim = min(r.end, (bin + 1u) * 200u);
            bin_add(bins, b0 + bin, lim - s);
            s = lim;
        }
    }
}

// general path: paint bits.  bit_base[r] = first bit (multiple of 32) of record r.  One warp per run.
__global__ void __launch_bounds__(256) k_paint_runs(const corn_run_t *__restrict__ runs, uint64_t n_run,
                                                    const uint32_t *__restrict__ rec_len,
                                                    const uint64_t *__restrict__ bit_base, uint32_t n_rec, uint32_t *bitmap)
{
    const uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= n_run) return;
    const corn_run_t r = runs[wid];
    if (r.rec >= n_rec) return;
    const uint32_t len = rec_len[r.rec];
    const uint32_t s = r.start, e = min(r.end, len);     // the reference would write out of bounds past len
    if (s >= e) return;
    uint32_t *bm = bitmap + (bit_base[r.rec] >> 5);
    const uint32_t w0 = s >> 5, w1 = (e - 1) >> 5;
    for (uint32_t w = w0 + lane; w <= w1; w += 32) {
        uint32_t m = 0xffffffffu;
        if (w == w0) m &= 0xffffffffu << (s & 31);
        if (w == w1) m &= 0xffffffffu >> (31 - ((e - 1) & 31));
        atomicOr(bm + w, m);
    }
}

// bitmap -> bins.  One thread per bin (pad bins stay 0).
__global__ void __launch_bounds__(256) k_bins_from_bitmap(const uint32_t *__restrict__ bitmap, const uint64_t *__restrict__ bit_base,
                                                          const uint32_t *__restrict__ bin_base, const uint32_t *__restrict__ rec_len,
                                                          uint32_t n_rec, uint32_t n_bins_total, uint8_t *bins)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_bins_total) return;
    const uint32_t rec = corn_upper_bound(bin_base, n_rec, g) - 1;
    const uint32_t k = g - bin_base[rec];
    const uint32_t len = rec_len[rec];
    const uint32_t s = k * 200u;
    if (s >= len) { bins[g] = 0; return; }
    const uint32_t e = min(len, s + 200u);
    const uint32_t *bm = bitmap + (bit_base[rec] >> 5);
    uint32_t cnt = 0;
    for (uint32_t w = s >> 5; w <= (e - 1) >> 5; ++w) {
        uint32_t m = 0xffffffffu;
        if (w == (s >> 5)) m &= 0xffffffffu << (s & 31);
        if (w == ((e - 1) >> 5)) m &= 0xffffffffu >> (31 - ((e - 1) & 31));
        cnt += __popc(__ldg(bm + w) & m);
    }
    bins[g] = (uint8_t)cnt;
}

// windows.  Bin k of record r is window i = 200k when k < n_win(r).  A thread owns 16 consecutive
// bin indices (one 16-byte load + the 4 bins that follow); with a positive threshold a stretch of
// 20 empty bins cannot hold a passing window (car = 0), which is the overwhelmingly common case.
// Pass 1 records a 16-bit pass mask per thread and counts per block; after the block scan pass 2
// re-derives only the passing windows and writes them in ascending (record, i) order.
#define WIN_PER_THREAD 16u

__device__ __forceinline__ bool window_at(const uint8_t *__restrict__ bins, const uint32_t *__restrict__ bin_base,
                                          const uint32_t *__restrict__ rec_len, uint32_t n_rec, uint32_t g, double thr,
                                          corn_window_t &w)
{
    const uint32_t rec = corn_upper_bound(bin_base, n_rec, g) - 1;
    const uint32_t k = g - bin_base[rec];
    const uint32_t len = rec_len[rec];
    if (k >= nwin_of(len)) return false;
    const uint8_t *b = bins + g;
    const uint32_t car = (uint32_t)b[0] + b[1] + b[2] + b[3] + b[4];
    const uint32_t i = k * 200u;
    const uint32_t den = (i + 1000u < len) ? 1000u : len - i;
    w.rec = rec; w.start = i; w.end = i + den; w.car = car;
    return ((double)car / (double)den) >= thr;           // same IEEE double division as src/telomere_windows.c:37
}

// car_min_full: smallest integer car with (double)car / 1000.0 >= thr, found on the host with the
// same double arithmetic (0 if every count passes, > 1000 if none can): full windows (den = 1000)
// are then decided by integer compares on sliding 5-bin sums.  Threads whose bins come within a
// few bins of a record end (partial windows, pad bins) take the exact per-window path.
__global__ void __launch_bounds__(256) k_windows_mark(const uint8_t *__restrict__ bins, const uint32_t *__restrict__ bin_base,
                                                      const uint32_t *__restrict__ rec_len, uint32_t n_rec, uint32_t n_bins_total,
                                                      double thr, uint32_t car_min_full, uint16_t *__restrict__ mask_out, uint32_t *__restrict__ blk_cnt,
                                                      uint32_t *__restrict__ hot_blocks, uint32_t *__restrict__ n_hot)
{
    __shared__ uint32_t warp_cnt[8];
    __shared__ uint32_t blk_rec[2];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t g0 = t * WIN_PER_THREAD;
    // record of the block's first and last bin: almost always the same one, then no thread searches
    if (threadIdx.x < 2) {
        const uint32_t g = min(n_bins_total - 1, (blockIdx.x * blockDim.x + (threadIdx.x ? blockDim.x - 1 : 0)) * WIN_PER_THREAD + (threadIdx.x ? WIN_PER_THREAD - 1 : 0));
        blk_rec[threadIdx.x] = corn_upper_bound(bin_base, n_rec, g) - 1;
    }
    __syncthreads();
    uint32_t mask = 0;
    if (g0 < n_bins_total) {
        const uint32_t rec = blk_rec[0] == blk_rec[1] ? blk_rec[0] : corn_upper_bound(bin_base, n_rec, g0) - 1;
        const uint32_t k0 = g0 - bin_base[rec];
        const uint32_t len = rec_len[rec];
        // windows k0 .. k0+15 all full and inside this record?
        const bool interior = len > 1000u && (uint64_t)(k0 + WIN_PER_THREAD - 1) * 200u + 1000u < len;
        if (interior) {
            const uint4 v = __ldg((const uint4 *)(bins + g0));
            const uint32_t tail = __ldg((const uint32_t *)(bins + g0 + 16));
            const uint32_t w[5] = { v.x, v.y, v.z, v.w, tail };
            uint32_t b[20];
#pragma unroll
            for (int i = 0; i < 20; ++i) b[i] = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
            uint32_t car = b[0] + b[1] + b[2] + b[3] + b[4];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (car >= car_min_full) mask |= 1u << j;
                if (j < 15) car += b[j + 5] - b[j];
            }
        } else {
            corn_window_t w;
            for (uint32_t j = 0; j < WIN_PER_THREAD && g0 + j < n_bins_total; ++j)
                if (window_at(bins, bin_base, rec_len, n_rec, g0 + j, thr, w)) mask |= 1u << j;
        }
        mask_out[t] = (uint16_t)mask;
    }
    const uint32_t c = corn_warp_sum(__popc(mask));
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int i = 0; i < 8; ++i) s += warp_cnt[i];
        blk_cnt[blockIdx.x] = s;
        if (s) hot_blocks[atomicAdd(n_hot, 1u)] = blockIdx.x;      // order is irrelevant: offsets come from the scan
    }
}

__global__ void __launch_bounds__(256) k_windows_write(const uint8_t *__restrict__ bins, const uint32_t *__restrict__ bin_base,
                                                       const uint32_t *__restrict__ rec_len, uint32_t n_rec, uint32_t n_bins_total,
                                                       double thr, const uint16_t *__restrict__ mask_in, const uint32_t *__restrict__ blk_cnt,
                                                       const uint32_t *__restrict__ blk_off, uint32_t capacity, corn_window_t *out,
                                                       const uint32_t *__restrict__ hot_blocks, const uint32_t *__restrict__ n_hot)
{
    __shared__ uint32_t warp_cnt[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n = *n_hot;
    for (uint32_t h = blockIdx.x; h < n; h += gridDim.x) {         // only the blocks of the mark pass that found something
        const uint32_t blk = hot_blocks[h];
        const uint32_t t = blk * blockDim.x + threadIdx.x;
        const uint32_t g0 = t * WIN_PER_THREAD;
        const uint32_t mask = g0 < n_bins_total ? mask_in[t] : 0u;
        const uint32_t k = __popc(mask);
        const uint32_t incl = corn_warp_iscan(k, lane);
        if (lane == 31) warp_cnt[warp] = incl;
        __syncthreads();
        if (mask) {
            uint32_t off = blk_off[blk] + incl - k;
            for (int i = 0; i < warp; ++i) off += warp_cnt[i];
            uint32_t m = mask;
            while (m) {
                const uint32_t j = __ffs(m) - 1;
                m &= m - 1;
                corn_window_t w;
                window_at(bins, bin_base, rec_len, n_rec, g0 + j, thr, w);
                if (off < capacity) out[off] = w;
                ++off;
            }
        }
        __syncthreads();
    }
}

// Fused path (bins and the hot-bin list come from telofind's run assembly).  car_min_full >= 5 * CORN_HOT_BIN, so a
// full window (den = 1000, i + 1000 < len) that passes holds a listed bin: the thread of a listed bin g looks at the
// five windows that contain it and reports those of which g is the FIRST listed bin.  The last window of every record
// (den = len - i <= 1000: the only one the reference tests with a shorter denominator, src/telomere_windows.c:33-41)
// is decided by one thread per record with the exact double-precision test.  Windows are written in any order; the
// host puts the few thousand of them into (record, start) order.
__global__ void __launch_bounds__(256) k_windows_hot(const uint8_t *__restrict__ bins, const uint32_t *__restrict__ bin_base,
                                                     const uint32_t *__restrict__ rec_len, uint32_t n_rec, double thr, uint32_t car_min_full,
                                                     const uint32_t *__restrict__ hot_list, const uint32_t *__restrict__ hot_count,
                                                     uint32_t capacity, corn_window_t *out, uint32_t *n_out)
{
    const uint32_t n_hot = min(*hot_count, CORN_HOT_CAP);
    const uint32_t n_items = n_hot + n_rec;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_items; t += gridDim.x * blockDim.x) {
        if (t < n_hot) {
            const uint32_t g = hot_list[t];
            const uint32_t rec = corn_upper_bound(bin_base, n_rec, g) - 1;
            const uint32_t b0 = bin_base[rec], k0 = g - b0, nwin = nwin_of(rec_len[rec]);
            if (nwin < 2) continue;                                   // no full window in this record
            const uint32_t k_hi = min(k0, nwin - 2u), k_lo = k0 >= 4u ? k0 - 4u : 0u;
            for (uint32_t k = k_lo; k <= k_hi; ++k) {
                bool first = true;
                for (uint32_t j = k; j < k0; ++j) first &= bins[b0 + j] < CORN_HOT_BIN;
                if (!first) continue;
                const uint8_t *b = bins + b0 + k;
                const uint32_t car = (uint32_t)b[0] + b[1] + b[2] + b[3] + b[4];
                if (car >= car_min_full) {
                    const uint32_t o = atomicAdd(n_out, 1u);
                    if (o < capacity) { corn_window_t w; w.rec = rec; w.start = k * 200u; w.end = k * 200u + 1000u; w.car = car; out[o] = w; }
                }
            }
        } else {
            const uint32_t rec = t - n_hot;
            const uint32_t nwin = nwin_of(rec_len[rec]);
            if (nwin == 0) continue;
            corn_window_t w;
            if (window_at(bins, bin_base, rec_len, n_rec, bin_base[rec] + nwin - 1u, thr, w)) {
                const uint32_t o = atomicAdd(n_out, 1u);
                if (o < capacity) out[o] = w;
            }
        }
    }
}

__global__ void k_bit_bases(const uint32_t *__restrict__ rec_len, uint32_t *__restrict__ words, uint32_t n_rec)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rec) words[r] = (rec_len[r] + 31u) / 32u + 1u;   // 32-bit words per record
}

__global__ void k_words_to_bits(const uint32_t *__restrict__ word_base, uint64_t *__restrict__ bit_base, uint32_t n)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) bit_base[r] = (uint64_t)word_base[r] * 32u;
}

}  // namespace

// (record, start) order = the reference's print order (records in file order, i ascending).  The key of a window is
// its global bin number bin_base[rec] + start / 200 -- unique and ascending in that order -- and a few thousand 32-bit
// keys are put in order by an LSD radix sort (three 11-bit passes, ~10 us; a comparison sort of the structs took ~100).
static void order_windows(corn_window_t *w, uint32_t n, const uint32_t *bin_base)
{
    if (n < 2) return;
    bool sorted = true;
    for (uint32_t i = 1; i < n && sorted; ++i) sorted = w[i - 1].rec < w[i].rec || (w[i - 1].rec == w[i].rec && w[i - 1].start < w[i].start);
    if (sorted) return;
    std::vector<uint32_t> key(n), idx(n), idx2(n);
    for (uint32_t i = 0; i < n; ++i) { key[i] = bin_base[w[i].rec] + w[i].start / 200u; idx[i] = i; }
    uint32_t cnt[2048];
    for (int pass = 0; pass < 3; ++pass) {
        const int sh = 11 * pass;
        memset(cnt, 0, sizeof cnt);
        for (uint32_t i = 0; i < n; ++i) ++cnt[(key[idx[i]] >> sh) & 2047u];
        uint32_t acc = 0;
        for (int b = 0; b < 2048; ++b) { const uint32_t c = cnt[b]; cnt[b] = acc; acc += c; }
        for (uint32_t i = 0; i < n; ++i) idx2[cnt[(key[idx[i]] >> sh) & 2047u]++] = idx[i];
        idx.swap(idx2);
    }
    std::vector<corn_window_t> tmp(w, w + n);
    for (uint32_t i = 0; i < n; ++i) w[i] = tmp[idx[i]];
}

// common tail of both window paths: hand the host array over, fill the timing, and settle an un-synced telofind
static int telowin_finish(corn_ctx *ctx, corn_windows_t *out, corn_window_t *h_win, uint32_t n_win, const uint32_t hv[16], int pending,
                          const corn_timing_t &t_find, const corn_hits_t *hits, const corn_contigs_t *contigs, double thr)
{
    ctx->last_n_win = n_win;
    out->n_win = n_win;
    out->win = h_win; out->_owner = h_win;
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[8], ctx->ev[9]);
    cudaEventElapsedTime(&ctx->timing.post_ms, ctx->ev[9], ctx->ev[10]);
    cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[10], ctx->ev[11]);
    ctx->timing.out_bytes = (uint64_t)n_win * sizeof(corn_window_t);
    if (pending) {
        // the stream is idle now: look at the telofind totals/flags that were left unchecked
        const corn_timing_t t_win = ctx->timing;
        ctx->timing = t_find;
        const uint64_t before = ctx->total_launches;
        int r = corn_telofind_resolve_with(ctx, hv);
        if (r != CORN_OK) { corn_gpu_windows_free(out); return r; }
        if (ctx->total_launches != before) {       // a buffer had been too small and the runs were rebuilt: redo the windows
            corn_gpu_windows_free(out);
            return corn_gpu_telowin(ctx, hits, contigs, thr, out);
        }
        // report the fused step as one: scan = the telofind scan kernel, post = every other kernel
        ctx->timing.post_ms += t_win.post_ms;
        ctx->timing.d2h_ms += t_win.d2h_ms;
        ctx->timing.launches += t_win.launches;
        ctx->timing.out_bytes += t_win.out_bytes;
    }
    return CORN_OK;
}

extern "C" int corn_gpu_telowin(corn_ctx_t *ctx, const corn_hits_t *hits, const corn_contigs_t *contigs,
                                double thr, corn_windows_t *out)
{
    if (!ctx || !out) return CORN_E_ARG;
    CORN_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (hits && ctx->pending) {                    // unrelated call: settle the outstanding telofind first
        CORN_CUDA(ctx, cudaStreamSynchronize(st));
        CORN_TRY(corn_telofind_resolve(ctx));
    }
    out->win = NULL; out->n_win = 0; out->_owner = NULL;

    // ---- record lengths on the device --------------------------------------------------------
    uint32_t n_rec = 0;
    const uint32_t *d_len = NULL, *h_len = NULL;
    uint64_t n_run = 0;
    const corn_run_t *d_runs = NULL;
    int disjoint = 0;
    const int pending = !hits && ctx->pending;     // fused call right after an un-synced telofind_dev(out == NULL)
    const corn_timing_t t_find = ctx->timing;      // its launch count so far (times are filled in by the resolve below)
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[8], st));
    if (!hits) {
        if (!ctx->last_db) return corn_set_err(ctx, CORN_E_STATE, "telowin(hits == NULL) needs a preceding telofind on this context");
        if (contigs && contigs->n != ctx->last_db->n_rec) return corn_set_err(ctx, CORN_E_ARG, "contigs->n != records of the last telofind");
        n_rec = ctx->last_db->n_rec;
        d_len = ctx->last_db->d_rec_len;
        h_len = ctx->last_db->h_rec_len;
        n_run = ctx->last_n_run;
        d_runs = (const corn_run_t *)ctx->runs.p;
        disjoint = ctx->last_runs_disjoint;
        if (pending && !disjoint) {                // the general path sizes its work from the run count: settle first
            CORN_CUDA(ctx, cudaStreamSynchronize(st));
            CORN_TRY(corn_telofind_resolve(ctx));
            return corn_gpu_telowin(ctx, hits, contigs, thr, out);
        }
        memset(&ctx->timing, 0, sizeof ctx->timing);
    } else {
        memset(&ctx->timing, 0, sizeof ctx->timing);
        if (!contigs || (contigs->n && !contigs->length) || (hits->n_run && !hits->run)) return CORN_E_ARG;
        n_rec = contigs->n;
        n_run = hits->n_run;
        const size_t need = sizeof(uint32_t) * ((size_t)n_rec + 1) + sizeof(corn_run_t) * (n_run + 1) + 64;
        CORN_TRY(corn_dbuf_reserve(ctx, &ctx->sd_tab, need));     // borrowed for the uploads
        uint32_t *dl = (uint32_t *)ctx->sd_tab.p;
        corn_run_t *dr = (corn_run_t *)((uint8_t *)ctx->sd_tab.p + ((sizeof(uint32_t) * ((size_t)n_rec + 1) + 15) & ~(size_t)15));
        if (n_rec) CORN_CUDA(ctx, cudaMemcpyAsync(dl, contigs->length, sizeof(uint32_t) * n_rec, cudaMemcpyHostToDevice, st));
        if (n_run) CORN_CUDA(ctx, cudaMemcpyAsync(dr, hits->run, sizeof(corn_run_t) * n_run, cudaMemcpyHostToDevice, st));
        d_len = dl; d_runs = dr;
        h_len = contigs->length;
        disjoint = 0;
    }
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[9], st));
    if (n_rec == 0) { CORN_CUDA(ctx, cudaStreamSynchronize(st)); return CORN_OK; }

    // ---- tables ---------------------------------------------------------------------------------
    // misc: [0,32) telofind totals/counters | [32,48) telowin counters (one readback fetches both)
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->misc, 4096));
    uint32_t *d_tot = (uint32_t *)((uint8_t *)ctx->misc.p + 32);
    CORN_CUDA(ctx, cudaMemsetAsync(d_tot, 0, 16, st));
    const size_t tab_bytes = ((size_t)n_rec + 2) * (2 * sizeof(uint32_t) + sizeof(uint64_t) + sizeof(uint32_t)) + 256;
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->wins, tab_bytes));
    // first bin of every record: host arithmetic on the (host-known) lengths, cached with a resident batch
    uint32_t *bin_base_buf = (uint32_t *)ctx->wins.p;    // [n_rec+1] (text-driven path)
    const uint32_t *bin_base = NULL;
    uint64_t bins64 = 0, words64 = 0;
    for (uint32_t r = 0; r < n_rec; ++r) words64 += (h_len[r] + 31u) / 32u + 1u;
    if (!hits) {
        bin_base = ctx->last_db->d_bin_base;
        bins64 = ctx->last_db->n_bins_total;
    } else {
        uint32_t *hb = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)n_rec + 1));
        if (!hb) return corn_set_err(ctx, CORN_E_NOMEM, "bin table");
        for (uint32_t r = 0; r < n_rec; ++r) { hb[r] = (uint32_t)(bins64 > 0xFFFFFF00ull ? 0xFFFFFF00ull : bins64); bins64 += corn_nbins_of(h_len[r]); }
        hb[n_rec] = (uint32_t)(bins64 > 0xFFFFFF00ull ? 0xFFFFFF00ull : bins64);
        cudaError_t ce = cudaMemcpyAsync(bin_base_buf, hb, sizeof(uint32_t) * ((size_t)n_rec + 1), cudaMemcpyHostToDevice, st);   // pageable: staged before return
        free(hb);
        if (ce != cudaSuccess) return corn_set_err(ctx, CORN_E_CUDA, "bin table upload: %s", cudaGetErrorString(ce));
        bin_base = bin_base_buf;
    }
    if (bins64 > 0xFFFFFF00ull || words64 > 0xFFFFFF00ull) return corn_set_err(ctx, CORN_E_TOOBIG, "too many bins");
    const uint32_t n_bins_total = (uint32_t)bins64;
    const unsigned gr = (n_rec + 255) / 256;
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->bins, (size_t)n_bins_total + 128));
    uint8_t *bins = (uint8_t *)ctx->bins.p;

    // integer form of the test for full windows, derived with the reference's own double expression
    uint32_t car_min_full = 1001;
    for (uint32_t c = 0; c <= 1000; ++c) if ((double)c / (double)1000 >= thr) { car_min_full = c; break; }
    const bool bins_ready = !hits && disjoint && ctx->bins_for_db == ctx->last_db;      // counted by telofind's run assembly
    if (hits) ctx->bins_for_db = NULL;                  // ctx->bins is about to hold the counts of other runs

    if (bins_ready && car_min_full >= 5u * CORN_HOT_BIN && !getenv("CORNETTO_NO_HOT_WINDOWS")) {
        // ---- fused fast path: only the windows around the listed bins and the last window of every record ----
        size_t cap_win = ctx->events.cap / sizeof(corn_window_t);
        if (cap_win < 65536) { CORN_TRY(corn_dbuf_reserve(ctx, &ctx->events, 65536 * sizeof(corn_window_t))); cap_win = ctx->events.cap / sizeof(corn_window_t); }
        corn_window_t *d_out = (corn_window_t *)ctx->events.p;
        const uint32_t *d_hot_count = (const uint32_t *)((uint8_t *)ctx->misc.p + CORN_MISC_HOT_COUNT);
        corn_window_t *h_win = NULL;
        uint32_t hv[16], n_win = 0;
        for (int attempt = 0; attempt < 2; ++attempt) {
            const uint64_t items = (uint64_t)n_rec + 4096u;
            const unsigned g = (unsigned)((items + 255) / 256 < (uint64_t)ctx->sm_count * 8u ? (items + 255) / 256 : (uint64_t)ctx->sm_count * 8u);
            k_windows_hot<<<g, 256, 0, st>>>(bins, bin_base, d_len, n_rec, thr, car_min_full, (const uint32_t *)ctx->hot.p, d_hot_count,
                                             (uint32_t)(cap_win > 0xFFFFFFFFull ? 0xFFFFFFFFull : cap_win), d_out, d_tot + 2);
            corn_count_launch(ctx);
            CORN_LAUNCH_CHECK(ctx);
            CORN_CUDA(ctx, cudaEventRecord(ctx->ev[10], st));
            size_t spec = (size_t)ctx->last_n_win + ctx->last_n_win / 4 + 4096;
            if (spec > cap_win) spec = cap_win;
            h_win = (corn_window_t *)corn_host_alloc(spec * sizeof(corn_window_t));
            if (!h_win) return corn_set_err(ctx, CORN_E_NOMEM, "pinned alloc of %zu windows", spec);
            CORN_CUDA(ctx, cudaMemcpyAsync(h_win, d_out, spec * sizeof(corn_window_t), cudaMemcpyDeviceToHost, st));
            CORN_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned_small, ctx->misc.p, 64, cudaMemcpyDeviceToHost, st));
            CORN_CUDA(ctx, cudaEventRecord(ctx->ev[11], st));
            CORN_CUDA(ctx, cudaStreamSynchronize(st));
            memcpy(hv, ctx->h_pinned_small, 64);
            n_win = hv[8 + 2];
            if (hv[CORN_MISC_HOT_COUNT / 4] > CORN_HOT_CAP) {       // more hot bins than the list holds (a batch made of telomeres): general path
                corn_host_free(h_win);
                ctx->bins_for_db = ctx->last_db;
                h_win = NULL;
                break;
            }
            if (n_win <= spec) break;
            corn_host_free(h_win);                                  // more windows than expected: make room and write them again
            h_win = NULL;
            if (attempt == 1) return corn_set_err(ctx, CORN_E_INTERNAL, "window count changed between two passes");
            CORN_TRY(corn_dbuf_reserve(ctx, &ctx->events, ((size_t)n_win + 1) * sizeof(corn_window_t)));
            cap_win = ctx->events.cap / sizeof(corn_window_t);
            d_out = (corn_window_t *)ctx->events.p;
            ctx->last_n_win = n_win;
            CORN_CUDA(ctx, cudaMemsetAsync(d_tot, 0, 16, st));
        }
        if (h_win) {
            order_windows(h_win, n_win, ctx->last_db->h_bin_base);
            return telowin_finish(ctx, out, h_win, n_win, hv, pending, t_find, hits, contigs, thr);
        }
        CORN_CUDA(ctx, cudaMemsetAsync(d_tot, 0, 16, st));
    }

    if (bins_ready) {
        // counted already by telofind's run assembly
    } else if (disjoint) {
        CORN_CUDA(ctx, cudaMemsetAsync(bins, 0, (size_t)n_bins_total + 64, st));
        if (n_run || pending) {
            const unsigned g = pending ? (unsigned)ctx->sm_count * 8u : (unsigned)((n_run + 255) / 256);
            k_bins_from_runs<<<g, 256, 0, st>>>(d_runs, n_run, pending ? (const uint4 *)ctx->misc.p : (const uint4 *)NULL,
                                                 ctx->pending_run_cap, bin_base, bins);
            corn_count_launch(ctx);
            CORN_LAUNCH_CHECK(ctx);
        }
    } else {
        // 1 bit per base, records word aligned
        uint32_t *words = bin_base_buf + (n_rec + 1);                     // [n_rec] then scanned in place
        uint64_t *bit_base = (uint64_t *)(((uintptr_t)(words + n_rec + 1) + 7) & ~(uintptr_t)7);
        k_bit_bases<<<gr, 256, 0, st>>>(d_len, words, n_rec);
        corn_count_launch(ctx);
        CORN_TRY(corn_scan_u32(ctx, words, words, n_rec, d_tot + 1));
        const uint32_t n_words = (uint32_t)words64;
        k_words_to_bits<<<gr, 256, 0, st>>>(words, bit_base, n_rec);
        corn_count_launch(ctx);
        CORN_TRY(corn_dbuf_reserve(ctx, &ctx->bitmap, ((size_t)n_words + 2) * sizeof(uint32_t)));
        CORN_CUDA(ctx, cudaMemsetAsync(ctx->bitmap.p, 0, ((size_t)n_words + 2) * sizeof(uint32_t), st));
        if (n_run) {
            const uint64_t threads = n_run * 32;
            k_paint_runs<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_runs, n_run, d_len, bit_base, n_rec, (uint32_t *)ctx->bitmap.p);
            corn_count_launch(ctx);
            CORN_LAUNCH_CHECK(ctx);
        }
        CORN_CUDA(ctx, cudaMemsetAsync(bins + n_bins_total, 0, 64, st));
        k_bins_from_bitmap<<<(n_bins_total + 255) / 256, 256, 0, st>>>((const uint32_t *)ctx->bitmap.p, bit_base, bin_base, d_len, n_rec, n_bins_total, bins);
        corn_count_launch(ctx);
        CORN_LAUNCH_CHECK(ctx);
    }

    // ---- windows: mark + count, scan, write (one host sync, at the end) --------------------------
    const uint32_t n_thr = (n_bins_total + WIN_PER_THREAD - 1) / WIN_PER_THREAD;
    const uint32_t n_blk = (n_thr + 255) / 256;
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->tile_tab, ((size_t)n_blk + 1) * 3 * sizeof(uint32_t) + (size_t)n_thr * sizeof(uint16_t) + 256));
    uint32_t *blk_cnt = (uint32_t *)ctx->tile_tab.p, *blk_off = blk_cnt + n_blk + 1, *hot = blk_off + n_blk + 1;
    uint16_t *wmask = (uint16_t *)(hot + n_blk + 1);
    uint32_t *n_hot = d_tot + 3;
    CORN_CUDA(ctx, cudaMemsetAsync(n_hot, 0, sizeof(uint32_t), st));
    // output capacity is speculative (grow-only, remembered across calls); the write kernel never
    // exceeds it and the total tells us afterwards whether a second write pass is needed
    size_t cap_win = ctx->events.cap / sizeof(corn_window_t);
    if (cap_win < 65536) { CORN_TRY(corn_dbuf_reserve(ctx, &ctx->events, 65536 * sizeof(corn_window_t))); cap_win = ctx->events.cap / sizeof(corn_window_t); }
    k_windows_mark<<<n_blk, 256, 0, st>>>(bins, bin_base, d_len, n_rec, n_bins_total, thr, car_min_full, wmask, blk_cnt, hot, n_hot);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    CORN_TRY(corn_scan_u32(ctx, blk_cnt, blk_off, n_blk, d_tot + 2));
    corn_window_t *d_out = (corn_window_t *)ctx->events.p;
    const unsigned g_write = (unsigned)(n_blk < (uint32_t)ctx->sm_count * 4 ? n_blk : (uint32_t)ctx->sm_count * 4);
    k_windows_write<<<g_write, 256, 0, st>>>(bins, bin_base, d_len, n_rec, n_bins_total, thr, wmask, blk_cnt, blk_off, (uint32_t)cap_win, d_out, hot, n_hot);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[10], st));
    // ONE host sync for the whole call: the counters (and, for a fused call, the telofind totals that
    // were left unchecked) come back together with a speculative copy of the first windows
    size_t spec = (size_t)ctx->last_n_win + ctx->last_n_win / 4 + 4096;
    if (spec > cap_win) spec = cap_win;
    corn_window_t *h_win = (corn_window_t *)corn_host_alloc(spec * sizeof(corn_window_t));
    if (!h_win) return corn_set_err(ctx, CORN_E_NOMEM, "pinned alloc of %zu windows", spec);
    CORN_CUDA(ctx, cudaMemcpyAsync(h_win, d_out, spec * sizeof(corn_window_t), cudaMemcpyDeviceToHost, st));
    CORN_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned_small, ctx->misc.p, 64, cudaMemcpyDeviceToHost, st));
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[11], st));
    CORN_CUDA(ctx, cudaStreamSynchronize(st));
    uint32_t hv[16];
    memcpy(hv, ctx->h_pinned_small, 64);
    const uint32_t n_win = hv[8 + 2];
    if (n_win > cap_win) {                       // first call with many windows (e.g. threshold 0): grow and rewrite
        CORN_TRY(corn_dbuf_reserve(ctx, &ctx->events, ((size_t)n_win + 1) * sizeof(corn_window_t)));
        d_out = (corn_window_t *)ctx->events.p;
        k_windows_write<<<g_write, 256, 0, st>>>(bins, bin_base, d_len, n_rec, n_bins_total, thr, wmask, blk_cnt, blk_off, n_win, d_out, hot, n_hot);
        corn_count_launch(ctx);
        CORN_LAUNCH_CHECK(ctx);
    }
    if (n_win > spec) {                          // more windows than the speculative copy covered
        corn_host_free(h_win);
        h_win = (corn_window_t *)corn_host_alloc((size_t)n_win * sizeof(corn_window_t));
        if (!h_win) return corn_set_err(ctx, CORN_E_NOMEM, "pinned alloc of %u windows", n_win);
        CORN_CUDA(ctx, cudaMemcpyAsync(h_win, d_out, (size_t)n_win * sizeof(corn_window_t), cudaMemcpyDeviceToHost, st));
        CORN_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return telowin_finish(ctx, out, h_win, n_win, hv, pending, t_find, hits, contigs, thr);
}

extern "C" void corn_gpu_windows_free(corn_windows_t *w)
{
    if (!w) return;
    corn_host_free(w->_owner);
    w->win = NULL; w->n_win = 0; w->_owner = NULL;
}
