// cornetto_b200/csrc/depthwin.cu -- windowed depth scan of `cornetto noboringbits` / `boringbits` on the GPU.
//
// Replaces get_regs() (src/boringbits_main.c:315-372) and the selection loops of print_fun_bits() (:425-446) and
// print_boring_bits() (:465-485).  The reference re-sums window_size (2500) depth values of two uint16 arrays for
// every window, stepping by window_inc (50): 50 reads per base and array.  Here a block owns DW_TILE consecutive
// windows of one contig: it sums the window_inc-sized bins they touch once (every base is read 1.05 times), builds an
// inclusive prefix over the bins in shared memory, and a window is then the difference of two prefix entries (plus up
// to window_inc - 1 single values when window_size is not a multiple of window_inc).  The integer means (C division),
// the double-precision mapq ratio test (mq_depth / (double)depth < low_mq_cov_thresh, a float promoted to double, :438)
// and the edge / contig-length rules are evaluated per window; selected windows are compacted in (contig, start) order.
// HBM bound: 4 bytes per base (two uint16 arrays), nothing else of size.
#include <vector>

#include "corn_internal.cuh"

namespace {

constexpr int DW_TILE = 512;        // windows per block
constexpr int DW_THREADS = 256;
constexpr int DW_MAX_SPAN = 1536;   // bins a window may span in the tiled kernel (window_size / window_inc rounded up)

struct DepthParams {
    const uint16_t *depth, *mq;
    const uint64_t *ctg_off;        // [n_ctg] first element
    const uint32_t *ctg_len;        // [n_ctg]
    const uint32_t *tile_base;      // [n_ctg + 1] first tile of each contig
    const uint32_t *win_base;       // [n_ctg + 1] first window of each contig
    uint32_t n_ctg;
    int w, inc, lo, hi, edge_len, min_ctg_len, boring;
    float mq_thr;
    uint32_t *flag;                 // [n_win] 1 = selected
    int32_t *out_depth, *out_mq;    // [n_win]
};

__device__ __forceinline__ uint32_t nreg_of(int length, int w, int inc)
{
    int n = (length - w + inc - 1) / inc + 1;          // C division: truncation toward zero, as :331
    return n < 1 ? 1u : (uint32_t)n;
}

// mask of the 16-bit halves of word i (values 2i and 2i + 1 of a lane's eight) whose value index is below cnt
__device__ __forceinline__ uint32_t below_mask(int cnt, int i)
{
    const int r = cnt - 2 * i;
    return r >= 2 ? 0xFFFFFFFFu : (r == 1 ? 0x0000FFFFu : 0u);
}

// is window [st, end) with these means printed?  (:436-441 for noboringbits, :474-481 for boringbits)
__device__ __forceinline__ bool selected(const DepthParams &P, int ctg_len, int st, int end, int depth, int mq)
{
    const bool fun = depth < P.lo || depth > P.hi || ((double)mq / (double)depth) < (double)P.mq_thr;
    if (!P.boring) return ctg_len >= P.min_ctg_len && fun;
    return ctg_len > P.min_ctg_len && st > P.edge_len && end < ctg_len - P.edge_len && !fun;
}

template <int VPL>
__global__ void __launch_bounds__(DW_THREADS, VPL == 8 ? 6 : 4) k_depth_windows(const DepthParams P)
{
    extern __shared__ uint32_t sm[];                    // prefix of depth bins | prefix of mq bins, n_bins + 1 entries each
    __shared__ uint32_t wsum[2][DW_THREADS / 32];
    const uint32_t tile = blockIdx.x;
    const uint32_t c = corn_upper_bound(P.tile_base, P.n_ctg, tile) - 1;
    const int len = (int)P.ctg_len[c];
    const uint32_t n_reg = nreg_of(len, P.w, P.inc);
    const uint32_t j0 = (tile - P.tile_base[c]) * DW_TILE, j1 = min(n_reg, j0 + DW_TILE);
    const int fw = P.w / P.inc;                          // whole bins in a full window
    const int nb_ctg = (len + P.inc - 1) / P.inc;        // bins of the contig (the last one may be short)
    const int b0 = (int)j0;
    int b1 = (int)j1 - 1 + fw + 1;                       // bins [b0, b1) cover every window of the tile (the partial tail is read directly)
    if (b1 > nb_ctg) b1 = nb_ctg;
    const int n_bins = b1 - b0;
    // All sums are kept modulo 2^32: the reference accumulates a window in an int (:353-359), so only the low 32 bits
    // of a sum reach its division, and differences of prefixes modulo 2^32 are those bits.
    uint32_t *Pd = sm, *Pq = sm + (DW_TILE + DW_MAX_SPAN + 2);
    const uint16_t *d = P.depth + P.ctg_off[c], *q = P.mq + P.ctg_off[c];

    // ---- bin sums with coalesced loads.  The tile's bases [lo, hi) are walked by the warps in steps of 32 x VPL values:
    // a lane takes VPL (8 or 16) consecutive values of each array with 16-byte loads (aligned on the ARRAY, whatever the
    // contig's offset; values outside [lo, hi) are masked) and splits them over the at most two bins they fall into
    // (window_inc >= VPL; increments below 8 take one value at a time).  The part that belongs to the lane's second bin
    // moves one lane up -- with window_inc >= VPL that lane starts in exactly that bin -- and a segmented scan over the
    // lanes (bin numbers rise with the lane) leaves each bin's total in the last lane that touches it: one
    // shared-memory add per bin and warp step.  The kernel is bound by instruction issue, not by the loads, which is
    // why the per-step work (bin number, scan, adds) is spread over 16 values where the increment allows.
    // (One thread per bin walking its 50 values gave 8 % of the HBM roofline -- 32 lanes x 2 bytes per request, 100
    // bytes apart; a warp reduction per bin with 64-bit shared atomics 19 %; 8 values per lane 54 %.)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i <= n_bins; i += DW_THREADS) { Pd[i] = 0; Pq[i] = 0; }
    __syncthreads();
    {
        constexpr int NW = VPL / 2, NL = VPL / 8;                         // words / 16-byte loads per lane and array
        const long long lo = (long long)b0 * P.inc, hi = min((long long)len, (long long)b1 * P.inc);
        const unsigned long long g0 = P.ctg_off[c];                       // element index of the contig's first value
        const unsigned long long first = (g0 + (unsigned long long)lo) & ~7ull;   // 8-aligned on the array, may lie before lo
        // 32-bit positions u relative to `first`: the tile's values are [lo_u, hi_u), value u lies in bin (u - lo_u) / inc
        const uint32_t lo_u = (uint32_t)(g0 + (unsigned long long)lo - first), hi_u = lo_u + (uint32_t)(hi - lo);
        const uint16_t *pd = P.depth + first, *pq = P.mq + first;
        const uint32_t uinc = (uint32_t)P.inc, stride = (DW_THREADS / 32) * 32 * VPL;
        const uint32_t inc_magic = 0xFFFFFFFFu / uinc;                    // x / inc = umulhi(x, magic) or one more
        const int max_seg = (P.inc - 1 + VPL - 1) / VPL + 1;              // lanes one bin can spread over
        // the loads of the warp's next step are issued before this step's values are reduced; a 16-byte piece that lies
        // wholly outside [lo_u, hi_u) is not read and stays zero
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        uint4 a[NL], e[NL], a_nx[NL], e_nx[NL];
        uint32_t u0 = (uint32_t)warp * 32 * VPL + (uint32_t)lane * VPL;   // this lane's first value
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            const uint32_t k = u0 + 8 * l;
            a[l] = zero4; e[l] = zero4;
            if (k + 8 > lo_u && k < hi_u) { a[l] = __ldg((const uint4 *)(pd + k)); e[l] = __ldg((const uint4 *)(pq + k)); }
        }
#pragma unroll 2
        for (; u0 - (uint32_t)lane * VPL < hi_u; u0 += stride) {
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                const uint32_t k = u0 + stride + 8 * l;
                a_nx[l] = zero4; e_nx[l] = zero4;
                if (k + 8 > lo_u && k < hi_u) { a_nx[l] = __ldg((const uint4 *)(pd + k)); e_nx[l] = __ldg((const uint4 *)(pq + k)); }
            }
            // the values of each array stay packed two to a word; masks blank what lies outside [lo_u, hi_u)
            uint32_t wa[NW], we[NW];
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                wa[4 * l] = a[l].x; wa[4 * l + 1] = a[l].y; wa[4 * l + 2] = a[l].z; wa[4 * l + 3] = a[l].w;
                we[4 * l] = e[l].x; we[4 * l + 1] = e[l].y; we[4 * l + 2] = e[l].z; we[4 * l + 3] = e[l].w;
            }
            if (u0 + VPL > lo_u && u0 < hi_u && !(u0 >= lo_u && u0 + VPL <= hi_u)) {
                const int t_lo = lo_u > u0 ? (int)(lo_u - u0) : 0, t_hi = hi_u - u0 >= (uint32_t)VPL ? VPL : (int)(hi_u - u0);
#pragma unroll
                for (int i = 0; i < NW; ++i) { const uint32_t m = below_mask(t_hi, i) & ~below_mask(t_lo, i); wa[i] &= m; we[i] &= m; }
            }
            if (VPL > 8 || P.inc >= 8) {
                // bin of the lane's first value and how many of its values stay in it
                const uint32_t uu = u0 < lo_u ? lo_u : u0;
                uint32_t qb = __umulhi(uu - lo_u, inc_magic);
                if (uu - lo_u - qb * uinc >= uinc) ++qb;
                const int ba = (int)qb;
                const uint32_t split = lo_u + ((uint32_t)ba + 1u) * uinc;  // first value of the next bin (> uu)
                const int keep = split - u0 >= (uint32_t)VPL ? VPL : (int)(split - u0);
                uint32_t td = 0, tq = 0;
#pragma unroll
                for (int i = 0; i < NW; ++i) { td = __dp2a_lo(wa[i], 0x0101u, td); tq = __dp2a_lo(we[i], 0x0101u, tq); }   // IDP.2A: both halves of a word times 1
                uint32_t xd = td, xq = tq;
                if (keep < VPL) {
                    xd = 0; xq = 0;
#pragma unroll
                    for (int i = 0; i < NW; ++i) { const uint32_t m = below_mask(keep, i); xd = __dp2a_lo(wa[i] & m, 0x0101u, xd); xq = __dp2a_lo(we[i] & m, 0x0101u, xq); }
                }
                const uint32_t nd = td - xd, nq = tq - xq;                 // what belongs to the next bin
                const uint32_t ud = __shfl_up_sync(0xffffffffu, nd, 1), uq = __shfl_up_sync(0xffffffffu, nq, 1);
                const int prev = __shfl_up_sync(0xffffffffu, ba, 1);
                if (lane) { xd += ud; xq += uq; }
                const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != ba);
                const int dist = lane - (31 - __clz((int)(heads & (0xffffffffu >> (31 - lane)))));   // lanes since the head of my bin
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    if (o < max_seg) {                                     // (same for the whole grid)
                        const uint32_t yd = __shfl_up_sync(0xffffffffu, xd, o), yq = __shfl_up_sync(0xffffffffu, xq, o);
                        if (dist >= o) { xd += yd; xq += yq; }
                    }
                }
                const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
                if (tail && ba < n_bins) { atomicAdd(&Pd[ba + 1], xd); atomicAdd(&Pq[ba + 1], xq); }
                if (lane == 31 && keep < VPL && ba + 1 < n_bins) { atomicAdd(&Pd[ba + 2], nd); atomicAdd(&Pq[ba + 2], nq); }
            } else {
#pragma unroll
                for (int t = 0; t < VPL; ++t) {
                    if (u0 + t >= lo_u && u0 + t < hi_u) {
                        const int bb = (int)((u0 + t - lo_u) / uinc);
                        atomicAdd(&Pd[bb + 1], (wa[t >> 1] >> (16 * (t & 1))) & 0xFFFFu); atomicAdd(&Pq[bb + 1], (we[t >> 1] >> (16 * (t & 1))) & 0xFFFFu);
                    }
                }
            }
#pragma unroll
            for (int l = 0; l < NL; ++l) { a[l] = a_nx[l]; e[l] = e_nx[l]; }
        }
    }
    __syncthreads();
    // inclusive prefix over the bin sums (block scan in rounds of DW_THREADS); Pd[0] = Pq[0] = 0
    uint32_t carry_d = 0, carry_q = 0;
    for (int base = 0; base < n_bins; base += DW_THREADS) {
        const int b = base + (int)threadIdx.x;
        uint32_t xd = b < n_bins ? Pd[b + 1] : 0u, xq = b < n_bins ? Pq[b + 1] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t yd = __shfl_up_sync(0xffffffffu, xd, o), yq = __shfl_up_sync(0xffffffffu, xq, o);
            if (lane >= o) { xd += yd; xq += yq; }
        }
        if (lane == 31) { wsum[0][warp] = xd; wsum[1][warp] = xq; }
        __syncthreads();
        uint32_t od = carry_d, oq = carry_q, td = 0, tq = 0;
        for (int w2 = 0; w2 < DW_THREADS / 32; ++w2) { if (w2 < warp) { od += wsum[0][w2]; oq += wsum[1][w2]; } td += wsum[0][w2]; tq += wsum[1][w2]; }
        if (b < n_bins) { Pd[b + 1] = od + xd; Pq[b + 1] = oq + xq; }
        carry_d += td; carry_q += tq;
        __syncthreads();
    }

    const uint32_t wb = P.win_base[c];
    for (uint32_t j = j0 + threadIdx.x; j < j1; j += DW_THREADS) {
        const int st = (int)j * P.inc;
        int end = st + P.w;
        uint32_t sd, sq;
        const int r = (int)(j - j0);                     // first bin of the window, relative to b0
        if (end >= len) {                                // clipped window: all bins to the end of the contig
            end = len;
            sd = Pd[n_bins] - Pd[r]; sq = Pq[n_bins] - Pq[r];
        } else {
            sd = Pd[r + fw] - Pd[r]; sq = Pq[r + fw] - Pq[r];
            for (int k = st + fw * P.inc; k < end; ++k) { sd += __ldg(d + k); sq += __ldg(q + k); }
        }
        // the reference accumulates in int: same low 32 bits, then C division by the window length (:357-358)
        const int depth = (int)sd / (end - st), mq = (int)sq / (end - st);
        P.out_depth[wb + j] = depth; P.out_mq[wb + j] = mq;
        P.flag[wb + j] = selected(P, len, st, end, depth, mq) ? 1u : 0u;
    }
}

// windows that span more than DW_MAX_SPAN bins (window_size / window_inc in the thousands): one thread per window,
// plain summation.  Nobody runs the tool that way; it only has to be right.
__global__ void __launch_bounds__(256) k_depth_windows_direct(const DepthParams P, uint32_t n_win_total)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_win_total) return;
    const uint32_t c = corn_upper_bound(P.win_base, P.n_ctg, g) - 1;
    const int len = (int)P.ctg_len[c];
    const uint32_t j = g - P.win_base[c];
    const int st = (int)j * P.inc;
    int end = st + P.w;
    if (end > len) end = len;
    const uint16_t *d = P.depth + P.ctg_off[c], *q = P.mq + P.ctg_off[c];
    uint32_t sd = 0, sq = 0;
    for (int k = st; k < end; ++k) { sd += __ldg(d + k); sq += __ldg(q + k); }
    const int depth = (int)sd / (end - st), mq = (int)sq / (end - st);
    P.out_depth[g] = depth; P.out_mq[g] = mq;
    P.flag[g] = selected(P, len, st, end, depth, mq) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_depth_compact(const DepthParams P, const uint32_t *__restrict__ off, uint32_t n_win_total,
                                                       corn_depth_window_t *out)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_win_total || !P.flag[g]) return;
    const uint32_t c = corn_upper_bound(P.win_base, P.n_ctg, g) - 1;
    const int len = (int)P.ctg_len[c];
    const uint32_t j = g - P.win_base[c];
    corn_depth_window_t w;
    w.ctg = c; w.st = j * (uint32_t)P.inc;
    const int end = (int)w.st + P.w;
    w.end = (uint32_t)(end > len ? len : end);
    w.depth = P.out_depth[g]; w.mq_depth = P.out_mq[g];
    out[off[g]] = w;
}

}  // namespace

extern "C" int corn_gpu_depthwin(corn_ctx_t *ctx, const corn_depth_batch_t *b, const corn_depth_params_t *prm, corn_depth_windows_t *out)
{
    if (!ctx || !b || !prm || !out) return CORN_E_ARG;
    out->win = NULL; out->n_win = 0; out->_owner = NULL;
    if (prm->window_size < 1 || prm->window_inc < 1) return corn_set_err(ctx, CORN_E_ARG, "window size and increment must be positive");
    if (b->n_ctg == 0) return CORN_OK;
    if (!b->depth || !b->mq_depth || !b->offset || !b->length) return CORN_E_ARG;
    CORN_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (ctx->pending) { CORN_CUDA(ctx, cudaStreamSynchronize(st)); CORN_TRY(corn_telofind_resolve(ctx)); }
    memset(&ctx->timing, 0, sizeof ctx->timing);
    const uint32_t n = b->n_ctg;
    // per-contig tables (host arithmetic on the lengths)
    std::vector<uint32_t> tile_base(n + 1), win_base(n + 1);
    uint64_t tiles = 0, wins = 0;
    for (uint32_t c = 0; c < n; ++c) {
        if (b->length[c] == 0 || b->length[c] > 0x7FFFFFFFu) return corn_set_err(ctx, CORN_E_ARG, "contig %u: length %u", c, b->length[c]);
        long long nr = ((long long)b->length[c] - prm->window_size + prm->window_inc - 1) / prm->window_inc + 1;
        if (nr < 1) nr = 1;
        tile_base[c] = (uint32_t)tiles; win_base[c] = (uint32_t)wins;
        tiles += (uint64_t)(nr + DW_TILE - 1) / DW_TILE; wins += (uint64_t)nr;
        if (wins > 0xFFFFFF00ull) return corn_set_err(ctx, CORN_E_TOOBIG, "more than 2^32 windows");
    }
    tile_base[n] = (uint32_t)tiles; win_base[n] = (uint32_t)wins;
    const uint32_t n_win = (uint32_t)wins;

    // device buffers: the two depth arrays go to ctx->cand, tables to sd_tab, per-window arrays to sd_slots
    const size_t elem_bytes = (size_t)b->n_total * sizeof(uint16_t);
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->cand, 2 * ((elem_bytes + 255) & ~(size_t)255) + 256));
    uint16_t *d_depth = (uint16_t *)ctx->cand.p, *d_mq = (uint16_t *)((uint8_t *)ctx->cand.p + ((elem_bytes + 255) & ~(size_t)255));
    const size_t tab_bytes = (size_t)n * (sizeof(uint64_t) + sizeof(uint32_t)) + 2 * ((size_t)n + 1) * sizeof(uint32_t) + 64;
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->sd_tab, tab_bytes));
    uint64_t *d_off = (uint64_t *)ctx->sd_tab.p;
    uint32_t *d_len = (uint32_t *)(d_off + n), *d_tile = d_len + n, *d_win = d_tile + n + 1;
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->sd_slots, 4 * ((size_t)n_win + 1) * sizeof(uint32_t) + 64));
    uint32_t *d_flag = (uint32_t *)ctx->sd_slots.p, *d_offw = d_flag + n_win + 1;
    int32_t *d_od = (int32_t *)(d_offw + n_win + 1), *d_oq = d_od + n_win + 1;
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->misc, 4096));
    uint32_t *d_total = (uint32_t *)((uint8_t *)ctx->misc.p + 3584);

    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    CORN_TRY(corn_h2d(ctx, d_depth, b->depth, elem_bytes));
    CORN_TRY(corn_h2d(ctx, d_mq, b->mq_depth, elem_bytes));
    CORN_TRY(corn_h2d(ctx, d_off, b->offset, (size_t)n * sizeof(uint64_t)));
    CORN_TRY(corn_h2d(ctx, d_len, b->length, (size_t)n * sizeof(uint32_t)));
    CORN_TRY(corn_h2d(ctx, d_tile, tile_base.data(), ((size_t)n + 1) * sizeof(uint32_t)));
    CORN_TRY(corn_h2d(ctx, d_win, win_base.data(), ((size_t)n + 1) * sizeof(uint32_t)));
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));

    DepthParams P;
    P.depth = d_depth; P.mq = d_mq; P.ctg_off = d_off; P.ctg_len = d_len; P.tile_base = d_tile; P.win_base = d_win; P.n_ctg = n;
    P.w = prm->window_size; P.inc = prm->window_inc; P.lo = prm->thresh_low_depth; P.hi = prm->thresh_high_depth;
    P.edge_len = prm->edge_len; P.min_ctg_len = prm->min_ctg_len; P.boring = prm->boring; P.mq_thr = prm->low_mq_cov_thresh;
    P.flag = d_flag; P.out_depth = d_od; P.out_mq = d_oq;
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    if ((prm->window_size + prm->window_inc - 1) / prm->window_inc + 1 <= DW_MAX_SPAN) {
        const size_t smem = 2 * (size_t)(DW_TILE + DW_MAX_SPAN + 2) * sizeof(uint32_t);
        if (prm->window_inc >= 16) k_depth_windows<16><<<(unsigned)tiles, DW_THREADS, smem, st>>>(P);
        else k_depth_windows<8><<<(unsigned)tiles, DW_THREADS, smem, st>>>(P);
    } else {
        k_depth_windows_direct<<<(n_win + 255) / 256, 256, 0, st>>>(P, n_win);
    }
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    CORN_TRY(corn_scan_u32(ctx, d_flag, d_offw, n_win, d_total));
    uint32_t total = 0;
    CORN_TRY(corn_read_small(ctx, &total, d_total, sizeof total));
    if (total) {
        CORN_TRY(corn_dbuf_reserve(ctx, &ctx->events, (size_t)total * sizeof(corn_depth_window_t)));
        k_depth_compact<<<(n_win + 255) / 256, 256, 0, st>>>(P, d_offw, n_win, (corn_depth_window_t *)ctx->events.p);
        corn_count_launch(ctx);
        CORN_LAUNCH_CHECK(ctx);
        out->win = (corn_depth_window_t *)corn_host_alloc((size_t)total * sizeof(corn_depth_window_t));
        if (!out->win) return corn_set_err(ctx, CORN_E_NOMEM, "pinned alloc of %u windows", total);
        out->_owner = out->win;
        CORN_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
        CORN_CUDA(ctx, cudaMemcpyAsync(out->win, ctx->events.p, (size_t)total * sizeof(corn_depth_window_t), cudaMemcpyDeviceToHost, st));
    } else CORN_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
    CORN_CUDA(ctx, cudaStreamSynchronize(st));
    out->n_win = total;
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->timing.scan_ms, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&ctx->timing.post_ms, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[4], ctx->ev[5]);
    ctx->timing.out_bytes = (uint64_t)total * sizeof(corn_depth_window_t);
    return CORN_OK;
}

extern "C" void corn_gpu_depth_windows_free(corn_depth_windows_t *w)
{
    if (!w) return;
    corn_host_free(w->_owner);
    w->win = NULL; w->n_win = 0; w->_owner = NULL;
}
