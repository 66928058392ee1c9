// cornetto_b200/csrc/ingest.cu -- FASTA/FASTQ text -> resident record batch, on the device.
//
// Replaces the byte loop of kseq_read() / ks_getuntil2() (src/kseq.h:102-141,184-224) for plain
// text input; ingest_core.cuh states which text is handled here and why the result equals the
// reference reader's.  The raw file bytes cross PCIe once and everything else happens in HBM:
//
//   k_ing_count      newlines per 8 KiB tile                               (reads the text)
//   scan             tile counts -> first line index of every tile
//   k_ing_positions  ordered list nl[] of newline positions                (reads the text again)
//   k_ing_classify   per line: header?  sequence bytes contributed?  regular?
//   scan(s)          cum[k] = sequence bytes before line k;  FASTA: header index of every line
//   k_ing_records    record table: header offset, length, first sequence byte in cum[] units
//   (host)           layout offsets (CORN_ALIGN), corn_dbatch allocation
//   k_ing_copy       output-driven gather: every 16-byte chunk of the batch finds its record
//                    (rec_off) and its source line (cum) by binary search and is assembled from
//                    unaligned source words; padding chunks are written as zeros.
//
// HBM traffic per text byte: 2 reads (count, positions) + 1 read + 1 write (copy) + ~0.3 B of line
// tables for 60-column FASTA.  The whole phase costs a few percent of the PCIe copy that feeds it.
#include "corn_internal.cuh"
#include "ingest_core.cuh"

namespace {

constexpr uint32_t ING_TILE = 8192;          // text bytes per warp in the two line-finding passes

__device__ __forceinline__ uint32_t nl_mask16(uint4 w)
{
    // bit j = byte j of the 16-byte word is '\n'
    const uint32_t K = 0x0A0A0A0Au, S = 0x08040201u, M = 0x01010101u;
    const uint32_t a = ((__vcmpeq4(w.x, K) & S) * M) >> 24, b = ((__vcmpeq4(w.y, K) & S) * M) >> 24;
    const uint32_t c = ((__vcmpeq4(w.z, K) & S) * M) >> 24, d = ((__vcmpeq4(w.w, K) & S) * M) >> 24;
    return a | (b << 4) | (c << 8) | (d << 12);
}

// text is zero-padded to a whole number of tiles: no bounds checks
__global__ void __launch_bounds__(256) k_ing_count(const uint8_t *__restrict__ text, uint32_t n_tiles, uint32_t *__restrict__ tile_cnt)
{
    const uint32_t tile = blockIdx.x * 8u + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const int lane = threadIdx.x & 31;
    const uint4 *p = (const uint4 *)(text + (size_t)tile * ING_TILE) + lane;
    uint32_t cnt = 0;
#pragma unroll 4
    for (int it = 0; it < (int)(ING_TILE / 512); ++it) cnt += __popc(nl_mask16(__ldg(p + it * 32)));
    cnt = corn_warp_sum(cnt);
    if (lane == 0) tile_cnt[tile] = cnt;
}

__global__ void __launch_bounds__(256) k_ing_positions(const uint8_t *__restrict__ text, uint32_t n_tiles,
                                                       const uint32_t *__restrict__ tile_base, uint32_t *__restrict__ nl)
{
    const uint32_t tile = blockIdx.x * 8u + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const int lane = threadIdx.x & 31;
    const uint4 *p = (const uint4 *)(text + (size_t)tile * ING_TILE) + lane;
    uint32_t run = tile_base[tile];
    for (int it = 0; it < (int)(ING_TILE / 512); ++it) {
        uint32_t m = nl_mask16(__ldg(p + it * 32));
        const uint32_t c = __popc(m);
        const uint32_t incl = corn_warp_iscan(c, lane);
        uint32_t dst = run + incl - c;
        const uint32_t pos0 = tile * ING_TILE + (uint32_t)it * 512u + (uint32_t)lane * 16u;
        while (m) {
            nl[dst++] = pos0 + (uint32_t)(__ffs(m) - 1);
            m &= m - 1;
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

struct LineParams {
    const uint8_t *text;
    const uint32_t *nl;        // [n_lines]: end of every line (position of its '\n', or n for a last line without one)
    uint32_t n_lines, n_groups4;
    int mode, final;
    uint32_t *contrib;         // [n_lines + 1], last entry 0
    uint32_t *hdr;             // [n_lines + 1] (FASTA mode only)
    uint32_t *flags;           // [0] irregular
};

__global__ void __launch_bounds__(256) k_ing_classify(const LineParams P)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    if (k < P.n_lines) {
        const uint32_t s = k ? P.nl[k - 1] + 1u : 0u, e = P.nl[k];
        uint32_t s2 = 0, e2 = 0;
        if (P.mode == ING_MODE_FASTQ && (k & 3u) == 3u) { s2 = P.nl[k - 3] + 1u; e2 = P.nl[k - 2]; }
        const ing_line r = ing_classify(P.text, P.mode, P.final, k, s, e, P.n_groups4, s2, e2);
        P.contrib[k] = r.contrib;
        if (P.mode == ING_MODE_FASTA) P.hdr[k] = r.header;
        bad = r.irregular != 0;
    } else if (k == P.n_lines) {
        P.contrib[k] = 0;
        if (P.mode == ING_MODE_FASTA) P.hdr[k] = 0;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(P.flags, 1u);
}

__global__ void __launch_bounds__(256) k_ing_hdr_compact(const uint32_t *__restrict__ hdr, const uint32_t *__restrict__ hidx,
                                                         uint32_t n_lines, uint32_t *__restrict__ hdr_line)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_lines && hdr[k]) hdr_line[hidx[k]] = k;
}

struct RecParams {
    const uint32_t *nl, *cum, *hdr_line;   // hdr_line == NULL: FASTQ, record r starts at line 4r
    uint32_t n_hdr, n_lines;
    uint32_t *hdr_off, *rec_len, *g0;      // [n_hdr]
};

__global__ void __launch_bounds__(256) k_ing_records(const RecParams P)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.n_hdr) return;
    const uint32_t h = P.hdr_line ? P.hdr_line[r] : 4u * r;
    const uint32_t nx = r + 1 < P.n_hdr ? (P.hdr_line ? P.hdr_line[r + 1] : 4u * (r + 1)) : P.n_lines;
    P.hdr_off[r] = h ? P.nl[h - 1] + 1u : 0u;
    P.g0[r] = P.cum[h];
    P.rec_len[r] = P.cum[nx] - P.cum[h];
}

struct CopyParams {
    const uint8_t *text;
    const uint32_t *nl, *cum;      // cum: [n_lines + 1]
    uint32_t n_lines;
    const uint32_t *rec_off;       // [n_rec + 1]
    const uint32_t *rec_len, *g0;  // [n_rec]
    uint32_t n_rec;
    uint32_t total_bytes;
    uint8_t *dst;
    uint32_t *flags;               // [0] |= 1 on a NUL byte inside a record
    const uint32_t *tile_rec, *tile_line;   // [n_tiles + 1]: search results at the start of every 512 B output tile
};

// upper bound in a[lo, hi): first index with a[i] > x
__device__ __forceinline__ uint32_t ub_range(const uint32_t *__restrict__ a, uint32_t lo, uint32_t hi, uint32_t x)
{
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// where a chunk's bytes come from: (record, position in cum[] units).  Padding chunks point at the last
// byte of their record (any valid position keeps the searches monotone).
__device__ __forceinline__ uint32_t chunk_g(const CopyParams &P, uint32_t r, uint32_t o, uint32_t *want)
{
    const uint32_t q = o - __ldg(P.rec_off + r), len = __ldg(P.rec_len + r);
    *want = q < len ? min(16u, len - q) : 0u;
    return __ldg(P.g0 + r) + (*want ? q : (len ? len - 1u : 0u));
}

// One full-table search per 512 B of output (= per warp of the copy kernel): the per-chunk searches of that
// kernel then run over the handful of records / lines between two consecutive index entries (positions, and
// therefore answers, ascend with the chunk).
__global__ void __launch_bounds__(256) k_ing_tile_index(const CopyParams P, uint32_t n_tiles, uint32_t *__restrict__ tile_rec,
                                                        uint32_t *__restrict__ tile_line)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    if (t == n_tiles) { tile_rec[t] = P.n_rec; tile_line[t] = P.n_lines + 1u; return; }
    const uint32_t o = t << 9;
    const uint32_t rr = ub_range(P.rec_off, 0, P.n_rec + 1u, o);
    uint32_t want;
    const uint32_t g = chunk_g(P, rr - 1u, o, &want);
    tile_rec[t] = rr;
    tile_line[t] = ub_range(P.cum, 0, P.n_lines + 1u, g);
}

// 16 text bytes from an arbitrary address (4-byte aligned loads + byte funnel)
__device__ __forceinline__ uint4 load16_unaligned(const uint8_t *__restrict__ p)
{
    const uint32_t sh = (uint32_t)((size_t)p & 3u);
    const uint32_t *q = (const uint32_t *)(p - sh);
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3), w4 = __ldg(q + 4);
    const uint32_t sel = 0x3210u + 0x1111u * sh;
    return make_uint4(__byte_perm(w0, w1, sel), __byte_perm(w1, w2, sel), __byte_perm(w2, w3, sel), __byte_perm(w3, w4, sel));
}

__device__ __forceinline__ uint32_t has_zero_byte(uint32_t v) { return (v - 0x01010101u) & ~v & 0x80808080u; }

__global__ void __launch_bounds__(256) k_ing_copy(const CopyParams P)
{
    const int lane = threadIdx.x & 31;
    const uint32_t n_chunks = P.total_bytes >> 4;
    const uint32_t c_raw = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t warp_first = c_raw - (uint32_t)lane;
    if (warp_first >= n_chunks) return;
    const bool live = c_raw < n_chunks;
    const uint32_t c = live ? c_raw : n_chunks - 1u;       // idle lanes shadow the last chunk (keeps the searches monotone)
    const uint32_t o = c << 4;
    const uint32_t tile = c_raw >> 5;                      // 32 lanes x 16 B = one 512 B index tile per warp
    const uint32_t r = ub_range(P.rec_off, __ldg(P.tile_rec + tile), __ldg(P.tile_rec + tile + 1), o) - 1u;
    uint32_t want;
    uint32_t g = chunk_g(P, r, o, &want);
    uint32_t k = ub_range(P.cum, __ldg(P.tile_line + tile), __ldg(P.tile_line + tile + 1), g) - 1u;
    uint32_t out[4] = { 0u, 0u, 0u, 0u };
    if (want) {
        uint32_t line_end_g = __ldg(P.cum + k + 1);
        uint32_t src = (k ? __ldg(P.nl + k - 1) + 1u : 0u) + (g - __ldg(P.cum + k));
        // One or two source lines (every chunk of ordinary 60/80-column text): branch-free.  The second
        // line's bytes are fetched from (its start - take), so that they already sit at byte positions
        // >= take of the vector; a byte mask merges the two.
        const uint32_t take = min(want, line_end_g - g);
        const uint32_t next_len = take < want ? __ldg(P.cum + k + 2) - line_end_g : 0u;   // (cum has n_lines + 1 entries and g + want <= total)
        if (take == want || next_len >= want - take) {
            const uint32_t src_b = take < want ? __ldg(P.nl + k) + 1u - take : src;
            const uint4 a = load16_unaligned(P.text + src), b = load16_unaligned(P.text + src_b);
            const uint32_t av[4] = { a.x, a.y, a.z, a.w }, bv[4] = { b.x, b.y, b.z, b.w };
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int keep = (int)take - 4 * w;        // bytes of this word that come from the first line
                const uint32_t m = keep >= 4 ? 0xFFFFFFFFu : (keep <= 0 ? 0u : (1u << (8 * keep)) - 1u);
                out[w] = (av[w] & m) | (bv[w] & ~m);
            }
        } else {                                            // three or more lines (lines shorter than 16 bytes, blank lines)
            uint32_t filled = 0;
            for (;;) {
                const uint32_t tk = min(want - filled, line_end_g - g);
                for (uint32_t i = 0; i < tk; ++i) {
                    const uint32_t b = __ldg(P.text + src + i), at = filled + i;
                    out[at >> 2] |= b << (8u * (at & 3u));
                }
                filled += tk; g += tk;
                if (filled >= want) break;
                ++k;                                        // next line (empty ones and headers contribute nothing)
                while (__ldg(P.cum + k + 1) == g) ++k;
                line_end_g = __ldg(P.cum + k + 1);
                src = __ldg(P.nl + k - 1) + 1u;
            }
        }
        // a NUL inside the record is outside the regular subset; behind the record's last byte the chunk is zeroed
        uint32_t z;
        if (want == 16u) z = has_zero_byte(out[0]) | has_zero_byte(out[1]) | has_zero_byte(out[2]) | has_zero_byte(out[3]);
        else {
            z = 0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int keep = (int)want - 4 * w;        // bytes of this word that belong to the record
                const uint32_t m = keep >= 4 ? 0xFFFFFFFFu : (keep <= 0 ? 0u : (1u << (8 * keep)) - 1u);
                z |= has_zero_byte(out[w] | ~m);
                out[w] &= m;
            }
        }
        if (z) atomicOr(P.flags, 1u);
    }
    if (live) *(uint4 *)(P.dst + o) = make_uint4(out[0], out[1], out[2], out[3]);
}

}  // namespace

// defined in context.cu
int corn_dbatch_from_lengths(corn_ctx *ctx, const uint32_t *length, uint32_t n_rec, corn_dbatch **out);

struct ingest_owner {
    uint64_t *hdr_off;
    uint32_t *length;
};

extern "C" void corn_gpu_ingest_free(corn_ingest_t *ing)
{
    if (!ing) return;
    ingest_owner *o = (ingest_owner *)ing->_owner;
    if (o) { free(o->hdr_off); free(o->length); free(o); }
    ing->hdr_off = NULL; ing->length = NULL; ing->_owner = NULL;
}

extern "C" int corn_gpu_host_register(void *p, uint64_t bytes)
{
    if (!p || !bytes) return CORN_E_ARG;
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorMemoryAllocation ? CORN_E_NOMEM : CORN_E_CUDA; }
    return CORN_OK;
}

extern "C" void corn_gpu_host_unregister(void *p)
{
    if (p && cudaHostUnregister(p) != cudaSuccess) cudaGetLastError();
}

static int ingest_run(corn_ctx *ctx, const uint8_t *text, uint64_t n_text, int final, corn_ingest_t *out)
{
    cudaStream_t st = ctx->stream;
    memset(&ctx->timing, 0, sizeof ctx->timing);
    if (n_text == 0) return CORN_OK;
    if (n_text > CORN_MAX_BATCH_BYTES) return corn_set_err(ctx, CORN_E_TOOBIG, "text block of %llu bytes", (unsigned long long)n_text);
    if (ing_precheck(text, n_text, final)) { out->irregular = 1; return CORN_OK; }
    const int mode = text[0] == '>' ? ING_MODE_FASTA : ING_MODE_FASTQ;
    const uint32_t n = (uint32_t)n_text;
    const uint32_t n_tiles = (n + ING_TILE - 1) / ING_TILE;
    const size_t padded = (size_t)n_tiles * ING_TILE + 64;

    // ---- text to the device ----
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->ing_text, padded));
    uint8_t *d_text = (uint8_t *)ctx->ing_text.p;
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    CORN_TRY(corn_h2d(ctx, d_text, text, n));
    CORN_CUDA(ctx, cudaMemsetAsync(d_text + n, 0, padded - n, st));
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));

    // ---- lines ----
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->misc, 4096));
    uint32_t *d_small = (uint32_t *)ctx->misc.p + 768;      // [0] newline total, [1] header total, [2] sequence bytes, [4] flags
    CORN_CUDA(ctx, cudaMemsetAsync(d_small, 0, 32, st));
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->ing_tab, sizeof(uint32_t) * 2 * ((size_t)n_tiles + 1)));
    uint32_t *tile_cnt = (uint32_t *)ctx->ing_tab.p, *tile_base = tile_cnt + n_tiles + 1;
    k_ing_count<<<(n_tiles + 7) / 8, 256, 0, st>>>(d_text, n_tiles, tile_cnt);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    CORN_TRY(corn_scan_u32(ctx, tile_cnt, tile_base, n_tiles, d_small));
    uint32_t hv[8];
    CORN_TRY(corn_read_small(ctx, hv, d_small, 32));
    const uint32_t n_nl = hv[0];
    const bool open_tail = text[n - 1] != '\n';             // bytes after the last newline
    const uint32_t n_lines = n_nl + ((final && open_tail) ? 1u : 0u);
    if (n_lines == 0) { out->consumed = 0; return CORN_OK; }   // not even one complete line yet
    // text that is mostly newlines (lines of < 8 bytes on average) would need line tables several times its own
    // size: not sequence data worth a device pass -- the serial reader takes it
    if ((uint64_t)n_lines * 8u > (uint64_t)n + 4096u) { out->irregular = 1; return CORN_OK; }

    // line tables: nl | contrib | cum | (FASTA) hdr | hidx | hdr_line
    const size_t L1 = (size_t)n_lines + 1;
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->ing_lines, sizeof(uint32_t) * L1 * (mode == ING_MODE_FASTA ? 6 : 3)));
    uint32_t *nl = (uint32_t *)ctx->ing_lines.p, *contrib = nl + L1, *cum = contrib + L1;
    uint32_t *hdr = cum + L1, *hidx = hdr + L1, *hdr_line = hidx + L1;
    k_ing_positions<<<(n_tiles + 7) / 8, 256, 0, st>>>(d_text, n_tiles, tile_base, nl);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    if (n_lines > n_nl) CORN_CUDA(ctx, cudaMemcpyAsync(nl + n_nl, &n, sizeof(uint32_t), cudaMemcpyHostToDevice, st));   // (n is a local: synced below)

    LineParams lp;
    lp.text = d_text; lp.nl = nl; lp.n_lines = n_lines; lp.n_groups4 = n_lines / 4; lp.mode = mode; lp.final = final;
    lp.contrib = contrib; lp.hdr = hdr; lp.flags = d_small + 4;
    k_ing_classify<<<(unsigned)((L1 + 255) / 256), 256, 0, st>>>(lp);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    CORN_TRY(corn_scan_u32(ctx, contrib, cum, L1, d_small + 2));
    if (mode == ING_MODE_FASTA) CORN_TRY(corn_scan_u32(ctx, hdr, hidx, L1, d_small + 1));
    CORN_TRY(corn_read_small(ctx, hv, d_small, 32));
    if (hv[4]) { out->irregular = 1; return CORN_OK; }
    const uint32_t n_hdr = mode == ING_MODE_FASTA ? hv[1] : n_lines / 4;
    // complete records: all of them at the end of the input, else all but the one still open
    // (FASTQ: a group of four complete lines is a complete record)
    const uint32_t n_rec = (mode == ING_MODE_FASTA && !final) ? (n_hdr ? n_hdr - 1 : 0) : n_hdr;
    if (n_hdr == 0) { out->consumed = final ? n : 0; return CORN_OK; }

    // ---- record table ----
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->ing_rec, sizeof(uint32_t) * 3 * (size_t)n_hdr));
    uint32_t *d_hdr_off = (uint32_t *)ctx->ing_rec.p, *d_len = d_hdr_off + n_hdr, *d_g0 = d_len + n_hdr;
    if (mode == ING_MODE_FASTA) {
        k_ing_hdr_compact<<<(unsigned)((n_lines + 255) / 256), 256, 0, st>>>(hdr, hidx, n_lines, hdr_line);
        corn_count_launch(ctx);
        CORN_LAUNCH_CHECK(ctx);
    }
    RecParams rp;
    rp.nl = nl; rp.cum = cum; rp.hdr_line = mode == ING_MODE_FASTA ? hdr_line : NULL;
    rp.n_hdr = n_hdr; rp.n_lines = n_lines; rp.hdr_off = d_hdr_off; rp.rec_len = d_len; rp.g0 = d_g0;
    k_ing_records<<<(n_hdr + 255) / 256, 256, 0, st>>>(rp);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);

    ingest_owner *own = (ingest_owner *)calloc(1, sizeof *own);
    if (!own) return CORN_E_NOMEM;
    out->_owner = own;
    uint32_t *h_off32 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n_hdr);
    own->hdr_off = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n_hdr);
    own->length = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n_hdr);
    if (!h_off32 || !own->hdr_off || !own->length) { free(h_off32); return corn_set_err(ctx, CORN_E_NOMEM, "record table of %u entries", n_hdr); }
    cudaError_t e = cudaMemcpyAsync(h_off32, d_hdr_off, sizeof(uint32_t) * (size_t)n_hdr, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(own->length, d_len, sizeof(uint32_t) * (size_t)n_hdr, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { free(h_off32); return corn_set_err(ctx, CORN_E_CUDA, "ingest: %s", cudaGetErrorString(e)); }
    for (uint32_t r = 0; r < n_hdr; ++r) own->hdr_off[r] = h_off32[r];
    free(h_off32);
    out->hdr_off = own->hdr_off; out->length = own->length;
    out->n_rec = n_rec;
    if (final) out->consumed = n;
    else if (mode == ING_MODE_FASTA) out->consumed = own->hdr_off[n_hdr - 1];
    else {                                                  // start of line 4 * n_rec = one past the newline before it
        uint32_t endpos = 0;
        CORN_TRY(corn_read_small(ctx, &endpos, nl + (4u * n_rec - 1u), sizeof(uint32_t)));
        out->consumed = (uint64_t)endpos + 1u;
    }
    if (n_rec == 0) return CORN_OK;
    for (uint32_t r = 0; r < n_rec; ++r)
        if (own->length[r] > 0x7FFFFFFFu) return corn_set_err(ctx, CORN_E_TOOBIG, "record %u of %u bytes", r, own->length[r]);

    // ---- layout + copy ----
    uint64_t total = 0;
    for (uint32_t r = 0; r < n_rec; ++r) total += ((uint64_t)own->length[r] + 1 + CORN_ALIGN - 1) / CORN_ALIGN * CORN_ALIGN;
    if (total > CORN_MAX_BATCH_BYTES) { out->irregular = 1; out->n_rec = 0; out->consumed = 0; return CORN_OK; }   // (tiny records: padding outgrew the limit)
    corn_dbatch *db = NULL;
    CORN_TRY(corn_dbatch_from_lengths(ctx, own->length, n_rec, &db));
    out->db = db;
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    CopyParams cp;
    cp.text = d_text; cp.nl = nl; cp.cum = cum; cp.n_lines = n_lines;
    cp.rec_off = db->d_rec_off; cp.rec_len = db->d_rec_len; cp.g0 = d_g0; cp.n_rec = n_rec;
    cp.total_bytes = (uint32_t)db->total_bytes; cp.dst = db->d_seq; cp.flags = d_small + 4;
    const uint32_t n_chunks = cp.total_bytes >> 4;
    if (n_chunks) {
        const uint32_t n_otiles = (n_chunks + 31) / 32;
        CORN_TRY(corn_dbuf_reserve(ctx, &ctx->ing_tab, sizeof(uint32_t) * 2 * ((size_t)n_otiles + 1)));   // (the newline tile counts are no longer needed)
        uint32_t *tile_rec = (uint32_t *)ctx->ing_tab.p, *tile_line = tile_rec + n_otiles + 1;
        cp.tile_rec = tile_rec; cp.tile_line = tile_line;
        k_ing_tile_index<<<(n_otiles + 1 + 255) / 256, 256, 0, st>>>(cp, n_otiles, tile_rec, tile_line);
        k_ing_copy<<<(n_chunks + 255) / 256, 256, 0, st>>>(cp);
        corn_count_launch(ctx, 2);
        CORN_LAUNCH_CHECK(ctx);
    }
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    CORN_TRY(corn_read_small(ctx, hv, d_small, 32));
    if (hv[4]) {                                            // a NUL byte inside a record
        corn_gpu_dbatch_free(ctx, db);
        out->db = NULL; out->irregular = 1; out->n_rec = 0; out->consumed = 0;
        return CORN_OK;
    }
    float a = 0, b = 0;
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&a, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&b, ctx->ev[2], ctx->ev[3]);
    ctx->timing.scan_ms = b;                                // the copy kernel
    ctx->timing.post_ms = a;                                // line finding, tables (incl. the host's share between them)
    ctx->timing.out_bytes = db->total_bytes;
    return CORN_OK;
}

extern "C" int corn_gpu_ingest(corn_ctx_t *ctx, const uint8_t *text, uint64_t n_text, int final, corn_ingest_t *out)
{
    if (!ctx || !out || (n_text && !text)) return CORN_E_ARG;
    memset(out, 0, sizeof *out);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return corn_set_err(ctx, CORN_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    if (ctx->pending) { cudaStreamSynchronize(ctx->stream); int r0 = corn_telofind_resolve(ctx); if (r0 != CORN_OK) return r0; }
    int r = ingest_run(ctx, text, n_text, final, out);
    if (r != CORN_OK || out->irregular) {
        if (out->db) { corn_gpu_dbatch_free(ctx, out->db); out->db = NULL; }
        corn_gpu_ingest_free(out);
        const int irr = out->irregular;
        memset(out, 0, sizeof *out);
        out->irregular = irr;
    }
    return r;
}
