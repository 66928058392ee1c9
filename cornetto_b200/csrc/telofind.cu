// cornetto_b200/csrc/telofind.cu -- telomere motif scan on the GPU.
//
// Replaces disambiguate()+find()+rc(), src/find_telomere.c:24-81: for every record, the maximal
// tandem runs of the motif (strand 0) and then of its reverse complement (strand 1).
//
// Pipeline (all on one stream, no host round trip except two 16-byte readbacks for sizing):
//   k_telofind_scan      HOT: one pass over the sequence bytes.  A warp owns a 32 KiB tile; each
//                        lane turns its 32 bytes into two bit planes, a shift-AND chain yields
//                        candidate occurrence masks for both strands, and the (rare) non-zero
//                        masks are appended in position order to the tile's candidate list with
//                        a ballot/popc rank.  At the end of the tile the same warp verifies the
//                        candidates byte-exactly (lane per candidate chunk) and classifies every
//                        occurrence as run start (no occurrence m bytes before) and/or run end
//                        (none m bytes after) by looking at the bytes themselves -- so tiles,
//                        warps and GPUs never exchange state.
//   scan (scan.cu)       exclusive prefix of the per-tile start/end counts.
//   k_telofind_scatter   writes the four ordered position lists (fwd/rev x start/end).
//   k_telofind_assemble  k-th start pairs with k-th end (per strand); runs are written in the
//                        reference's print order: per record all strand-0 runs, then strand-1.
//   Motifs that can overlap themselves (e.g. AAAAAA, TATATA) need the reference's greedy
//   left-to-right semantics (src/find_telomere.c:49-58): the scan then emits every occurrence
//   and k_telofind_greedy_* walks them per (record, strand).
#include "corn_internal.cuh"
#include "telofind_core.cuh"

namespace {

struct ScanParams {
    const uint8_t *seq;        // d_seq (position 0 of the batch)
    uint32_t       n_tiles;
    uint32_t       n_groups, group_len;   // ticket -> tile permutation (see tile_of_ticket)
    uint32_t      *tile_counter;
    // tile candidate lists, SoA, n_tiles * CORN_TILE_CHUNKS entries each
    uint32_t *c_idx;           // global chunk index (byte position / 32)
    uint32_t *c_a, *c_b;       // scan: candidate masks fwd / rev.  after classification: start masks fwd / rev
    uint32_t *c_c, *c_d;       // after classification: end masks fwd / rev
    uint4    *tile_cnt;        // per tile: #start_f, #end_f, #start_r, #end_r
    uint32_t *tile_ncand;
    uint32_t *dense_list;      // tiles whose candidate list is long (telomeres, satellites): classified by
    uint32_t *dense_count;     // k_telofind_classify_dense with the whole GPU instead of by the scanning warp
    const uint8_t *pat;        // device: fwd[256] then rev[256]
    uint64_t fc, rc;           // packed 2-bit codes (plane matcher)
    int      m;
    int      bordered;         // emit every occurrence as a "start", no ends
};

__device__ __forceinline__ void ld256(const uint8_t *p, uint32_t w[8])
{
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
}

// exact occurrence test with early exit (generic matcher: most positions fail on the first byte)
__device__ __forceinline__ bool occ_dev(const uint8_t *p, const uint8_t *__restrict__ pat, int m)
{
    for (int d = 0; d < m; ++d)
        if (corn_fold(__ldg(p + d)) != __ldg(pat + d)) return false;
    return true;
}

// exact occurrence test without early exit: the byte loads are independent, so a candidate costs one
// load latency instead of m chained ones (candidates are true occurrences almost always, and a
// dense tile -- a telomere -- is verified by a single warp)
__device__ __forceinline__ bool occ_all(const uint8_t *p, const uint8_t *__restrict__ pat, int m)
{
    uint32_t diff = 0;
#pragma unroll 8
    for (int d = 0; d < m; ++d) diff |= (uint32_t)(corn_fold(__ldg(p + d)) ^ __ldg(pat + d));
    return diff == 0;
}

// Verify + classify the candidates of one tile.  Called by the warp that scanned the tile,
// one lane per candidate chunk.  Every candidate bit is verified byte-exactly; whether an
// occurrence starts (ends) a run is read from the verified mask itself when the position m bytes
// before (after) lies in the same chunk, and from the bytes otherwise -- so no state is shared
// between lanes, tiles or GPUs.
__device__ __forceinline__ uint32_t verify_mask(const uint8_t *chunk, uint32_t cand, const uint8_t *__restrict__ pat, int m)
{
    uint32_t v = 0;
    while (cand) {
        const int b = __ffs(cand) - 1;
        cand &= cand - 1;
        if (occ_all(chunk + b, pat, m)) v |= 1u << b;
    }
    return v;
}

__device__ __forceinline__ void heads_tails(const uint8_t *chunk, uint32_t v, const uint8_t *__restrict__ pat, int m,
                                            uint32_t &heads, uint32_t &tails)
{
    if (m < 32) {
        const uint32_t lo = (1u << m) - 1u;                 // bits whose predecessor lies in the previous chunk
        const uint32_t hi = ~(0xffffffffu >> m);            // bits whose successor lies in the next chunk
        heads = v & ~(v << m) & ~lo;
        tails = v & ~(v >> m) & ~hi;
        uint32_t edge = v & lo;
        while (edge) {
            const int b = __ffs(edge) - 1;
            edge &= edge - 1;
            if (!occ_all(chunk + b - m, pat, m)) heads |= 1u << b;
        }
        edge = v & hi;
        while (edge) {
            const int b = __ffs(edge) - 1;
            edge &= edge - 1;
            if (!occ_all(chunk + b + m, pat, m)) tails |= 1u << b;
        }
    } else {
        heads = tails = 0;
        uint32_t rest = v;
        while (rest) {
            const int b = __ffs(rest) - 1;
            rest &= rest - 1;
            if (!occ_all(chunk + b - m, pat, m)) heads |= 1u << b;
            if (!occ_all(chunk + b + m, pat, m)) tails |= 1u << b;
        }
    }
}

__device__ __forceinline__ void classify_entry(const ScanParams &P, size_t slot, uint32_t &n_sf, uint32_t &n_ef, uint32_t &n_sr, uint32_t &n_er)
{
    const int m = P.m;
    const uint32_t idx = __ldcg(P.c_idx + slot);
    const uint8_t *chunk = P.seq + (size_t)idx * CORN_CHUNK_BYTES;
    const uint32_t vf = verify_mask(chunk, __ldcg(P.c_a + slot), P.pat, m);
    const uint32_t vr = verify_mask(chunk, __ldcg(P.c_b + slot), P.pat + 256, m);
    uint32_t sf = vf, ef = 0, sr = vr, er = 0;
    if (!P.bordered) {
        heads_tails(chunk, vf, P.pat, m, sf, ef);
        heads_tails(chunk, vr, P.pat + 256, m, sr, er);
    }
    P.c_a[slot] = sf; P.c_b[slot] = sr;
    P.c_c[slot] = ef; P.c_d[slot] = er;
    n_sf += __popc(sf); n_ef += __popc(ef); n_sr += __popc(sr); n_er += __popc(er);
}

// A tile with more candidate chunks than this is not classified by the warp that scanned it: a
// telomere fills all 1024 slots of its tile, and one warp chewing through them while the rest of the
// GPU has run out of tiles was the kernel's tail.  Such tiles are queued for k_telofind_classify_dense.
#define CORN_DENSE_TILE 96u

__device__ void classify_tile(const ScanParams &P, uint32_t tile, uint32_t cnt, int lane)
{
    const size_t base = (size_t)tile * CORN_TILE_CHUNKS;
    if (cnt > CORN_DENSE_TILE) {
        if (lane == 0) {
            P.tile_ncand[tile] = cnt;
            P.tile_cnt[tile] = make_uint4(0, 0, 0, 0);             // accumulated by k_telofind_classify_dense
            P.dense_list[atomicAdd(P.dense_count, 1u)] = tile;
        }
        return;
    }
    uint32_t n_sf = 0, n_ef = 0, n_sr = 0, n_er = 0;
    for (uint32_t e = lane; e < cnt; e += 32) classify_entry(P, base + e, n_sf, n_ef, n_sr, n_er);
    n_sf = corn_warp_sum(n_sf); n_ef = corn_warp_sum(n_ef);
    n_sr = corn_warp_sum(n_sr); n_er = corn_warp_sum(n_er);
    if (lane == 0) {
        P.tile_cnt[tile] = make_uint4(n_sf, n_ef, n_sr, n_er);
        P.tile_ncand[tile] = cnt;
    }
}

// one block per 256 candidate chunks of a queued tile (grid-stride over the queue), one thread per chunk
__global__ void __launch_bounds__(256) k_telofind_classify_dense(const ScanParams P)
{
    __shared__ uint32_t red[4][8];
    const uint32_t n_items = *P.dense_count * (CORN_TILE_CHUNKS / 256u);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
        const uint32_t tile = P.dense_list[it / (CORN_TILE_CHUNKS / 256u)];
        const uint32_t e = (it % (CORN_TILE_CHUNKS / 256u)) * 256u + threadIdx.x;
        const uint32_t cnt = P.tile_ncand[tile];
        uint32_t n_sf = 0, n_ef = 0, n_sr = 0, n_er = 0;
        if (e < cnt) classify_entry(P, (size_t)tile * CORN_TILE_CHUNKS + e, n_sf, n_ef, n_sr, n_er);
        n_sf = corn_warp_sum(n_sf); n_ef = corn_warp_sum(n_ef);
        n_sr = corn_warp_sum(n_sr); n_er = corn_warp_sum(n_er);
        if (lane == 0) { red[0][warp] = n_sf; red[1][warp] = n_ef; red[2][warp] = n_sr; red[3][warp] = n_er; }
        __syncthreads();
        if (threadIdx.x < 4) {
            uint32_t t = 0;
            for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
            if (t) atomicAdd(&((uint32_t *)&P.tile_cnt[tile])[threadIdx.x], t);
        }
        __syncthreads();
    }
}

// Tickets are handed out in increasing order, but consecutive tickets sweep round-robin over
// n_groups equal slices of the buffer, so the LAST tickets fall at slice ends scattered over the
// genome rather than on the final bytes of the batch.  Assemblies end in a telomere: those last
// tiles are the densest ones, and a dense tile claimed last would run alone on an idle GPU.
__device__ __forceinline__ uint32_t tile_of_ticket(const ScanParams &P, uint32_t t)
{
    const uint32_t full = P.n_groups * P.group_len;          // tickets covered by the permutation
    if (t >= full) return t;                                 // remainder (< n_groups tiles) in natural order
    // (+ half a slice: the last tickets land in the middle of the slices, away from both ends of the batch)
    uint32_t j = t / P.n_groups + P.group_len / 2;
    if (j >= P.group_len) j -= P.group_len;
    return (t % P.n_groups) * P.group_len + j;
}

// ordered append of the lanes' non-zero masks to the tile list
__device__ __forceinline__ void push_candidates(const ScanParams &P, size_t base, uint32_t &cnt, uint32_t chunk_idx,
                                                uint32_t mf, uint32_t mr, int lane)
{
    const bool has = (mf | mr) != 0;
    const uint32_t any = __ballot_sync(0xffffffffu, has);
    if (any) {
        if (has) {
            const uint32_t slot = cnt + __popc(any & corn_lanemask_lt());
            P.c_idx[base + slot] = chunk_idx;
            P.c_a[base + slot] = mf;
            P.c_b[base + slot] = mr;
        }
        cnt += __popc(any);
    }
}

// -------------------------------------------------------------------------------------------
// HOT KERNEL: plane matcher.  M_CT > 0 fixes the motif length at compile time (the default
// 6-mer); the packed codes stay run-time values in uniform registers.
// -------------------------------------------------------------------------------------------
constexpr uint64_t pack_codes(const char *s, int i = 0)
{
    return s[i] ? ((uint64_t)(s[i] == 'A' ? 0 : s[i] == 'C' ? 1 : s[i] == 'T' ? 2 : 3) << (2 * i)) | pack_codes(s, i + 1) : 0;
}
constexpr uint64_t FC_TTAGGG = pack_codes("TTAGGG"), RC_TTAGGG = pack_codes("CCCTAA");

// M_CT > 0: motif length AND codes (FC_CT/RC_CT) are compile-time, so the class selects fold
// into the LOP3 truth tables.  M_CT == 0: run-time motif (length P.m, codes P.fc/P.rc).
template <int M_CT, uint64_t FC_CT, uint64_t RC_CT>
__global__ void __launch_bounds__(256, 5) k_telofind_scan(const ScanParams P)
{
    const int lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) P.tile_counter[1] = 0;      // error counter of the run assembly that follows
    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(P.tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= P.n_tiles) break;
        tile = tile_of_ticket(P, tile);

        const uint8_t *lane_ptr = P.seq + (size_t)tile * CORN_TILE_BYTES + (size_t)lane * CORN_CHUNK_BYTES;
        const size_t   base = (size_t)tile * CORN_TILE_CHUNKS;
        const uint32_t chunk0 = tile * CORN_TILE_CHUNKS + lane;
        uint32_t cnt = 0;
        uint32_t pp1 = 0, pp2 = 0;

        // Rows 0..CORN_TILE_ROWS (the extra one is the next tile's first row: only its planes are
        // needed, for lane 31 of the last row).  Three register buffers rotate so that two rows
        // (2 KiB per warp) are always in flight behind the one being processed.
        uint32_t b0[8], b1[8], b2[8];
        ld256(lane_ptr, b0);
        ld256(lane_ptr + CORN_ROW_BYTES, b1);
        ld256(lane_ptr + 2 * CORN_ROW_BYTES, b2);

        // planes of `buf` (row r); candidates of row r-1 from (pp, planes); then refill buf with row r+3
#define CORN_ROW_STEP(buf, r)                                                                                   \
        {                                                                                                       \
            uint32_t c1, c2;                                                                                    \
            corn_planes32(buf, c1, c2);                                                                         \
            /* row CORN_TILE_ROWS belongs to the next tile: only its first chunk is needed (lane 0) */        \
            if ((r) + 3 < CORN_TILE_ROWS || ((r) + 3 == CORN_TILE_ROWS && lane == 0))                           \
                ld256(lane_ptr + (size_t)((r) + 3) * CORN_ROW_BYTES, buf);                                      \
            if ((r) > 0) {                                                                                      \
                /* lane i needs the planes of the 32 bytes after its chunk of row r-1: lane i+1's previous  */ \
                /* planes, or (lane 31) lane 0's current planes                                              */ \
                const uint32_t n1 = __shfl_sync(0xffffffffu, lane == 0 ? c1 : pp1, (lane + 1) & 31);            \
                const uint32_t n2 = __shfl_sync(0xffffffffu, lane == 0 ? c2 : pp2, (lane + 1) & 31);            \
                uint32_t mf, mr;                                                                                \
                corn_match32<M_CT>(pp1, pp2, n1, n2, M_CT ? FC_CT : P.fc, M_CT ? RC_CT : P.rc, P.m, mf, mr);    \
                push_candidates(P, base, cnt, chunk0 + ((r) - 1) * 32u, mf, mr, lane);                          \
            }                                                                                                   \
            pp1 = c1; pp2 = c2;                                                                                 \
        }
        static_assert((CORN_TILE_ROWS + 1) % 3 == 0, "row loop is unrolled by three");
#pragma unroll 1
        for (uint32_t row = 0; row <= CORN_TILE_ROWS; row += 3) {
            CORN_ROW_STEP(b0, row)
            CORN_ROW_STEP(b1, row + 1)
            CORN_ROW_STEP(b2, row + 2)
        }
#undef CORN_ROW_STEP
        __syncwarp();
        classify_tile(P, tile, cnt, lane);
        __syncwarp();
    }
}

// -------------------------------------------------------------------------------------------
// Generic matcher for motifs with characters outside {A,C,G,T} or longer than 32: plain
// byte compare of every position.  Same tile/candidate machinery, exact masks.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_telofind_scan_generic(const ScanParams P)
{
    const int lane = threadIdx.x & 31;
    const int m = P.m;
    const uint8_t f0 = __ldg(P.pat), r0 = __ldg(P.pat + 256);
    if (blockIdx.x == 0 && threadIdx.x == 0) P.tile_counter[1] = 0;      // error counter of the run assembly that follows
    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(P.tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= P.n_tiles) break;
        tile = tile_of_ticket(P, tile);
        const uint8_t *lane_ptr = P.seq + (size_t)tile * CORN_TILE_BYTES + (size_t)lane * CORN_CHUNK_BYTES;
        const size_t   base = (size_t)tile * CORN_TILE_CHUNKS;
        const uint32_t chunk0 = tile * CORN_TILE_CHUNKS + lane;
        uint32_t cnt = 0;
        for (uint32_t row = 0; row < CORN_TILE_ROWS; ++row) {
            const uint8_t *p = lane_ptr + (size_t)row * CORN_ROW_BYTES;
            uint32_t mf = 0, mr = 0;
            for (int b = 0; b < 32; ++b) {
                const uint8_t c = corn_fold(__ldg(p + b));
                if (c == f0 && occ_dev(p + b, P.pat, m)) mf |= 1u << b;
                if (c == r0 && occ_dev(p + b, P.pat + 256, m)) mr |= 1u << b;
            }
            push_candidates(P, base, cnt, chunk0 + row * 32u, mf, mr, lane);
        }
        __syncwarp();
        classify_tile(P, tile, cnt, lane);
        __syncwarp();
    }
}

// -------------------------------------------------------------------------------------------
// ordered position lists.  One warp per tile.
// -------------------------------------------------------------------------------------------
struct ScatterParams {
    const uint32_t *c_idx, *c_a, *c_b, *c_c, *c_d;
    const uint4    *tile_off;
    const uint32_t *tile_ncand;
    uint32_t n_tiles;
    uint32_t *ev;              // event lists, laid out [start_f | end_f | start_r | end_r] from the totals
    const uint4 *totals;       // device: #start_f, #end_f, #start_r, #end_r
    uint32_t capacity;         // entries available in ev
    int m;
    uint32_t *counters;        // [0] tile ticket counter, [2] dense-tile queue length (reset here)
};

__device__ __forceinline__ void emit_bits(uint32_t mask, uint32_t *dst, uint32_t pos0)
{
    while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        *dst++ = pos0 + b;
    }
}

// one batch = G consecutive candidate chunks of a tile, one per lane of a G-lane group (G = 8 for
// ordinary tiles: their lists hold ~16 entries, so four tiles share a warp; G = 32 for dense
// tiles).  `off` = where this batch's events start in the four lists.  Returns the batch's counts.
template <int G>
__device__ __forceinline__ uint4 scatter_batch(const ScatterParams &P, size_t base, uint32_t e0, uint32_t cnt, uint4 off,
                                               uint32_t *start_f, uint32_t *end_f, uint32_t *start_r, uint32_t *end_r,
                                               int gl /* lane within the group */, bool write)
{
    const uint32_t e = e0 + gl;
    uint32_t idx = 0, sf = 0, sr = 0, ef = 0, er = 0;
    if (e < cnt) {
        idx = P.c_idx[base + e];
        sf = P.c_a[base + e]; sr = P.c_b[base + e];
        ef = P.c_c[base + e]; er = P.c_d[base + e];
    }
    // two 16-bit counters per word: per-lane counts <= 32, group totals <= 1024
    const uint32_t k_f = __popc(sf) | (__popc(ef) << 16), k_r = __popc(sr) | (__popc(er) << 16);
    uint32_t i_f = k_f, i_r = k_r;
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, i_f, o, G), b = __shfl_up_sync(0xffffffffu, i_r, o, G);
        if (gl >= o) { i_f += a; i_r += b; }
    }
    if (write) {
        const uint32_t pos0 = idx * CORN_CHUNK_BYTES;
        const uint32_t x_f = i_f - k_f, x_r = i_r - k_r;          // exclusive
        emit_bits(sf, start_f + off.x + (x_f & 0xFFFFu), pos0);
        emit_bits(ef, end_f + off.y + (x_f >> 16), pos0 + P.m);
        emit_bits(sr, start_r + off.z + (x_r & 0xFFFFu), pos0);
        emit_bits(er, end_r + off.w + (x_r >> 16), pos0 + P.m);
    }
    const uint32_t t_f = __shfl_sync(0xffffffffu, i_f, G - 1, G), t_r = __shfl_sync(0xffffffffu, i_r, G - 1, G);
    return make_uint4(t_f & 0xFFFFu, t_f >> 16, t_r & 0xFFFFu, t_r >> 16);
}

// A block owns 32 consecutive tiles.  Ordinary tiles: one 8-lane group each.  Dense tiles (see
// CORN_DENSE_TILE): all eight warps, batch counts -> scan over the batches -> write.
__global__ void __launch_bounds__(256) k_telofind_scatter(const ScatterParams P)
{
    __shared__ uint4 bat[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4 tot = *P.totals;
    // the scan and the dense-tile pass are done with their ticket counter and queue length: leave them zeroed for
    // the next call (one launch and one dependency less than a reset in front of every scan)
    if (blockIdx.x == 0 && threadIdx.x == 0) { P.counters[0] = 0; P.counters[2] = 0; }
    if ((uint64_t)tot.x + tot.y + tot.z + tot.w > P.capacity) return;     // host grows the buffer and relaunches
    uint32_t *start_f = P.ev, *end_f = start_f + tot.x, *start_r = end_f + tot.y, *end_r = start_r + tot.z;
    const uint32_t tile0 = blockIdx.x * 32u;
    {
        const uint32_t tile = tile0 + (threadIdx.x >> 3);
        const int gl = threadIdx.x & 7;
        uint32_t cnt = tile < P.n_tiles ? P.tile_ncand[tile] : 0u;
        if (cnt > CORN_DENSE_TILE) cnt = 0;
        uint4 off = cnt ? P.tile_off[tile] : make_uint4(0, 0, 0, 0);
        const size_t base = (size_t)tile * CORN_TILE_CHUNKS;
        // the four groups of a warp loop together (shuffles are warp-wide): until the longest list is done
        uint32_t longest = cnt;
        longest = max(longest, __shfl_xor_sync(0xffffffffu, longest, 8));
        longest = max(longest, __shfl_xor_sync(0xffffffffu, longest, 16));
        for (uint32_t e0 = 0; e0 < longest; e0 += 8) {
            const uint4 c = scatter_batch<8>(P, base, e0, cnt, off, start_f, end_f, start_r, end_r, gl, true);
            off.x += c.x; off.y += c.y; off.z += c.z; off.w += c.w;
        }
    }
    for (uint32_t t = 0; t < 32; ++t) {
        const uint32_t tile = tile0 + t;
        if (tile >= P.n_tiles) break;
        const uint32_t cnt = P.tile_ncand[tile];
        if (cnt <= CORN_DENSE_TILE) continue;                              // uniform over the block
        const size_t base = (size_t)tile * CORN_TILE_CHUNKS;
        const uint32_t nb = (cnt + 31) / 32;
        const uint4 zero = make_uint4(0, 0, 0, 0);
        for (uint32_t b = warp; b < nb; b += 8) {
            const uint4 c = scatter_batch<32>(P, base, b * 32, cnt, zero, start_f, end_f, start_r, end_r, lane, false);
            if (lane == 0) bat[b] = c;
        }
        __syncthreads();
        if (warp == 0) {                                                   // exclusive scan over <= 32 batches
            const uint4 c = (uint32_t)lane < nb ? bat[lane] : zero;
            const uint4 o = P.tile_off[tile];
            const uint32_t ix = corn_warp_iscan(c.x, lane), iy = corn_warp_iscan(c.y, lane);
            const uint32_t iz = corn_warp_iscan(c.z, lane), iw = corn_warp_iscan(c.w, lane);
            bat[lane] = make_uint4(o.x + ix - c.x, o.y + iy - c.y, o.z + iz - c.z, o.w + iw - c.w);
        }
        __syncthreads();
        for (uint32_t b = warp; b < nb; b += 8)
            scatter_batch<32>(P, base, b * 32, cnt, bat[b], start_f, end_f, start_r, end_r, lane, true);
        __syncthreads();
    }
}

// -------------------------------------------------------------------------------------------
// run assembly, border-free motifs.  Thread k < n_f handles forward run k, thread n_f + k reverse
// run k.  Output slot: the reference prints, per record, all strand-0 runs then all strand-1
// runs (src/find_telomere.c:49-72), so
//   fwd run k of record r  ->  k + #rev runs in records < r
//   rev run k of record r  ->  k + #fwd runs in records <= r
// -------------------------------------------------------------------------------------------
struct AssembleParams {
    const uint32_t *ev;        // [start_f | end_f | start_r | end_r]
    const uint4 *totals;
    uint32_t ev_capacity, run_capacity;
    const uint32_t *rec_off;   // [n_rec+1]
    uint32_t n_rec;
    uint32_t *rank_f, *rank_r; // [n_rec+1]: number of fwd / rev runs that start before record r
    corn_run_t *out;
    uint32_t *err;
    // telowin's 200-bp bin counts, filled here when runs cannot overlap (bins != NULL): replaces the separate pass over
    // the finished run list, src/telomere_windows.c:75-79 (the paint loop)
    const uint32_t *bin_base;  // [n_rec+1] first bin of each record
    uint8_t  *bins;            // zeroed before the launch
    uint32_t *hot_list;        // bins that reached CORN_HOT_BIN (any order), CORN_HOT_CAP entries
    uint32_t *hot_count;
};

// -------------------------------------------------------------------------------------------
// tile prefix + record ranks in ONE launch (replaces a three-launch scan and k_telofind_ranks).
// Exclusive prefix of the per-tile (start_f, end_f, start_r, end_r) counts by a single-pass scan with decoupled
// look-back: a block owns TP_TILES consecutive tiles, publishes its aggregate, then adds up the aggregates (or the
// first inclusive prefix it meets) of the blocks before it.  A 4 GiB batch has 131 072 tiles = 128 blocks, all
// resident at once, so waiting on a predecessor cannot deadlock.  The block then ranks the records that start inside
// its tiles: rank_f[r] / rank_r[r] = forward / reverse run starts before rec_off[r] = the tile's prefix plus the starts
// of the tile's candidate chunks that lie before the record (records are chunk aligned).
// -------------------------------------------------------------------------------------------
constexpr int TP_THREADS = 256, TP_ITEMS = 4, TP_TILES = TP_THREADS * TP_ITEMS;

struct TilePrefixParams {
    const uint4 *tile_cnt;
    uint4 *tile_off;
    uint32_t n_tiles;
    uint4 *totals;
    uint4 *blk_val;            // [n_blocks][2]: aggregate, inclusive prefix
    uint32_t *blk_flag;        // [n_blocks]: 0 nothing, 1 aggregate published, 2 inclusive prefix published (zeroed before the launch)
    // ranks
    const uint32_t *rec_off; uint32_t n_rec;
    uint32_t *rank_f, *rank_r;
    const uint32_t *tile_ncand, *c_idx, *c_sf, *c_sr;
};

__device__ __forceinline__ uint4 add4(uint4 a, uint4 b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__global__ void __launch_bounds__(TP_THREADS) k_telofind_tile_prefix(const TilePrefixParams P)
{
    __shared__ uint4 warp_tot[TP_THREADS / 32];
    __shared__ uint4 blk_excl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t t0 = blockIdx.x * TP_TILES + threadIdx.x * TP_ITEMS;
    uint4 v[TP_ITEMS], sum = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < TP_ITEMS; ++i) {
        v[i] = t0 + i < P.n_tiles ? P.tile_cnt[t0 + i] : make_uint4(0, 0, 0, 0);
        sum = add4(sum, v[i]);
    }
    // block-level exclusive scan of the per-thread sums
    uint4 inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, inc.x, o), b = __shfl_up_sync(0xffffffffu, inc.y, o);
        const uint32_t c = __shfl_up_sync(0xffffffffu, inc.z, o), d = __shfl_up_sync(0xffffffffu, inc.w, o);
        if (lane >= o) { inc.x += a; inc.y += b; inc.z += c; inc.w += d; }
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint4 wbase = make_uint4(0, 0, 0, 0), agg = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int w = 0; w < TP_THREADS / 32; ++w) { if (w < warp) wbase = add4(wbase, warp_tot[w]); agg = add4(agg, warp_tot[w]); }
    // publish the aggregate, then look back with one warp, 32 predecessors at a time (a single thread walking them one
    // by one made the last of ~100 blocks wait for ~100 dependent global reads: 29 us for a 5 us job)
    if (warp == 0) {
        uint4 excl = make_uint4(0, 0, 0, 0);
        if (lane == 0) {
            P.blk_val[2 * blockIdx.x] = agg;
            if (blockIdx.x == 0) P.blk_val[1] = agg;
            __threadfence();
            atomicExch(&P.blk_flag[blockIdx.x], blockIdx.x == 0 ? 2u : 1u);
        }
        for (int base = (int)blockIdx.x - 1; base >= 0; base -= 32) {
            const int b = base - lane;                          // lane 0 looks at the nearest predecessor
            uint32_t f = 2u;                                    // (in front of block 0: an empty prefix)
            if (b >= 0) { while ((f = atomicAdd(&P.blk_flag[b], 0u)) == 0u) { } __threadfence(); }
            const uint32_t pm = __ballot_sync(0xffffffffu, f == 2u);
            const int stop = pm ? __ffs(pm) - 1 : 32;           // nearest predecessor that already knows its inclusive prefix
            uint4 v = make_uint4(0, 0, 0, 0);
            if (b >= 0 && lane <= stop) {
                const volatile uint4 *pv = (const volatile uint4 *)&P.blk_val[2 * b + (lane == stop ? 1 : 0)];
                v = make_uint4(pv->x, pv->y, pv->z, pv->w);
            }
            v.x = corn_warp_sum(v.x); v.y = corn_warp_sum(v.y); v.z = corn_warp_sum(v.z); v.w = corn_warp_sum(v.w);
            excl = add4(excl, v);
            if (pm) break;
        }
        if (lane == 0) {
            if (blockIdx.x != 0) {
                P.blk_val[2 * blockIdx.x + 1] = add4(excl, agg);
                __threadfence();
                atomicExch(&P.blk_flag[blockIdx.x], 2u);
            }
            blk_excl = excl;
            if (blockIdx.x == gridDim.x - 1) *P.totals = add4(excl, agg);
        }
    }
    __syncthreads();
    uint4 run = add4(blk_excl, add4(wbase, make_uint4(inc.x - sum.x, inc.y - sum.y, inc.z - sum.z, inc.w - sum.w)));
#pragma unroll
    for (int i = 0; i < TP_ITEMS; ++i) {
        if (t0 + i < P.n_tiles) P.tile_off[t0 + i] = run;
        run = add4(run, v[i]);
    }
    __syncthreads();                                    // this block's tile_off entries are visible to the whole block below

    // ---- ranks of the records that start inside this block's tiles (the last block also takes the sentinel n_rec) ----
    const uint32_t tile_lo = blockIdx.x * TP_TILES, tile_hi = min(P.n_tiles, tile_lo + TP_TILES);
    const uint32_t r_lo = corn_lower_bound(P.rec_off, P.n_rec, tile_lo * CORN_TILE_BYTES);
    uint32_t r_hi = tile_hi >= P.n_tiles ? P.n_rec + 1 : corn_lower_bound(P.rec_off, P.n_rec, tile_hi * CORN_TILE_BYTES);
    for (uint32_t r = r_lo + warp; r < r_hi; r += TP_THREADS / 32) {     // one warp per record
        const uint32_t pos = P.rec_off[r];                // (rec_off[n_rec] = total bytes)
        const uint32_t tile = pos / CORN_TILE_BYTES;
        uint32_t f = 0, q = 0, f0, q0;
        if (tile < P.n_tiles) {
            const uint4 off = P.tile_off[tile];
            f0 = off.x; q0 = off.z;
            const size_t base = (size_t)tile * CORN_TILE_CHUNKS;
            const uint32_t n = P.tile_ncand[tile], chunk = pos / CORN_CHUNK_BYTES;
            for (uint32_t e = lane; e < n; e += 32)
                if (P.c_idx[base + e] < chunk) { f += __popc(P.c_sf[base + e]); q += __popc(P.c_sr[base + e]); }
        } else {                                          // behind the last tile: everything
            const uint4 tot = add4(blk_excl, agg);
            f0 = tot.x; q0 = tot.z;
        }
        f = corn_warp_sum(f); q = corn_warp_sum(q);
        if (lane == 0) { P.rank_f[r] = f0 + f; P.rank_r[r] = q0 + q; }
    }
}

__global__ void __launch_bounds__(256) k_telofind_assemble(const AssembleParams P)
{
    const uint4 tot = *P.totals;
    const uint32_t n_f = tot.x, n_r = tot.z;
    if ((uint64_t)tot.x + tot.y + tot.z + tot.w > P.ev_capacity || (uint64_t)n_f + n_r > P.run_capacity) return;
    const uint32_t *start_f = P.ev, *end_f = start_f + tot.x, *start_r = end_f + tot.y, *end_r = start_r + tot.z;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_f + n_r; t += gridDim.x * blockDim.x) {
        const bool rev = t >= n_f;
        const uint32_t k = rev ? t - n_f : t;
        const uint32_t s = rev ? start_r[k] : start_f[k];
        const uint32_t e = rev ? end_r[k] : end_f[k];
        const uint32_t rec = corn_upper_bound(P.rec_off, P.n_rec, s) - 1;   // rec_off[rec] <= s
        const uint32_t r0 = P.rec_off[rec];
        const uint32_t slot = rev ? k + P.rank_f[rec + 1] : k + P.rank_r[rec];
        if (e <= s || e > P.rec_off[rec + 1]) atomicAdd(P.err, 1u);        // would mean a start/end mismatch
        corn_run_t r;
        r.rec = rec; r.strand = rev ? 1u : 0u; r.start = s - r0; r.end = e - r0;
        P.out[slot] = r;
        if (P.bins) {
            // marked bases per 200-bp bin: runs are disjoint here, so every run adds its overlap with each bin it
            // touches (bytes inside a word: a bin never exceeds 200, so the add cannot carry into its neighbour)
            const uint32_t b0 = P.bin_base[rec];
            uint32_t a = r.start;
            while (a < r.end) {
                const uint32_t bin = a / 200u, lim = min(r.end, (bin + 1u) * 200u), g = b0 + bin, sh = 8u * (g & 3u);
                const uint32_t before = (atomicAdd((uint32_t *)(P.bins + (g & ~3u)), (lim - a) << sh) >> sh) & 0xFFu;
                if (before < CORN_HOT_BIN && before + (lim - a) >= CORN_HOT_BIN) {       // crossed once per bin: counts only grow
                    const uint32_t h = atomicAdd(P.hot_count, 1u);
                    if (h < CORN_HOT_CAP) P.hot_list[h] = g;
                }
                a = lim;
            }
        }
    }
}

// -------------------------------------------------------------------------------------------
// greedy run assembly for self-overlapping motifs: thread per (record, strand), two passes
// (count, then write at the scanned offset).  occ lists hold EVERY occurrence, ascending.
//   pos = 0; loop: p = first occurrence >= pos; q = p; while occurrence at q: q += m;
//   emit [p,q); pos = q + 1            (src/find_telomere.c:49-58)
// -------------------------------------------------------------------------------------------
struct GreedyParams {
    const uint32_t *occ_f, *occ_r;
    uint32_t n_f, n_r;
    const uint32_t *rec_off;
    uint32_t n_rec;
    int m;
    uint32_t *cnt;          // [2*n_rec]  order: rec0 fwd, rec0 rev, rec1 fwd, ...
    const uint32_t *off;    // exclusive scan of cnt
    corn_run_t *out;        // NULL in the counting pass
};

__global__ void __launch_bounds__(128) k_telofind_greedy(const GreedyParams P)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2u * P.n_rec) return;
    const uint32_t rec = t >> 1, strand = t & 1u;
    const uint32_t *occ = strand ? P.occ_r : P.occ_f;
    const uint32_t n = strand ? P.n_r : P.n_f;
    const uint32_t r0 = P.rec_off[rec], r1 = P.rec_off[rec + 1];
    uint32_t i = corn_lower_bound(occ, n, r0);
    const uint32_t hi = corn_lower_bound(occ, n, r1);
    uint32_t k = 0;
    corn_run_t *out = P.out ? P.out + P.off[t] : NULL;
    while (i < hi) {
        const uint32_t p = occ[i];
        uint32_t q = p + P.m;
        uint32_t j = i + 1;
        // extend while there is an occurrence exactly at q
        for (;;) {
            while (j < hi && occ[j] < q) ++j;
            if (j < hi && occ[j] == q) { q += P.m; ++j; } else break;
        }
        if (out) { corn_run_t r; r.rec = rec; r.strand = strand; r.start = p - r0; r.end = q - r0; out[k] = r; }
        ++k;
        // resume at q + 1 (an occurrence at q is impossible here)
        i = j;
        while (i < hi && occ[i] <= q) ++i;
    }
    if (!P.out) P.cnt[t] = k;
}

__global__ void k_reset_counter(uint32_t *p, int n) { if ((int)threadIdx.x < n) p[threadIdx.x] = 0; }

}  // namespace

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
template <typename K>
static int scan_grid(corn_ctx *ctx, K kernel)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 4; }
    if (const char *e = getenv("CORNETTO_SCAN_CTAS_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= per_sm) per_sm = v; }
    return ctx->sm_count * per_sm;
}

static int telofind_run(corn_ctx *ctx, const corn_dbatch *db, const char *motif, corn_hits_t *out, int allow_async)
{
    if (ctx->pending) {                            // settle an earlier un-synced call before reusing its buffers
        CORN_CUDA(ctx, cudaSetDevice(ctx->device));
        CORN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        CORN_TRY(corn_telofind_resolve(ctx));
    }
    if (!motif || !motif[0]) return corn_set_err(ctx, CORN_E_ARG, "empty motif");
    if (strlen(motif) > 255) return corn_set_err(ctx, CORN_E_ARG, "motif longer than 255");
    CORN_CUDA(ctx, cudaSetDevice(ctx->device));
    corn_motif_info mi;
    corn_analyse_motif(motif, &mi);
    cudaStream_t st = ctx->stream;
    const float keep_h2d = ctx->timing.h2d_ms;
    memset(&ctx->timing, 0, sizeof ctx->timing);
    ctx->timing.h2d_ms = keep_h2d;

    const uint32_t n_tiles = (uint32_t)((db->total_bytes + CORN_TILE_BYTES - 1) / CORN_TILE_BYTES);
    const size_t n_slots = (size_t)n_tiles * CORN_TILE_CHUNKS;
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->cand, 5 * n_slots * sizeof(uint32_t) + 256));
    // tile_tab: counts (uint4) | offsets (uint4) | ncand (u32) ; misc: totals uint4, counter, err, pattern
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->tile_tab, (size_t)n_tiles * (2 * sizeof(uint4) + 2 * sizeof(uint32_t)) + 256));
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->misc, 4096));

    uint32_t *cand = (uint32_t *)ctx->cand.p;
    ScanParams sp;
    sp.seq = db->d_seq;
    sp.n_tiles = n_tiles;
    sp.n_groups = n_tiles >= 4096 ? 61u : 1u;               // odd, unrelated to contig counts
    sp.group_len = n_tiles / sp.n_groups;
    sp.c_idx = cand; sp.c_a = cand + n_slots; sp.c_b = cand + 2 * n_slots; sp.c_c = cand + 3 * n_slots; sp.c_d = cand + 4 * n_slots;
    sp.tile_cnt = (uint4 *)ctx->tile_tab.p;
    uint4 *tile_off = sp.tile_cnt + n_tiles;
    sp.tile_ncand = (uint32_t *)(tile_off + n_tiles);
    sp.dense_list = sp.tile_ncand + n_tiles;
    uint8_t *misc = (uint8_t *)ctx->misc.p;
    uint4 *d_totals = (uint4 *)misc;                    // 16 B
    sp.tile_counter = (uint32_t *)(misc + 16);          // [0] tile counter, [1] error counter, [2] dense-tile queue length
    uint32_t *d_err = sp.tile_counter + 1;
    sp.dense_count = sp.tile_counter + 2;
    uint8_t *d_pat = misc + 64;                         // 512 B
    sp.pat = d_pat;
    sp.fc = mi.fc; sp.rc = mi.rc; sp.m = mi.m; sp.bordered = mi.bordered;

    // pattern upload (512 bytes), only when the motif changed since the last call on this context
    if (strcmp(ctx->cached_motif, motif) != 0 || ctx->cached_motif[0] == 0) {
        uint8_t hpat[512];
        memset(hpat, 0, sizeof hpat);
        memcpy(hpat, mi.fwd, mi.m); memcpy(hpat + 256, mi.rev, mi.m);
        CORN_CUDA(ctx, cudaStreamSynchronize(st));           // h_pinned_small may still be the target of an earlier readback
        memcpy(ctx->h_pinned_small, hpat, 512);
        CORN_CUDA(ctx, cudaMemcpyAsync(d_pat, ctx->h_pinned_small, 512, cudaMemcpyHostToDevice, st));
        CORN_CUDA(ctx, cudaStreamSynchronize(st));
        snprintf(ctx->cached_motif, sizeof ctx->cached_motif, "%s", motif);
    }
    if (!ctx->counters_clean) {                       // first call on this context, or the previous one did not get as far as its scatter pass
        k_reset_counter<<<1, 32, 0, st>>>(sp.tile_counter, 3);
        corn_count_launch(ctx);
    }
    ctx->counters_clean = 0;
    ctx->bins_for_db = NULL;                          // ctx->bins is about to be rewritten (or left stale)

    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    if (n_tiles) {
        // persistent grid: every SM filled to the occupancy limit (a partial last wave would leave
        // whole SMs without a CTA: the block scheduler packs, it does not balance)
        if (mi.acgt && mi.m <= CORN_MAX_FAST_MOTIF) {
            if (mi.m == 6 && mi.fc == FC_TTAGGG && mi.rc == RC_TTAGGG)
                k_telofind_scan<6, FC_TTAGGG, RC_TTAGGG><<<scan_grid(ctx, k_telofind_scan<6, FC_TTAGGG, RC_TTAGGG>), 256, 0, st>>>(sp);
            else
                k_telofind_scan<0, 0, 0><<<scan_grid(ctx, k_telofind_scan<0, 0, 0>), 256, 0, st>>>(sp);
        } else {
            k_telofind_scan_generic<<<scan_grid(ctx, k_telofind_scan_generic), 256, 0, st>>>(sp);
        }
        corn_count_launch(ctx);
        CORN_LAUNCH_CHECK(ctx);
    }
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    if (n_tiles) {
        k_telofind_classify_dense<<<ctx->sm_count * 4, 256, 0, st>>>(sp);
        corn_count_launch(ctx);
        CORN_LAUNCH_CHECK(ctx);
    }

    // tile prefix + record ranks, one launch
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->ranks, 2 * ((size_t)db->n_rec + 2) * sizeof(uint32_t) + 64));
    uint32_t *rank_f = (uint32_t *)ctx->ranks.p, *rank_r = rank_f + db->n_rec + 2;
    if (n_tiles) {
        const uint32_t n_tp = (n_tiles + TP_TILES - 1) / TP_TILES;
        CORN_TRY(corn_dbuf_reserve(ctx, &ctx->scan_tmp, (size_t)n_tp * (2 * sizeof(uint4) + sizeof(uint32_t)) + 64));
        TilePrefixParams tp;
        tp.tile_cnt = sp.tile_cnt; tp.tile_off = tile_off; tp.n_tiles = n_tiles; tp.totals = d_totals;
        tp.blk_val = (uint4 *)ctx->scan_tmp.p; tp.blk_flag = (uint32_t *)(tp.blk_val + 2 * (size_t)n_tp);
        tp.rec_off = db->d_rec_off; tp.n_rec = db->n_rec; tp.rank_f = rank_f; tp.rank_r = rank_r;
        tp.tile_ncand = sp.tile_ncand; tp.c_idx = sp.c_idx; tp.c_sf = sp.c_a; tp.c_sr = sp.c_b;
        CORN_CUDA(ctx, cudaMemsetAsync(tp.blk_flag, 0, (size_t)n_tp * sizeof(uint32_t), st));
        k_telofind_tile_prefix<<<n_tp, TP_THREADS, 0, st>>>(tp);
        corn_count_launch(ctx);
        CORN_LAUNCH_CHECK(ctx);
    } else {
        CORN_CUDA(ctx, cudaMemsetAsync(d_totals, 0, sizeof(uint4), st));
        CORN_CUDA(ctx, cudaMemsetAsync(rank_f, 0, 2 * ((size_t)db->n_rec + 2) * sizeof(uint32_t), st));
    }

    // Everything below is launched against SPECULATIVE buffer capacities (grow-only, remembered across
    // calls) with the exact totals read by the kernels from device memory; the host looks at the
    // totals once, at the end, and only repeats the sparse phase if a buffer was too small.
    uint64_t n_run = 0;
    uint32_t tot[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };     // totals (4), tile counter, error counter
    for (int attempt = 0; attempt < 2; ++attempt) {
        size_t ev_cap = ctx->events.cap / sizeof(uint32_t), run_cap = ctx->runs.cap / sizeof(corn_run_t);
        const size_t guess = (size_t)(db->total_bytes / 256) + 4096;          // ~8x the random-DNA hit rate
        if (attempt == 0 && ev_cap < 4 * guess) { CORN_TRY(corn_dbuf_reserve(ctx, &ctx->events, 4 * guess * sizeof(uint32_t))); ev_cap = ctx->events.cap / sizeof(uint32_t); }
        if (attempt == 0 && run_cap < guess) { CORN_TRY(corn_dbuf_reserve(ctx, &ctx->runs, guess * sizeof(corn_run_t))); run_cap = ctx->runs.cap / sizeof(corn_run_t); }
        if (ev_cap > 0xFFFFFFFFull) ev_cap = 0xFFFFFFFFull;
        if (run_cap > 0xFFFFFFFFull) run_cap = 0xFFFFFFFFull;
        uint32_t *ev = (uint32_t *)ctx->events.p;
        ScatterParams sc;
        sc.c_idx = sp.c_idx; sc.c_a = sp.c_a; sc.c_b = sp.c_b; sc.c_c = sp.c_c; sc.c_d = sp.c_d;
        sc.tile_off = tile_off; sc.tile_ncand = sp.tile_ncand; sc.n_tiles = n_tiles;
        sc.ev = ev; sc.totals = d_totals; sc.capacity = (uint32_t)ev_cap; sc.m = mi.m; sc.counters = sp.tile_counter;
        if (n_tiles) {
            k_telofind_scatter<<<(n_tiles + 31) / 32, 256, 0, st>>>(sc);
            corn_count_launch(ctx);
            CORN_LAUNCH_CHECK(ctx);
            ctx->counters_clean = 1;
        }
        if (!mi.bordered) {
            AssembleParams ap;
            ap.ev = ev; ap.totals = d_totals; ap.ev_capacity = (uint32_t)ev_cap; ap.run_capacity = (uint32_t)run_cap;
            ap.rec_off = db->d_rec_off; ap.n_rec = db->n_rec;
            ap.rank_f = rank_f; ap.rank_r = rank_r;
            ap.out = (corn_run_t *)ctx->runs.p; ap.err = d_err;
            ap.bin_base = NULL; ap.bins = NULL; ap.hot_list = NULL; ap.hot_count = NULL;
            if (!mi.strands_overlap && db->n_rec && db->d_bin_base && db->n_bins_total <= 0xFFFFFF00ull) {
                // runs cannot overlap: the run assembly also counts telowin's bins (and lists the hot ones)
                CORN_TRY(corn_dbuf_reserve(ctx, &ctx->bins, (size_t)db->n_bins_total + 128));
                CORN_TRY(corn_dbuf_reserve(ctx, &ctx->hot, (size_t)CORN_HOT_CAP * sizeof(uint32_t)));
                ap.bin_base = db->d_bin_base; ap.bins = (uint8_t *)ctx->bins.p;
                ap.hot_list = (uint32_t *)ctx->hot.p; ap.hot_count = (uint32_t *)(misc + CORN_MISC_HOT_COUNT);
                CORN_CUDA(ctx, cudaMemsetAsync(ap.bins, 0, (size_t)db->n_bins_total + 64, st));
                CORN_CUDA(ctx, cudaMemsetAsync(ap.hot_count, 0, sizeof(uint32_t), st));
            }
            if (db->n_rec) {
                k_telofind_assemble<<<ctx->sm_count * 8, 256, 0, st>>>(ap);
                corn_count_launch(ctx);
                CORN_LAUNCH_CHECK(ctx);
                ctx->bins_for_db = ap.bins ? db : NULL;       // (a capacity overflow leaves them empty; the repeat below refills them)
            }
            CORN_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
            if (!out && attempt == 0 && allow_async) {
                // resident batch, results stay on the device: return without a host sync.  The totals and
                // the consistency flags are looked at by the next call that synchronises
                // (corn_gpu_telowin(hits == NULL), corn_gpu_last_timing, ...): corn_telofind_resolve().
                ctx->pending = 1;
                snprintf(ctx->pending_motif, sizeof ctx->pending_motif, "%s", motif);
                ctx->pending_ev_cap = (uint32_t)ev_cap; ctx->pending_run_cap = (uint32_t)run_cap;
                ctx->last_db = db;
                ctx->last_n_run = 0;
                ctx->last_runs_disjoint = !mi.strands_overlap;
                ctx->last_motif_len = mi.m;
                return CORN_OK;
            }
            CORN_TRY(corn_read_small(ctx, tot, d_totals, 32));
            if (tot[0] != tot[1] || tot[2] != tot[3])
                return corn_set_err(ctx, CORN_E_INTERNAL, "start/end counts differ: %u/%u %u/%u", tot[0], tot[1], tot[2], tot[3]);
            n_run = (uint64_t)tot[0] + tot[2];
            const uint64_t n_ev = 2 * n_run;
            if (n_ev <= ev_cap && n_run <= run_cap) break;
            CORN_TRY(corn_dbuf_reserve(ctx, &ctx->events, (n_ev + 4) * sizeof(uint32_t)));
            CORN_TRY(corn_dbuf_reserve(ctx, &ctx->runs, (n_run + 1) * sizeof(corn_run_t)));
        } else {
            // self-overlapping motif: the lists hold every occurrence; greedy pass per (record, strand)
            CORN_TRY(corn_read_small(ctx, tot, d_totals, 16));
            if ((uint64_t)tot[0] + tot[2] > ev_cap) {
                CORN_TRY(corn_dbuf_reserve(ctx, &ctx->events, ((uint64_t)tot[0] + tot[2] + 4) * sizeof(uint32_t)));
                continue;
            }
            if (db->n_rec) {
                const size_t n2 = 2 * (size_t)db->n_rec;
                CORN_TRY(corn_dbuf_reserve(ctx, &ctx->bins, (2 * n2 + 8) * sizeof(uint32_t)));   // (bins is free here: bordered motifs never fuse them)
                GreedyParams gp;
                gp.occ_f = ev; gp.occ_r = ev + tot[0]; gp.n_f = tot[0]; gp.n_r = tot[2];
                gp.rec_off = db->d_rec_off; gp.n_rec = db->n_rec; gp.m = mi.m;
                gp.cnt = (uint32_t *)ctx->bins.p; uint32_t *goff = gp.cnt + n2; gp.off = goff; gp.out = NULL;
                k_telofind_greedy<<<(unsigned)((n2 + 127) / 128), 128, 0, st>>>(gp);
                corn_count_launch(ctx);
                CORN_LAUNCH_CHECK(ctx);
                uint32_t *d_tot = (uint32_t *)d_totals;
                CORN_TRY(corn_scan_u32(ctx, gp.cnt, goff, n2, d_tot));
                uint32_t t32 = 0;
                CORN_TRY(corn_read_small(ctx, &t32, d_tot, 4));
                n_run = t32;
                CORN_TRY(corn_dbuf_reserve(ctx, &ctx->runs, (n_run + 1) * sizeof(corn_run_t)));
                gp.out = (corn_run_t *)ctx->runs.p;
                k_telofind_greedy<<<(unsigned)((n2 + 127) / 128), 128, 0, st>>>(gp);
                corn_count_launch(ctx);
                CORN_LAUNCH_CHECK(ctx);
            }
            CORN_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
            break;
        }
    }

    ctx->last_db = db;
    ctx->last_n_run = n_run;
    ctx->last_runs_disjoint = !mi.bordered && !mi.strands_overlap;
    ctx->last_motif_len = mi.m;
    ctx->timing.out_bytes = n_run * sizeof(corn_run_t);

    if (out) {
        out->run = NULL; out->n_run = n_run; out->_owner = NULL;
        if (n_run) {
            out->run = (corn_run_t *)corn_host_alloc(n_run * sizeof(corn_run_t));
            if (!out->run) return corn_set_err(ctx, CORN_E_NOMEM, "pinned alloc of %llu runs", (unsigned long long)n_run);
            out->_owner = out->run;
            CORN_CUDA(ctx, cudaMemcpyAsync(out->run, ctx->runs.p, n_run * sizeof(corn_run_t), cudaMemcpyDeviceToHost, st));
        }
    }
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
    CORN_CUDA(ctx, cudaEventSynchronize(ctx->ev[5]));
    if (tot[5]) return corn_set_err(ctx, CORN_E_INTERNAL, "%u runs failed the start/end consistency check", tot[5]);
    float all_ms = 0;
    cudaEventElapsedTime(&ctx->timing.scan_ms, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&all_ms, ctx->ev[3], ctx->ev[4]);
    ctx->timing.post_ms = all_ms;
    cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[4], ctx->ev[5]);
    return CORN_OK;
}

int corn_telofind_resolve(corn_ctx *ctx)
{
    if (!ctx->pending) return CORN_OK;
    uint32_t tot[8];
    CORN_TRY(corn_read_small(ctx, tot, ctx->misc.p, 32));
    return corn_telofind_resolve_with(ctx, tot);
}

int corn_telofind_resolve_with(corn_ctx *ctx, const uint32_t tot[8])
{
    if (!ctx->pending) return CORN_OK;
    ctx->pending = 0;
    const uint64_t n_run = (uint64_t)tot[0] + tot[2];
    if (tot[0] != tot[1] || tot[2] != tot[3])
        return corn_set_err(ctx, CORN_E_INTERNAL, "start/end counts differ: %u/%u %u/%u", tot[0], tot[1], tot[2], tot[3]);
    if (2 * n_run > ctx->pending_ev_cap || n_run > ctx->pending_run_cap) {
        // a speculative buffer was too small: nothing was written; run again, synchronously
        char motif[256];
        snprintf(motif, sizeof motif, "%s", ctx->pending_motif);
        return telofind_run(ctx, ctx->last_db, motif, NULL, 0);
    }
    if (tot[5]) return corn_set_err(ctx, CORN_E_INTERNAL, "%u runs failed the start/end consistency check", tot[5]);
    ctx->last_n_run = n_run;
    ctx->timing.out_bytes = n_run * sizeof(corn_run_t);
    float all_ms = 0;
    cudaEventElapsedTime(&ctx->timing.scan_ms, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&all_ms, ctx->ev[3], ctx->ev[4]);
    ctx->timing.post_ms = all_ms;
    return CORN_OK;
}

extern "C" int corn_gpu_telofind_dev(corn_ctx_t *ctx, const corn_dbatch_t *db, const char *motif, corn_hits_t *out)
{
    if (!ctx || !db) return CORN_E_ARG;
    return telofind_run(ctx, db, motif, out, 1);
}

extern "C" int corn_gpu_telofind(corn_ctx_t *ctx, const corn_batch_t *batch, const char *motif, corn_hits_t *out)
{
    if (!ctx || !batch || !out) return CORN_E_ARG;
    corn_dbatch_t *db = NULL;
    corn_ctx_adopt(ctx, NULL);   // retire the previous resident batch first: its buffer is reused by the upload
    CORN_TRY(corn_gpu_upload(ctx, batch, &db));
    const float h2d = ctx->timing.h2d_ms;
    corn_ctx_adopt(ctx, db);   // stays resident for a fused corn_gpu_telowin(hits == NULL)
    ctx->timing.h2d_ms = h2d;
    return telofind_run(ctx, db, motif, out, 1);
}

extern "C" void corn_gpu_hits_free(corn_hits_t *hits)
{
    if (!hits) return;
    corn_host_free(hits->_owner);
    hits->run = NULL; hits->n_run = 0; hits->_owner = NULL;
}
