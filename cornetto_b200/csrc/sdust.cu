// cornetto_b200/csrc/sdust.cu -- symmetric DUST low-complexity masking on the GPU.
//
// Replaces sdust_core(), src/sdust/sdust.c:130-160.  The state machine is inherently serial
// along a sequence, so the parallelism is across chunks: one THREAD per chunk of a record, with
// the exact mid-record start and the seam merge described in sdust_core.cuh.  Per-thread state
// (window ring, 2x64 counters, W slots) lives in shared memory, one 4-byte column per thread so
// that data-dependent indices never cause bank conflicts.
//
// This kernel is instruction-issue / shared-memory bound, not HBM bound (about 60 instructions
// and 10 shared-memory accesses per base against 1 byte of HBM traffic); DESIGN.md says so.
#include "corn_internal.cuh"
#include "sdust_core.cuh"

using namespace sd_narrow;

namespace {

constexpr int SD_BLOCK = 128;

__global__ void k_sdust_nchunks(const uint32_t *__restrict__ rec_len, uint32_t *__restrict__ nch, uint32_t n_rec, uint32_t C)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rec) nch[r] = (rec_len[r] + C - 1) / C;
}

// seq_nt4_table (sdust_core.cuh: sd_nt4) for four bytes at once: 2-bit codes in bits 0..7, validity in
// bits 0..3 of *valid4.  Letters: bits 2..1 of the byte give A 0, C 1, T 2, G 3; the byte must then equal
// that letter once its case bit is cleared; Gray-decoding the index gives A 0, C 1, G 2, T 3.  Bytes
// 0..3 are their own code.
__device__ __forceinline__ uint32_t nt4x4(uint32_t w, uint32_t *valid4)
{
    const uint32_t g = (w >> 1) & 0x03030303u;
    uint32_t sel = (g | (g >> 4)) & 0x00FF00FFu;
    sel = (sel | (sel >> 8)) & 0xFFFFu;                         // one selector nibble per byte
    const uint32_t expect = __byte_perm(0x47544341u, 0u, sel);   // 'A','C','T','G' by index
    const uint32_t is_letter = __vcmpeq4(w & 0xDFDFDFDFu, expect);
    const uint32_t is_low = __vcmpltu4(w, 0x04040404u);
    const uint32_t code = (is_letter & (g ^ ((g >> 1) & 0x01010101u))) | (is_low & w & 0x03030303u);
    uint32_t p = (code | (code >> 6)) & 0x000F000Fu;
    p = (p | (p >> 12)) & 0xFFu;                                // c0 | c1 << 2 | c2 << 4 | c3 << 6
    *valid4 = (((is_letter | is_low) & 0x01010101u) * 0x01020408u) >> 24;
    return p;
}

// byte stream over one record, 16 bytes per global load
struct DevFetch {
    const uint8_t *seq;
    uint32_t codes, valid;          // the 16 bytes of block cblk, decoded (nt4)
    int cblk;
    // sd_nt4(byte i), sixteen positions decoded per load
    __device__ __forceinline__ int nt4(int i)
    {
        const int b = i >> 4;
        if (b != cblk) {
            const uint4 v = __ldg((const uint4 *)(seq + ((size_t)b << 4)));
            uint32_t v0, v1, v2, v3;
            codes = nt4x4(v.x, &v0) | (nt4x4(v.y, &v1) << 8) | (nt4x4(v.z, &v2) << 16) | (nt4x4(v.w, &v3) << 24);
            valid = v0 | (v1 << 4) | (v2 << 8) | (v3 << 12);
            cblk = b;
        }
        const uint32_t k = (uint32_t)i & 15u;
        return ((valid >> k) & 1u) ? (int)((codes >> (2u * k)) & 3u) : 4;
    }
};

// sd_warm_start / sd_warm_quiet (sdust_core.cuh) on the device: the same walk to the left until W triplet positions have
// been passed, but a 16-byte block without a single A/C/G/T is stepped over at once (a chunk or item that starts behind a
// 50 kb gap walks across all of it, and the other 31 lanes wait for that one)
__device__ __forceinline__ int dev_warm_walk(DevFetch &fetch, int p, int W)
{
    if (p <= 0) return 0;
    int need = W, run = 0;
    while (p > 0 && need > 0) {
        if ((p & 15) == 0 && p >= 16) {
            (void)fetch.nt4(p - 1);                       // decodes block [p - 16, p)
            if (fetch.valid == 0u) { p -= 16; run = 0; continue; }
        }
        --p;
        if (fetch.nt4(p) < 4) { if (++run >= 3) --need; }
        else run = 0;
    }
    return p;
}

#define SD_QUICK_POPS 4

struct SdParams {
    const uint8_t  *seq;
    const uint32_t *rec_off, *rec_len;
    const uint32_t *chunk_base;     // [n_rec+1]
    uint32_t n_rec, n_chunks;
    int T, W, C;
    uint32_t cap;
    uint64_t *slots;                // n_chunks * cap
    uint32_t *gslots;               // n_chunks * (W | 1) words, zeroed: the perfect-interval rings
    uint32_t *cnt;                  // [n_chunks]
    uint32_t *err;
    uint32_t *task_counter;
    const uint32_t *task_list;      // ticket -> warp-task, expensive tasks first
    // two-phase execution: the work units are ITEMS (runs of active 64-base blocks) instead of the chunks of a fixed grid;
    // n_chunks then counts the entries of it_list, and slots / gslots / cnt are indexed by item number
    const uint32_t *it_list;        // NULL: chunk mode.  lane's entry -> item number
    const uint32_t *it_rec, *it_c0, *it_c1, *it_flags;
    const uint32_t *list_lo, *list_hi;   // device: this launch serves the entries [*list_lo, *list_hi) of it_list
};

// ---- shared-memory layout of a block -------------------------------------------------------------
//   [ cw | cv columns : 32 rows x SD_BLOCK x 4 B ][ ring arrays : SD_BLOCK x ring_words x 4 B ]
//   [ 64 counter bytes per warp ]
// ring_words is odd, so that the same index in consecutive threads falls into consecutive banks.
// The W slots of the perfect-interval ring live in GLOBAL memory (one 4*slot_words-byte row per
// chunk, L2 resident): they are only touched inside low-complexity sequence, and keeping them out
// of shared memory doubles the number of resident warps (196 B instead of 456 B per thread).
struct SdLayout {
    int ring_words, slot_words;
    __host__ __device__ SdLayout(int W) : ring_words(((W + 3) >> 2) | 1), slot_words(W | 1) {}
    __host__ __device__ size_t bytes() const { return (size_t)SD_BLOCK * 4 * (32 + ring_words) + 64 * (SD_BLOCK / 32); }
};

__device__ __forceinline__ sd_mem sd_mem_of(uint32_t *smem, const SdLayout &lay, int tid, uint32_t *slots)
{
    sd_mem m;
    m.pitch = SD_BLOCK * 4;
    m.cw = (uint8_t *)(smem + tid);
    m.cv = m.cw + 16 * (size_t)m.pitch;
    uint32_t *rings = smem + 32 * SD_BLOCK;
    m.ring = (uint8_t *)(rings + (size_t)tid * lay.ring_words);
    m.slot = slots;
    return m;
}

// a lane's slot row, broadcast from the leader of a cooperative call
__device__ __forceinline__ uint32_t *bcast_ptr(uint32_t *p, int leader)
{
    const unsigned long long v = (unsigned long long)p;
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, leader), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), leader);
    return (uint32_t *)(((unsigned long long)hi << 32) | lo);
}

// ---- warp-cooperative routines ---------------------------------------------------------------------
// find_perfect is needed on ~2 % of the positions of ordinary sequence but costs ~50 dependent
// iterations, and the suffix-shrink loop of shift_window runs ~12 iterations on ~1 % of them; run
// per lane they leave the other 31 lanes idle for thousands of issue slots.  Instead the lanes
// that need them are served one after the other by the WHOLE warp: window index i = 32*b + lane,
// equal-triplet ranks by __match_any_sync, reverse warp scans for the suffix sums and for the
// running maximum of score ratios (the data-parallel form derived in sdust_core.cuh).

// shrink v until the first occurrence of t has been dropped (sd_shift_window_pop)
template <int NB>
__device__ __forceinline__ void pop_coop(int leader, int lane, sd_state &s, int my_t, uint32_t *smem, const SdLayout &lay, int W)   // (does not touch the slots)
{
    const uint32_t FULL = 0xffffffffu;
    __syncwarp();                                     // the leader's own writes to its ring and counts (its last steps) are
                                                      // visible to the lanes that now read them (shuffles order nothing)
    const int wn = __shfl_sync(FULL, s.wn, leader), whead = __shfl_sync(FULL, s.whead, leader);
    const int L = __shfl_sync(FULL, s.L, leader), t = __shfl_sync(FULL, my_t, leader);
    const sd_mem m = sd_mem_of(smem, lay, (threadIdx.x & ~31) + leader, NULL);
    const int v0 = wn - L;                            // first index of v
    const uint32_t le = corn_lanemask_lt() | (1u << lane);
    int x[NB];
    bool inv[NB];
    int p = 1 << 30;                                  // index of the first occurrence of t in v
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int i = 32 * b + lane;
        inv[b] = i >= v0 && i < wn;
        int ri = whead + i; if (ri >= W) ri -= W;
        x[b] = inv[b] ? (int)SD_RING(ri) : -1;
        const uint32_t hit = __ballot_sync(FULL, x[b] == t);
        if (hit && p == (1 << 30)) p = 32 * b + __ffs(hit) - 1;
    }
    int sub = 0;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int i = 32 * b + lane;
        const bool popped = inv[b] && i <= p;
        const int key = popped ? x[b] : 64 + lane;
        const uint32_t mm = __match_any_sync(FULL, key);
        const int cvx = popped ? (int)SD_U8(m.cv, x[b]) : 0;          // count before this block's pops
        __syncwarp();                                 // every lane of a group has read cv[x] before its leader updates it
        if (popped) {
            sub += cvx - __popc(mm & le);             // "rv -= --cv[x]" for the rank-th equal element
            if (lane == __ffs(mm) - 1) SD_U8(m.cv, x[b]) = (uint8_t)(cvx - __popc(mm));
        }
        __syncwarp();
    }
    sub = (int)corn_warp_sum((uint32_t)sub);
    if (lane == leader) { s.rv -= sub; s.L -= p - v0 + 1; }
}

template <int NB>   // 32-position blocks covering the window: 2 for W <= 66, 4 for W <= 128
__device__ __forceinline__ void fp_coop(int leader, int lane, sd_state &s, int my_start, uint32_t *smem, const SdLayout &lay,
                                        uint32_t *my_slots, uint8_t *cnt /* 64 bytes per warp */, int T, int W)
{
    const uint32_t FULL = 0xffffffffu;
    const int wn = __shfl_sync(FULL, s.wn, leader), whead = __shfl_sync(FULL, s.whead, leader);
    const int L = __shfl_sync(FULL, s.L, leader), rv = __shfl_sync(FULL, s.rv, leader);
    const int start = __shfl_sync(FULL, my_start, leader);
    int base = __shfl_sync(FULL, s.pslot, leader) + (start - __shfl_sync(FULL, s.pstart, leader));
    if (base >= W || base < 0) base = (int)((uint32_t)start % (uint32_t)W);
    const int i0 = wn - L - 1;
    const sd_mem m = sd_mem_of(smem, lay, (threadIdx.x & ~31) + leader, bcast_ptr(my_slots, leader));

    if (lane < 16) ((uint32_t *)cnt)[lane] = 0;
    __syncwarp();
    const uint32_t le = corn_lanemask_lt() | (1u << lane);

    int c[NB];
    bool valid[NB];
    const int nb_c = (i0 >> 5) + 1;                   // blocks that hold an index <= i0: the others contribute no c_i
                                                      // (after a long shrink on ordinary sequence i0 < 32: half the work)
#pragma unroll
    for (int b = 0; b < NB; ++b) {                    // ranks, ascending blocks
        const int i = 32 * b + lane;
        valid[b] = i < wn;
        c[b] = 0;
        if (b >= nb_c) continue;                      // (warp-uniform)
        int ri = whead + i; if (ri >= W) ri -= W;
        const int t = valid[b] ? (int)SD_RING(ri) : 64 + lane;
        const int before = valid[b] ? (int)cnt[t] : 0;
        const uint32_t mm = __match_any_sync(FULL, t);
        const int rank = before + __popc(mm & le);
        __syncwarp();                                 // every lane of a group has read cnt[t] before its leader updates it
        if (valid[b] && lane == __ffs(mm) - 1) cnt[t] = (uint8_t)(before + __popc(mm));
        __syncwarp();
        c[b] = (valid[b] && i <= i0) ? (int)SD_U8(m.cw, t) - rank : 0;
    }
    // suffix sums (descending index): new_r(i) = rv + sum_{k >= i} c_k
    int nr[NB];
    bool cand[NB];
    int carry = 0, fmin = 0x7fffffff;
    bool any_cand = false;
#pragma unroll
    for (int b = NB - 1; b >= 0; --b) {
        nr[b] = 0; cand[b] = false;
        if (b >= nb_c) continue;                      // no candidate up there, nothing to add to the carry
        int x = c[b];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_down_sync(FULL, x, o); if (lane + o < 32) x += y; }
        nr[b] = rv + x + carry;
        carry += __shfl_sync(FULL, x, 0);
        const int i = 32 * b + lane, new_l = wn - i - 1;
        const bool eligible = valid[b] && i <= i0;
        cand[b] = eligible && nr[b] * 10 > T * new_l;
        any_cand |= cand[b];
        if (eligible) fmin = min(fmin, T * new_l - 10 * nr[b]);
    }
    if (!__any_sync(FULL, any_cand)) {                // nothing can be inserted: P is left untouched
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) fmin = min(fmin, __shfl_xor_sync(FULL, fmin, o));
        if (lane == leader) s.slack = min(fmin, 1 << 20);     // covers the following steps (sd_slack_push)
        return;
    }
    if (lane == leader) sd_slack_unknown(s);

    // elements (existing slot, candidate) and their exclusive running maximum (descending index)
    int er[NB], el[NB], pr[NB], pl[NB], si[NB];
    bool sv[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int i = 32 * b + lane, new_l = wn - i - 1;
        int k = base + i; if (k >= W) k -= W;
        si[b] = k;
        const uint32_t v = valid[b] ? SD_SLOT(k) : 0u;
        sv[b] = (v & SD_SLOT_VALID) != 0;
        pr[b] = sv[b] ? sd_slot_r(v) : 0;
        pl[b] = sv[b] ? sd_slot_l(v) : 1;
        er[b] = pr[b]; el[b] = pl[b];
        if (cand[b]) sd_fracmax(er[b], el[b], nr[b], new_l);
    }
    int cr = 0, cl = 1;                               // maximum over the blocks above the current one
    uint32_t fresh = 0;
#pragma unroll
    for (int b = NB - 1; b >= 0; --b) {
        int xr = er[b], xl = el[b];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int yr = __shfl_down_sync(FULL, xr, o), yl = __shfl_down_sync(FULL, xl, o);
            if (lane + o < 32) sd_fracmax(xr, xl, yr, yl);
        }
        int mr = __shfl_down_sync(FULL, xr, 1), ml = __shfl_down_sync(FULL, xl, 1);   // exclusive: lanes above me
        if (lane == 31) { mr = 0; ml = 1; }
        sd_fracmax(mr, ml, cr, cl);
        const int tr = __shfl_sync(FULL, xr, 0), tl = __shfl_sync(FULL, xl, 0);
        sd_fracmax(cr, cl, tr, tl);
        bool ins = false;
        if (cand[b]) {
            const int i = 32 * b + lane, new_l = wn - i - 1;
            sd_fracmax(mr, ml, pr[b], pl[b]);
            if (nr[b] * ml >= mr * new_l) {
                SD_SLOT(si[b]) = sd_slot_pack(nr[b], new_l, wn + 2 - i);
                ins = !sv[b];
            }
        }
        fresh += __popc(__ballot_sync(FULL, ins));
    }
    if (lane == leader && fresh) {
        if (s.nslot == 0) sd_anchor_pstart(s, start, base);      // first slot after a stretch without any
        s.nslot += (int)fresh;
    }
    __syncwarp();
}

// one warp-task: 32 chunks (one per lane) stepped in lockstep
// NB = 32-position blocks covering the window (2 for W <= 66, else 4); COOP = warp-cooperative
// find_perfect (T >= 5; below that a candidate can have new_l == 0 and the serial form is used).
// Template parameters rather than run-time switches: the register allocation of the default
// W = 64 kernel is then not sized by the four-block arrays of the W = 128 one.
template <int NB, bool COOP, bool SSM>   // SSM: the perfect-interval slot rows live in shared memory (dense items), else in global memory
__device__ __forceinline__ void sdust_warp_task(const SdParams &P, uint32_t warp_id, uint32_t n_warps, uint32_t first_entry, uint32_t n_entries,
                                                uint32_t *smem, const SdLayout &lay, uint8_t *cnt, int lane)
{
    const uint32_t FULL = 0xffffffffu;
    // chunk of this lane: consecutive chunks go to DIFFERENT warp-tasks (lane l of task w takes chunk
    // l * n_warps + w).  Low-complexity stretches (satellites, telomeres) make their chunks many times
    // more expensive; spread over the tasks they cost each one slow lane instead of leaving one
    // warp with 32 of them as the kernel's tail.
    // (item mode: a task is 32 CONSECUTIVE entries of this launch's part of it_list)
    const uint32_t e = P.it_list ? warp_id * 32u + (uint32_t)lane : (uint32_t)lane * n_warps + warp_id;
    const bool have = e < n_entries;                  // lanes without a chunk still serve the warp's cooperative calls
    const uint32_t j = (have && P.it_list) ? P.it_list[first_entry + e] : e;      // chunk number, or item number in item mode
    const int T = P.T, W = P.W, cv_max = (P.T << 1) / 10;
    uint32_t *my_slots = P.gslots + (size_t)(have ? j : 0) * lay.slot_words;
    if (SSM) {
        // inside low-complexity sequence the slots are read and written at every step: ~60 dependent accesses per
        // find_perfect call, which from L2 made a step take ~20 k cycles.  One odd-strided row per thread, zeroed here
        // (the global rows were zeroed by a memset before the launch).
        my_slots = smem + (size_t)SD_BLOCK * (32 + lay.ring_words) + 16 * (SD_BLOCK / 32) + (size_t)threadIdx.x * lay.slot_words;
        for (int q = 0; q < lay.slot_words; ++q) my_slots[q] = 0;
    }
    const sd_mem m = sd_mem_of(smem, lay, threadIdx.x, my_slots);

    uint32_t rec = 0, k = 0;
    int len = 0, c0 = 0, c1 = 0;
    bool quiet = false;                               // item whose preceding block is quiet: window-only warm-up (P is empty at c0)
    if (have && P.it_list) {
        rec = P.it_rec[j];
        len = (int)P.rec_len[rec];
        c0 = (int)P.it_c0[j];
        c1 = (int)P.it_c1[j];
        quiet = (P.it_flags[j] & SD_ITEM_QUIET) != 0;
    } else if (have) {
        rec = corn_upper_bound(P.chunk_base, P.n_rec, j) - 1;
        k = j - P.chunk_base[rec];
        len = (int)P.rec_len[rec];
        c0 = (int)k * P.C;
        c1 = min(len, c0 + P.C);
    }
    (void)k;
    sd_sink sink;
    sd_sink_init(sink, P.slots + (size_t)(have ? j : 0) * P.cap, P.cap);
    DevFetch fetch;
    fetch.seq = P.seq + (have ? P.rec_off[rec] : 0);
    fetch.cblk = -1;
    fetch.codes = 0; fetch.valid = 0;

    sd_state s;
    sd_reset_counters(s, m);                          // (the slot rows were zeroed by a memset before the launch)
    int p0 = 0, n_steps = 0;
    if (have) {
        p0 = dev_warm_walk(fetch, quiet ? c0 : c0 - 2 * W, W) & ~15;     // = sd_warm_quiet / sd_warm_start
                                                      // (a longer warm-up is always valid) all lanes then refill their
                                                      // 16-byte fetch buffer on the same steps
        s.pstart = p0; s.pslot = (int)((uint32_t)p0 % (uint32_t)W);
        const int stop = c1 < len ? c1 : len;
        n_steps = (stop - p0) + (c1 >= len ? 1 : 0);  // + the reference's i == l_seq iteration for the record's last chunk
    }

    // `skipped` = positions of the warm-up that were jumped over: after the first byte of a run of non-ACGT bytes (which
    // flushes P and resets l, t) the following ones change nothing (:152-156), and a warm-up that starts in front of a
    // 50 kb gap would otherwise step through all of it with 31 lanes waiting
    int skipped = 0, cur_i = 0;
    for (int step = 0;; ++step) {
        const bool live = step + skipped < n_steps;
        if (!__any_sync(FULL, live)) break;
        bool emit = false, need_pop = false;
        int start = 0;
        if (live) {
            const int i = p0 + step + skipped;
            cur_i = i;
            if (i >= c0) sink.on = 1;
            const int b = i < len ? fetch.nt4(i) : 4;
            if (b >= 4) {
                int nx = i + 1;
                while (nx < c0 && fetch.nt4(nx) >= 4) ++nx;
                skipped += nx - (i + 1);
            }
            if (b < 4) {
                ++s.l;
                s.t = (s.t << 2 | (unsigned)b) & 63u;
                if (s.l >= 3) {
                    start = (s.l - W > 0 ? s.l - W : 0) + (i + 1 - s.l);
                    sd_save(s, m, sink, start, W);
                    need_pop = sd_shift_window_push(s, m, (int)s.t, T, cv_max, W);
                    emit = true;
                }
            } else {
                sd_flush(s, m, sink, (s.l - W + 1 > 0 ? s.l - W + 1 : 0) + (i + 1 - s.l), W);
                s.l = 0; s.t = 0;
            }
        }
        if (need_pop) sd_slack_unknown(s);                  // L is about to shrink: new suffixes become eligible
        uint32_t todo = __ballot_sync(FULL, need_pop);
        if (todo) {
            // Inside a tandem repeat the first element of v IS the triplet that just went over the limit, so the
            // shrink loop (sd_shift_window_pop) ends after one pop -- at every step.  SD_QUICK_POPS pops are tried per
            // lane, all lanes at once; only what is left (ordinary sequence needs ~10) goes to the cooperative form.
            // Measured at 3 Gb (Gbases/s, assembly / plain): 0 -> 55.7 / 87.5, 1 -> 66.6 / 92.9, 4 -> 67.2 / 96.7, 8 -> 66.2 / 97.9.
#pragma unroll
            for (int q = 0; q < SD_QUICK_POPS; ++q) {
                if (need_pop) {
                    const int x = SD_RING(sd_ring_idx(s, s.wn - s.L, W));
                    const int e2 = SD_U8(m.cv, x) - 1;
                    SD_U8(m.cv, x) = (uint8_t)e2;
                    s.rv -= e2;
                    --s.L;
                    need_pop = x != (int)s.t;
                }
            }
            todo = __ballot_sync(FULL, need_pop);
        }
        while (todo) {
            const int leader = __ffs(todo) - 1;
            todo &= todo - 1;
            pop_coop<NB>(leader, lane, s, (int)s.t, smem, lay, W);
        }
        bool trig = false;
        if (emit && s.rw * 10 > s.L * T && !(quiet && cur_i < c0)) {         // (quiet warm-up: the true run calls nothing here)
            if (!COOP) sd_find_perfect(s, m, T, start, W);
            else trig = s.wn - s.L - 1 >= 0 && s.slack < 0;   // no index to examine / provably no candidate otherwise
        }
        todo = __ballot_sync(FULL, trig);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            todo &= todo - 1;
            if (COOP) fp_coop<NB>(leader, lane, s, start, smem, lay, my_slots, cnt, T, W);
        }
    }
    sd_sink_close(sink);
    if (have) {
        P.cnt[j] = sink.n;
        if (sink.overflow) atomicAdd(P.err, 1u);
    }
}

// ---- cost-aware task order -----------------------------------------------------------------------
// A chunk inside a tandem repeat (telomere, satellite) needs both cooperative routines at every
// step, which makes its whole warp-task 3-4x as long as an ordinary one.  If such a task is claimed
// late, it is the tail of the kernel.  A probe looks at 64 bases in the middle of every chunk: few
// distinct triplets => expensive.  Tasks with an expensive lane are put at the front of the ticket
// order.  (Purely a scheduling hint: results do not depend on it.)
__global__ void __launch_bounds__(256) k_sdust_probe(const SdParams P, uint8_t *flag)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n_chunks) return;
    const uint32_t rec = corn_upper_bound(P.chunk_base, P.n_rec, j) - 1;
    const uint32_t k = j - P.chunk_base[rec];
    const uint32_t len = P.rec_len[rec];
    const uint32_t c0 = k * (uint32_t)P.C, c1 = min(len, c0 + (uint32_t)P.C);
    uint32_t a = (c0 + (c1 - c0) / 2) & ~15u;                       // 64 bytes around the middle, 16-byte aligned
    if (a + 64 > c1) a = c1 >= 64 ? (c1 - 64) & ~15u : 0;
    const uint4 *p = (const uint4 *)(P.seq + P.rec_off[rec] + a);
    uint64_t seen = 0;
    uint32_t t = 0, run = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 v = __ldg(p + q);
        const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int b = sd_nt4((uint8_t)(w[i >> 2] >> (8 * (i & 3))));
            if (b < 4) { t = (t << 2 | (uint32_t)b) & 63u; if (++run >= 3) seen |= 1ull << t; }
            else run = 0;
        }
    }
    flag[j] = (uint8_t)(__popcll(seen) <= 24);                       // random DNA shows ~40 distinct triplets in 62
}

__global__ void __launch_bounds__(256) k_sdust_order(const uint8_t *__restrict__ flag, uint32_t n_chunks, uint32_t n_warps,
                                                     uint32_t *task_list, uint32_t *counters /* [0] front, [1] back */)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_warps) return;
    bool heavy = false;
    for (uint32_t l = 0; l < 32; ++l) {
        const uint32_t j = l * n_warps + w;
        if (j < n_chunks && flag[j]) { heavy = true; break; }
    }
    if (heavy) task_list[atomicAdd(&counters[0], 1u)] = w;
    else       task_list[n_warps - 1u - atomicAdd(&counters[1], 1u)] = w;
}

// Persistent grid: warps claim warp-tasks from a counter, so the kernel ends one task -- not one wave
// of blocks -- after the last claim.
template <int NB, bool COOP, bool SSM = false>
__global__ void __launch_bounds__(SD_BLOCK, (NB == 2 && COOP) ? 8 : (SSM ? 3 : 5)) k_sdust_scan(const SdParams P)
{
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31;
    const SdLayout lay(P.W);
    uint8_t *cnt = (uint8_t *)(smem + (size_t)SD_BLOCK * (32 + lay.ring_words)) + 64 * (threadIdx.x >> 5);
    uint32_t first_entry = 0, n_entries = P.n_chunks;
    if (P.it_list) { first_entry = *P.list_lo; n_entries = *P.list_hi - first_entry; }
    const uint32_t n_warps = (n_entries + 31u) / 32u;
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(P.task_counter, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_warps) break;
        sdust_warp_task<NB, COOP, SSM>(P, P.task_list ? P.task_list[w] : w, n_warps, first_entry, n_entries, smem, lay, cnt, lane);
        __syncwarp();
    }
}

// =====================================================================================================
// Two-phase execution (sdust_core.cuh, "Two-phase execution"): scout -> items -> full machine on the items only
// =====================================================================================================
constexpr int SC_BLOCK = 128;

struct ScoutParams {
    const uint8_t  *seq;
    const uint32_t *rec_off, *rec_len;
    const uint32_t *chunk_base;     // [n_rec+1] scout chunks (C bases each) before every record
    const uint32_t *blk_base;       // [n_rec+1] 64-base blocks before every record
    uint32_t n_rec, n_chunks;
    int T, W, C;
    uint8_t *active;                // [n_blk_total + 1], zeroed
    uint8_t *tcnt;                  // [n_blk_total + 1], zeroed: trigger positions per block (written by the block's own chunk only)
    uint8_t *nflag;                 // [n_blk_total + 1], zeroed: the block holds a non-ACGT byte
};

// shared-memory accessors of the scout: one 32-bit word per triplet value as a column (word index = row * SC_BLOCK +
// thread: whatever a lane indexes stays in its own bank) and a 64-entry byte ring per thread as a plain row with an odd
// word stride (one add per access; lanes that index the same entry -- the usual case -- hit 32 different banks)
constexpr int SC_RING_STRIDE = 68;
struct ScoutWords {
    uint32_t *col;
    __device__ __forceinline__ uint32_t &operator()(uint32_t i) { return col[i * SC_BLOCK]; }
};
struct ScoutRing {
    uint8_t *row;
    __device__ __forceinline__ uint8_t &operator()(uint32_t i) { return row[i]; }
};

// Phase 1.  One thread per chunk of C bases: the window half of the state machine in its loop-free form; wherever the
// reference would call find_perfect, the 64-base block of that position and the next one (the drain) are marked active.
__global__ void __launch_bounds__(SC_BLOCK) k_sdust_scout(const ScoutParams P)
{
    __shared__ uint32_t sm_words[64 * SC_BLOCK];
    __shared__ uint32_t sm_ring[SC_RING_STRIDE / 4 * SC_BLOCK];
    const uint32_t g = blockIdx.x * SC_BLOCK + threadIdx.x;
    const bool have = g < P.n_chunks;
    ScoutWords words = { sm_words + threadIdx.x };
    ScoutRing ring = { (uint8_t *)sm_ring + threadIdx.x * SC_RING_STRIDE };
    sd_scout sc;
    sd_scout_reset(sc, words, ring);
    uint32_t rec = 0;
    int len = 0, c0 = 0, c1 = 0, p0 = 0;
    const uint8_t *seq = P.seq;
    if (have) {
        rec = corn_upper_bound(P.chunk_base, P.n_rec, g) - 1;
        len = (int)P.rec_len[rec];
        c0 = (int)(g - P.chunk_base[rec]) * P.C;
        c1 = min(len, c0 + P.C);
        seq = P.seq + P.rec_off[rec];
        DevFetch fetch;
        fetch.seq = seq; fetch.cblk = -1; fetch.codes = 0; fetch.valid = 0;
        p0 = dev_warm_walk(fetch, c0, P.W) & ~15;            // = sd_warm_quiet
    }
    uint8_t *act = P.active + (have ? P.blk_base[rec] : 0);
    uint8_t *tcn = P.tcnt + (have ? P.blk_base[rec] : 0);
    uint8_t *nfl = P.nflag + (have ? P.blk_base[rec] : 0);
    const int n_blk = (len + SD_BLK - 1) / SD_BLK;
    const int T = P.T, W = P.W;
    // sixteen positions per 16-byte load, decoded at once (nt4x4) and stepped through with compile-time shifts.  Every
    // lane runs (nearly) the same number of steps and no step has a data-dependent loop: the warp stays converged.
    for (int ib = p0; have && ib < c1; ib += 16) {
        const uint4 v = __ldg((const uint4 *)(seq + ib));
        uint32_t v0, v1, v2, v3;
        const uint32_t codes = nt4x4(v.x, &v0) | (nt4x4(v.y, &v1) << 8) | (nt4x4(v.z, &v2) << 16) | (nt4x4(v.w, &v3) << 24);
        const uint32_t valid = v0 | (v1 << 4) | (v2 << 8) | (v3 << 12);
        const int rem = c1 - ib;                              // positions of this group inside the chunk
        const bool own = ib >= c0;                            // (c0 is a multiple of 64, p0 of 16)
        const int k = ib >> 6;
        const uint32_t real = rem < 16 ? (1u << rem) - 1u : 0xFFFFu;
        if ((valid & real) != real) nfl[k] = 1;               // a non-ACGT byte in this block (any thread may say so)
        if ((valid & real) == 0u) { sc.l = 0; sc.t = 0; continue; }      // sixteen non-ACGT bytes (inside a gap): nothing else to do
        uint32_t ntrig = 0;
        bool stale = false;                                   // a trigger less than W bases after a non-ACGT byte
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if (q < rem) {
                if ((valid >> q) & 1u) {
                    ++sc.l;
                    sc.t = (sc.t << 2 | ((codes >> (2 * q)) & 3u)) & 63u;
                    if (sc.l >= 3 && sd_scout_push(sc, words, ring, sc.t, T, W)) { ++ntrig; stale |= sc.l < W; }
                } else { sc.l = 0; sc.t = 0; }
            }
        }
        if (ntrig && own) {                                   // (a 16-base group lies inside one 64-base block)
            act[k] = 1;
            if (k + 1 < n_blk) act[k + 1] = 1;
            if (stale && k + 2 < n_blk) act[k + 2] = 1;       // the longer drain of the stale phase (sdust_core.cuh)
            tcn[k] = (uint8_t)(tcn[k] + ntrig);               // (chunks are multiples of 64 bases: block k is this thread's alone)
        }
    }
}

// Items.  Block j starts an item iff it is active and (the block before it in its record is not, or j lies on the
// SD_ITEM_MAX grid of its record).  start[] -> exclusive scan -> k_sdust_items writes the item table in position order.
struct ItemParams {
    const uint8_t  *active, *tcnt, *nflag;
    const uint32_t *blk_base, *rec_len;
    uint32_t n_rec, n_blk;
    uint32_t *start;                // [n_blk]: 1 where an item starts; scanned in place to the item number
    const uint32_t *total;          // number of items (device)
    uint32_t cap_items;
    uint32_t *it_rec, *it_c0, *it_c1, *it_flags;
    uint32_t dense_pct;             // an item is dense when at least this share (%) of its positions are trigger positions
    uint32_t *it_dense;             // [n_items] 1 = dense item (k_sdust_items); scanned to the item's rank in its class
};

__device__ __forceinline__ bool item_starts_at(const uint8_t *__restrict__ active, uint32_t j, uint32_t first_blk_of_rec)
{
    if (!active[j]) return false;
    const uint32_t k = j - first_blk_of_rec;
    return k == 0 || !active[j - 1] || (k % (SD_ITEM_MAX / SD_BLK)) == 0;
}

__global__ void __launch_bounds__(256) k_sdust_item_starts(const ItemParams P)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n_blk) return;
    uint32_t v = 0;
    if (P.active[j]) {
        const uint32_t rec = corn_upper_bound(P.blk_base, P.n_rec, j) - 1;
        v = item_starts_at(P.active, j, P.blk_base[rec]) ? 1u : 0u;
    }
    P.start[j] = v;
}

__global__ void __launch_bounds__(256) k_sdust_items(const ItemParams P, const uint32_t *__restrict__ item_no)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    bool have = j < P.n_blk && P.active[j];
    uint32_t rec = 0, b0 = 0, it = 0;
    if (have) {
        rec = corn_upper_bound(P.blk_base, P.n_rec, j) - 1;
        b0 = P.blk_base[rec];
        have = item_starts_at(P.active, j, b0);
    }
    if (have) { it = item_no[j]; have = it < P.cap_items; }            // (beyond the tables: host grows them and repeats)
    if (!have) return;
    bool dense = false;
    const uint32_t len = P.rec_len[rec], k = j - b0, nb = (len + SD_BLK - 1) / SD_BLK;
    uint32_t e = k + 1;                                                // the item runs to the next cut or to the end of the active run
    while (e < nb && P.active[b0 + e] && (e % (SD_ITEM_MAX / SD_BLK)) != 0) ++e;
    const bool prev = k > 0 && P.active[j - 1];
    P.it_rec[it] = rec;
    P.it_c0[it] = k * SD_BLK;
    P.it_c1[it] = min(len, e * SD_BLK);
    P.it_flags[it] = (prev ? SD_ITEM_CHAIN : 0u) | ((k > 0 && !prev) ? SD_ITEM_QUIET : 0u);
    // Two classes.  DENSE items lie inside low-complexity sequence (a microsatellite, a telomere): find_perfect runs at
    // (nearly) every step there; one warp runs one such item with the window spread over its lanes (k_sdust_dense).
    // SPARSE items are isolated bursts on ordinary sequence -- a call every few dozen steps -- and 32 of them share a
    // warp, one per lane, with the cooperative routines serving whichever lane needs them (k_sdust_scan<2, true>).
    uint32_t trig = 0;
    for (uint32_t b = k; b < e; ++b) trig += P.tcnt[b0 + b];
    dense = trig * 100u >= (e - k) * SD_BLK * P.dense_pct;
    // the dense kernel only takes items without a non-ACGT byte from their warm start (at most 3W + 2 = 194 bases, i.e.
    // four blocks, before c0) to their end
    for (uint32_t b = k >= 4u ? k - 4u : 0u; dense && b < e; ++b) dense = P.nflag[b0 + b] == 0;
    P.it_dense[it] = dense ? 1u : 0u;
}

// item numbers by class, each class in position order: dense items are it_list[0 .. *n_dense), sparse ones follow.
// rank = exclusive scan of it_dense.  (Placing them with one atomicAdd per warp and class on two counters cost 2.2 ms of a
// 3 Gb batch: ~1.5 M returning atomics on two addresses.)
__global__ void __launch_bounds__(256) k_sdust_item_lists(const uint32_t *__restrict__ it_dense, const uint32_t *__restrict__ rank,
                                                           uint32_t n_items, const uint32_t *__restrict__ n_dense, uint32_t *__restrict__ it_list)
{
    const uint32_t it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= n_items) return;
    const uint32_t r = rank[it];
    it_list[it_dense[it] ? r : *n_dense + (it - r)] = it;
}

struct ItemGatherParams {
    const uint64_t *slots;
    const uint32_t *cnt, *it_c0, *it_flags, *it_rec;
    const uint32_t *n_items;        // device
    uint32_t cap;
    int W;
    uint32_t *out_cnt;
    const uint32_t *out_off;
    uint64_t *out;
};

// it_off[] of the fold routines: every item owns `cap` consecutive slots
struct FixedOff { uint32_t cap; };

__global__ void __launch_bounds__(256) k_sdust_item_gather(const ItemGatherParams P, const uint32_t *__restrict__ off, int write)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = *P.n_items;
    if (j >= n) return;
    if (!write) { P.out_cnt[j] = sd_item_gather_count(P.slots, off, P.cnt, P.it_c0, P.it_flags, j, P.W); return; }
    sd_item_gather_write(P.slots, off, P.cnt, P.it_c0, P.it_flags, j, n, P.W, P.out + P.out_off[j]);
}

__global__ void __launch_bounds__(256) k_sdust_item_off(uint32_t *off, uint32_t n, uint32_t cap)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) off[j] = j * cap;
}

// rec_first[r] = offset of the first interval of the first item of a record >= r
__global__ void k_sdust_item_rec_first(const uint32_t *__restrict__ it_rec, const uint32_t *__restrict__ out_off, const uint32_t *__restrict__ n_items,
                                       const uint32_t *__restrict__ total, uint32_t n_rec, uint64_t *rec_first)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rec) return;
    const uint32_t n = *n_items;
    const uint32_t it = corn_lower_bound(it_rec, n, r);               // first item whose record is >= r
    rec_first[r] = it < n ? out_off[it] : *total;
}

// -------------------------------------------------------------------------------------------------------
// DENSE items: one warp per item, the window spread over the lanes in AGE order (lane = age, two ages per lane).
// This is sd_run_item_dense() of sdust_core.cuh (the plain-array statement, checked against the oracle in tests/sim)
// with the loops over ages turned into ballots, shuffles and warp scans.  Only items without a non-ACGT byte in
// [p0, c1) come here (k_sdust_items checks the scout's per-block flags).
// -------------------------------------------------------------------------------------------------------
struct DenseParams {
    const uint8_t  *seq;
    const uint32_t *rec_off, *rec_len;
    const uint32_t *it_list, *it_rec, *it_c0, *it_c1, *it_flags;
    const uint32_t *n_dense;        // device: the dense items are it_list[0 .. *n_dense)
    int T, W;
    uint32_t cap;
    uint64_t *slots;                // item number * cap
    uint32_t *cnt, *err, *task_counter;
};

__device__ __forceinline__ void frac_scan_step(int &r, int &l, int o, int lane)
{
    const int yr = __shfl_up_sync(0xffffffffu, r, o), yl = __shfl_up_sync(0xffffffffu, l, o);
    if (lane >= o) sd_fracmax(r, l, yr, yl);
}

__global__ void __launch_bounds__(128) k_sdust_dense(const DenseParams P)
{
    const uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t n_dense = *P.n_dense;
    const int T = P.T, W = P.W;
    const uint32_t le = (2u << lane) - 1u;            // ages 0..lane of a block
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(P.task_counter, 1u);
        w = __shfl_sync(FULL, w, 0);
        if (w >= n_dense) break;
        const uint32_t it = P.it_list[w];
        const uint32_t rec = P.it_rec[it];
        const int len = (int)P.rec_len[rec], c0 = (int)P.it_c0[it], c1 = (int)P.it_c1[it];
        const bool quiet = (P.it_flags[it] & SD_ITEM_QUIET) != 0;
        int p0 = quiet ? c0 - (W + 2) : c0 - 2 * W - (W + 2);
        if (p0 < 0 || (!quiet && c0 - 2 * W <= 0)) p0 = 0;
        const uint8_t *seq = P.seq + P.rec_off[rec];
        sd_sink sink;
        sd_sink_init(sink, P.slots + (size_t)it * P.cap, P.cap);

        int x0 = 0xFF, x1 = 0xFF, R0 = 0, R1 = 0;    // per lane: ages lane and 32 + lane
        uint32_t s0 = 0, s1 = 0;
        int wn = 0, L = 0, rw = 0;
        unsigned t = 0;
        const int stop = c1 < len ? c1 : len;
        for (int ib = p0; ib < stop; ib += 32) {
            const int mine = ib + lane < stop ? sd_nt4(__ldg(seq + ib + lane)) : 0;
            const int n_here = min(32, stop - ib);
            for (int q = 0; q < n_here; ++q) {
                const int i = ib + q;
                if (i == c0) sink.on = 1;
                t = (t << 2 | (unsigned)__shfl_sync(FULL, mine, q)) & 63u;
                const int l = i - p0 + 1;
                if (l < 3) continue;
                const int start = p0 + (l - W > 0 ? l - W : 0);
                if (wn >= W - 2) {                    // the oldest element leaves; its interval, if any, is saved
                    const int o = wn - 1;
                    const uint32_t so = __shfl_sync(FULL, o < 32 ? s0 : s1, o & 31);
                    const int xo = __shfl_sync(FULL, o < 32 ? x0 : x1, o & 31);
                    if (so & SD_SLOT_VALID) sd_sink_put(sink, start - 1, start - 1 + sd_slot_flen(so));
                    rw -= __popc(__ballot_sync(FULL, x0 == xo)) + __popc(__ballot_sync(FULL, x1 == xo)) - 1;
                    if (lane == (o & 31)) { if (o < 32) { x0 = 0xFF; s0 = 0; } else { x1 = 0xFF; s1 = 0; } }
                    --wn;
                }
                if (L > wn) L = wn;
                // every element is one step older
                {
                    const int cx = __shfl_sync(FULL, x0, 31), cR = __shfl_sync(FULL, R0, 31);
                    const uint32_t cs = __shfl_sync(FULL, s0, 31);
                    x1 = __shfl_up_sync(FULL, x1, 1); R1 = __shfl_up_sync(FULL, R1, 1); s1 = __shfl_up_sync(FULL, s1, 1);
                    x0 = __shfl_up_sync(FULL, x0, 1); R0 = __shfl_up_sync(FULL, R0, 1); s0 = __shfl_up_sync(FULL, s0, 1);
                    if (lane == 0) { x1 = cx; R1 = cR; s1 = cs; x0 = 0xFF; }
                }
                const uint32_t m0 = __ballot_sync(FULL, x0 == (int)t), m1 = __ballot_sync(FULL, x1 == (int)t);
                const int c0t = __popc(m0), cnt = c0t + __popc(m1);
                R0 += __popc(m0 & le);
                R1 += c0t + __popc(m1 & le);
                if (lane == 0) { x0 = (int)t; R0 = 0; s0 = 0; }
                rw += cnt;
                ++wn;
                ++L;
                if (cnt >= 4) {
                    const int d4 = c0t >= 4 ? (int)__fns(m0, 0, 4) : 32 + (int)__fns(m1, 0, 4 - c0t);
                    if (d4 < L) L = d4;
                }
                if (rw * 10 > L * T && L < wn && !(quiet && i < c0)) {
                    // find_perfect in age order: exclusive running maximum of (slot, candidate) ratios over younger ages
                    const int j0 = lane, j1 = 32 + lane;
                    const bool v0 = (s0 & SD_SLOT_VALID) != 0, v1 = (s1 & SD_SLOT_VALID) != 0;
                    const int pr0 = v0 ? sd_slot_r(s0) : 0, pl0 = v0 ? sd_slot_l(s0) : 1;
                    const int pr1 = v1 ? sd_slot_r(s1) : 0, pl1 = v1 ? sd_slot_l(s1) : 1;
                    const bool cand0 = j0 < wn && j0 >= L && R0 * 10 > T * j0;
                    const bool cand1 = j1 < wn && j1 >= L && R1 * 10 > T * j1;
                    int er0 = pr0, el0 = pl0, er1 = pr1, el1 = pl1;
                    if (cand0) sd_fracmax(er0, el0, R0, j0);
                    if (cand1) sd_fracmax(er1, el1, R1, j1);
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { frac_scan_step(er0, el0, o, lane); frac_scan_step(er1, el1, o, lane); }
                    const int t0r = __shfl_sync(FULL, er0, 31), t0l = __shfl_sync(FULL, el0, 31);     // all of block 0
                    int xr0 = __shfl_up_sync(FULL, er0, 1), xl0 = __shfl_up_sync(FULL, el0, 1);
                    int xr1 = __shfl_up_sync(FULL, er1, 1), xl1 = __shfl_up_sync(FULL, el1, 1);
                    if (lane == 0) { xr0 = 0; xl0 = 1; xr1 = 0; xl1 = 1; }
                    sd_fracmax(xr1, xl1, t0r, t0l);
                    if (cand0) { sd_fracmax(xr0, xl0, pr0, pl0); if (R0 * xl0 >= xr0 * j0) s0 = sd_slot_pack(R0, j0, j0 + 3); }
                    if (cand1) { sd_fracmax(xr1, xl1, pr1, pl1); if (R1 * xl1 >= xr1 * j1) s1 = sd_slot_pack(R1, j1, j1 + 3); }
                }
            }
        }
        if (c1 >= len) {                              // the record ends here: flush (:152-154)
            sink.on = 1;
            const int l = len - p0;
            if (l >= 3) {
                int start = (l - W + 1 > 0 ? l - W + 1 : 0) + (len + 1 - l);
                const int a0 = len - 3;
                for (;;) {
                    const uint32_t vm0 = __ballot_sync(FULL, (s0 & SD_SLOT_VALID) != 0), vm1 = __ballot_sync(FULL, (s1 & SD_SLOT_VALID) != 0);
                    if (!(vm0 | vm1)) break;
                    const int oldest = vm1 ? 63 - __clz(vm1) : 31 - __clz(vm0);
                    const int a = a0 - oldest;
                    if (a >= start) start = a + 1;    // (the rounds in between find nothing below their start)
                    const uint32_t so = __shfl_sync(FULL, oldest < 32 ? s0 : s1, oldest & 31);
                    sd_sink_put(sink, a, a + sd_slot_flen(so));
                    if (a0 - lane < start) s0 = 0;
                    if (a0 - 32 - lane < start) s1 = 0;
                    ++start;
                }
            }
        }
        sd_sink_close(sink);
        if (lane == 0) {
            P.cnt[it] = sink.n;
            if (sink.overflow) atomicAdd(P.err, 1u);
        }
        __syncwarp();
    }
}

struct GatherParams {
    const uint64_t *slots;
    const uint32_t *cnt;
    const uint32_t *chunk_base, *rec_len;
    uint32_t n_rec, n_chunks, cap;
    int C, W;
    uint32_t *out_cnt;
    const uint32_t *out_off;
    uint64_t *out;
};

__global__ void __launch_bounds__(256) k_sdust_gather(const GatherParams P, int write)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n_chunks) return;
    const uint32_t rec = corn_upper_bound(P.chunk_base, P.n_rec, j) - 1;
    const uint32_t k = j - P.chunk_base[rec];
    if (!write) { P.out_cnt[j] = sd_gather_count(P.slots, P.cnt, P.cap, j, k, P.C, P.W); return; }
    const uint32_t nch = P.chunk_base[rec + 1] - P.chunk_base[rec];
    sd_gather_write(P.slots, P.cnt, P.cap, j, k, nch, P.C, P.W, P.out + P.out_off[j]);
}

// rec_first[r] = offset of the first interval of record r; rec_first[n_rec] = total
__global__ void k_sdust_rec_first(const uint32_t *__restrict__ chunk_base, const uint32_t *__restrict__ out_off,
                                  const uint32_t *__restrict__ total, uint32_t n_rec, uint32_t n_chunks, uint64_t *rec_first)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rec) return;
    const uint32_t cb = r < n_rec ? chunk_base[r] : n_chunks;
    rec_first[r] = cb < n_chunks ? out_off[cb] : *total;
}

}  // namespace

// --------------------------------------------------------------------------------------------
// two-phase path (W <= 64, floor(2T/10) == 4: the defaults)
// --------------------------------------------------------------------------------------------
static int sdust_run_fast(corn_ctx *ctx, const corn_dbatch *db, int T, int W, corn_intervals_t *out)
{
    cudaStream_t st = ctx->stream;
    const uint32_t n_rec = db->n_rec;
    uint64_t *h_first = (uint64_t *)corn_host_alloc(sizeof(uint64_t) * ((size_t)n_rec + 1));
    if (!h_first) return corn_set_err(ctx, CORN_E_NOMEM, "pinned alloc");
    out->rec_first = h_first;
    for (uint32_t r = 0; r <= n_rec; ++r) h_first[r] = 0;

    // scout chunk length: every chunk pays ~W + 16 warm-up positions; 2048 keeps that at 4 % and still gives a 3 Gb
    // batch 1.5 M threads
    int C = 2048;
    if (const char *e = getenv("CORNETTO_SDUST_SCOUT_CHUNK")) { int v = atoi(e); if (v >= 64 && v <= (1 << 20)) C = v / 64 * 64; }
    uint64_t n_chunks64 = 0, n_blk64 = 0;
    for (uint32_t r = 0; r < n_rec; ++r) {
        n_chunks64 += ((uint64_t)db->h_rec_len[r] + C - 1) / C;
        n_blk64 += ((uint64_t)db->h_rec_len[r] + SD_BLK - 1) / SD_BLK;
    }
    if (n_chunks64 == 0) return CORN_OK;
    if (n_blk64 > 0xFFFFFFF0ull) return corn_set_err(ctx, CORN_E_TOOBIG, "too many sdust blocks");
    const uint32_t n_chunks = (uint32_t)n_chunks64, n_blk = (uint32_t)n_blk64;

    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->misc, 4096));
    uint32_t *d_tot = (uint32_t *)((uint8_t *)ctx->misc.p + 2048);   // [0] intervals, [1] items, [4] overflow errors, [8] task counter, [11] dense items, [12] dense task counter
    uint32_t *d_err = d_tot + 4;
    CORN_CUDA(ctx, cudaMemsetAsync(d_tot, 0, 64, st));

    // tables: nch | chunk_base | nblk | blk_base (n_rec + 1 each), then start / item_no [n_blk + 1], then active bytes
    const size_t tab_words = 4 * ((size_t)n_rec + 1) + ((size_t)n_blk + 2) + 3 * (((size_t)n_blk + 8) / 4 + 2) + 64;
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->sd_tab, tab_words * sizeof(uint32_t)));
    uint32_t *nch = (uint32_t *)ctx->sd_tab.p, *chunk_base = nch + n_rec + 1, *nblk = chunk_base + n_rec + 1, *blk_base = nblk + n_rec + 1;
    uint32_t *item_no = blk_base + n_rec + 1;
    uint8_t *active = (uint8_t *)(item_no + n_blk + 2);
    const size_t act_bytes = (((size_t)n_blk + 8) / 4 + 2) * 4;
    uint8_t *tcnt = active + act_bytes, *nflag = tcnt + act_bytes;

    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    k_sdust_nchunks<<<(n_rec + 255) / 256, 256, 0, st>>>(db->d_rec_len, nch, n_rec, (uint32_t)C);
    k_sdust_nchunks<<<(n_rec + 255) / 256, 256, 0, st>>>(db->d_rec_len, nblk, n_rec, (uint32_t)SD_BLK);
    corn_count_launch(ctx, 2);
    CORN_TRY(corn_scan_u32(ctx, nch, chunk_base, n_rec, chunk_base + n_rec));
    CORN_TRY(corn_scan_u32(ctx, nblk, blk_base, n_rec, blk_base + n_rec));
    CORN_CUDA(ctx, cudaMemsetAsync(active, 0, 3 * act_bytes, st));

    // ---- phase 1: scout ----
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    ScoutParams sc;
    sc.seq = db->d_seq; sc.rec_off = db->d_rec_off; sc.rec_len = db->d_rec_len; sc.chunk_base = chunk_base; sc.blk_base = blk_base;
    sc.n_rec = n_rec; sc.n_chunks = n_chunks; sc.T = T; sc.W = W; sc.C = C; sc.active = active; sc.tcnt = tcnt; sc.nflag = nflag;
    k_sdust_scout<<<(n_chunks + SC_BLOCK - 1) / SC_BLOCK, SC_BLOCK, 0, st>>>(sc);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[12], st));

    // ---- items ----
    ItemParams ip;
    memset(&ip, 0, sizeof ip);
    ip.active = active; ip.tcnt = tcnt; ip.nflag = nflag; ip.blk_base = blk_base; ip.rec_len = db->d_rec_len; ip.n_rec = n_rec; ip.n_blk = n_blk;
    ip.start = item_no; ip.total = d_tot + 1;
    ip.dense_pct = 25;
    if (const char *e = getenv("CORNETTO_SDUST_DENSE_PCT")) { const int v = atoi(e); if (v >= 1 && v <= 1000) ip.dense_pct = (uint32_t)v; }
    k_sdust_item_starts<<<(n_blk + 255) / 256, 256, 0, st>>>(ip);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    CORN_TRY(corn_scan_u32(ctx, item_no, item_no, n_blk, d_tot + 1));
    uint32_t hv[16];
    CORN_TRY(corn_read_small(ctx, hv, d_tot, 64));
    const uint32_t n_items = hv[1];
    if (n_items == 0) {                                  // nothing anywhere that could be masked
        CORN_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
        CORN_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
        CORN_CUDA(ctx, cudaEventRecord(ctx->ev[6], st));
        CORN_CUDA(ctx, cudaStreamSynchronize(st));
        cudaEventElapsedTime(&ctx->timing.scan_ms, ctx->ev[3], ctx->ev[4]);
        return CORN_OK;
    }
    const uint32_t cap = (uint32_t)(SD_ITEM_MAX + 2 * W) / 4 + 2;
    if ((uint64_t)n_items * cap > 0xFFFFFFF0ull) {       // (tens of millions of tiny items: slot offsets are 32-bit) the chunk grid takes over
        corn_host_free(h_first);
        out->rec_first = NULL;
        return CORN_E_STATE;
    }
    const SdLayout lay(W);
    const size_t slot_words = (size_t)lay.slot_words;
    // item tables: rec | c0 | c1 | flags | list | off | cnt | out_cnt | out_off  (n_items + 1 each)
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->sd_out, 9 * ((size_t)n_items + 1) * sizeof(uint32_t)));
    uint32_t *it_rec = (uint32_t *)ctx->sd_out.p, *it_c0 = it_rec + n_items + 1, *it_c1 = it_c0 + n_items + 1, *it_flags = it_c1 + n_items + 1;
    uint32_t *it_list = it_flags + n_items + 1, *it_off = it_list + n_items + 1, *cnt = it_off + n_items + 1;
    uint32_t *out_cnt = cnt + n_items + 1, *out_off = out_cnt + n_items + 1;
    const size_t iv_bytes = ((size_t)n_items * cap * sizeof(uint64_t) + 255) & ~(size_t)255;
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->sd_slots, iv_bytes + (size_t)n_items * slot_words * sizeof(uint32_t)));
    uint64_t *slots = (uint64_t *)ctx->sd_slots.p;
    uint32_t *gslots = (uint32_t *)((uint8_t *)ctx->sd_slots.p + iv_bytes);
    CORN_CUDA(ctx, cudaMemsetAsync(gslots, 0, (size_t)n_items * slot_words * sizeof(uint32_t), st));
    ip.cap_items = n_items;
    ip.it_rec = it_rec; ip.it_c0 = it_c0; ip.it_c1 = it_c1; ip.it_flags = it_flags; ip.it_dense = out_cnt;          // (out_cnt / out_off are free until the gather)
    k_sdust_items<<<(n_blk + 255) / 256, 256, 0, st>>>(ip, item_no);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    CORN_TRY(corn_scan_u32(ctx, out_cnt, out_off, n_items, d_tot + 11));
    k_sdust_item_lists<<<(n_items + 255) / 256, 256, 0, st>>>(out_cnt, out_off, n_items, d_tot + 11, it_list);
    k_sdust_item_off<<<(n_items + 255) / 256, 256, 0, st>>>(it_off, n_items, cap);
    corn_count_launch(ctx, 2);
    CORN_LAUNCH_CHECK(ctx);

    // ---- phase 2: the full machine on the items ----
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[13], st));
    const size_t smem = lay.bytes();
    typedef void (*sd_kernel_t)(const SdParams);
    SdParams sp;
    memset(&sp, 0, sizeof sp);
    sp.seq = db->d_seq; sp.rec_off = db->d_rec_off; sp.rec_len = db->d_rec_len; sp.chunk_base = chunk_base;
    sp.n_rec = n_rec; sp.n_chunks = n_items; sp.T = T; sp.W = W; sp.C = SD_ITEM_MAX; sp.cap = cap;
    sp.slots = slots; sp.gslots = gslots; sp.cnt = cnt; sp.err = d_err; sp.task_list = NULL;
    sp.it_list = it_list; sp.it_rec = it_rec; sp.it_c0 = it_c0; sp.it_c1 = it_c1; sp.it_flags = it_flags;
    // d_tot: [1] items, [11] dense items (they fill it_list from the front).  The two classes are independent and could
    // share the SMs from two streams ($CORNETTO_SDUST_OVERLAP=1: sparse 6 of its 8 possible blocks per SM, dense 3), but
    // both kernels are issue bound: measured at 3 Gb, 25.1 ms together against 12.4 + 10.1 ms one after the other.
    const bool overlap = getenv("CORNETTO_SDUST_OVERLAP") && atoi(getenv("CORNETTO_SDUST_OVERLAP")) > 0;
    if (overlap && !ctx->aux_stream) {
        CORN_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
        CORN_CUDA(ctx, cudaEventCreateWithFlags(&ctx->aux_ev[0], cudaEventDisableTiming));
        CORN_CUDA(ctx, cudaEventCreateWithFlags(&ctx->aux_ev[1], cudaEventDisableTiming));
    }
    cudaStream_t st_dense = overlap ? ctx->aux_stream : st;
    if (overlap) {
        CORN_CUDA(ctx, cudaEventRecord(ctx->aux_ev[0], st));
        CORN_CUDA(ctx, cudaStreamWaitEvent(st_dense, ctx->aux_ev[0], 0));
    }
    {
        DenseParams dp;
        dp.seq = db->d_seq; dp.rec_off = db->d_rec_off; dp.rec_len = db->d_rec_len;
        dp.it_list = it_list; dp.it_rec = it_rec; dp.it_c0 = it_c0; dp.it_c1 = it_c1; dp.it_flags = it_flags;
        dp.n_dense = d_tot + 11; dp.T = T; dp.W = W; dp.cap = cap; dp.slots = slots; dp.cnt = cnt; dp.err = d_err; dp.task_counter = d_tot + 12;
        k_sdust_dense<<<ctx->sm_count * (overlap ? 3 : 8), 128, 0, st_dense>>>(dp);
        corn_count_launch(ctx);
        CORN_LAUNCH_CHECK(ctx);
        if (overlap) CORN_CUDA(ctx, cudaEventRecord(ctx->aux_ev[1], st_dense));
        CORN_CUDA(ctx, cudaEventRecord(ctx->ev[14], st));
    }
    {
        const sd_kernel_t kern = k_sdust_scan<2, true, false>;
        CORN_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int blocks_per_sm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, SD_BLOCK, smem) != cudaSuccess || blocks_per_sm < 1) { cudaGetLastError(); blocks_per_sm = 1; }
        sp.list_lo = d_tot + 11;
        sp.list_hi = d_tot + 1;
        sp.task_counter = d_tot + 8;
        if (overlap && blocks_per_sm > 6) blocks_per_sm = 6;
        const unsigned want = (n_items + SD_BLOCK - 1) / SD_BLOCK, resident = (unsigned)(ctx->sm_count * blocks_per_sm);
        kern<<<want < resident ? want : resident, SD_BLOCK, smem, st>>>(sp);
        corn_count_launch(ctx);
        CORN_LAUNCH_CHECK(ctx);
        if (overlap) CORN_CUDA(ctx, cudaStreamWaitEvent(st, ctx->aux_ev[1], 0));      // the fold needs both classes
    }
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));

    // ---- fold across item seams, compact ----
    ItemGatherParams gp;
    gp.slots = slots; gp.cnt = cnt; gp.it_c0 = it_c0; gp.it_flags = it_flags; gp.it_rec = it_rec; gp.n_items = d_tot + 1;
    gp.cap = cap; gp.W = W; gp.out_cnt = out_cnt; gp.out_off = out_off; gp.out = NULL;
    k_sdust_item_gather<<<(n_items + 255) / 256, 256, 0, st>>>(gp, it_off, 0);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    CORN_TRY(corn_scan_u32(ctx, out_cnt, out_off, n_items, d_tot));
    CORN_TRY(corn_read_small(ctx, hv, d_tot, 64));
    const uint32_t n_iv = hv[0];
    if (hv[4]) return corn_set_err(ctx, CORN_E_INTERNAL, "sdust: %u items overflowed their interval slot", hv[4]);
    // (the item tables live in sd_out: the compacted intervals and rec_first go to a buffer of their own)
    const size_t need_out = ((size_t)n_iv + 1) * sizeof(uint64_t) + ((size_t)n_rec + 1) * sizeof(uint64_t);
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->bitmap, need_out));
    gp.out = (uint64_t *)ctx->bitmap.p;
    uint64_t *d_first = gp.out + n_iv + 1;
    if (n_iv) {
        k_sdust_item_gather<<<(n_items + 255) / 256, 256, 0, st>>>(gp, it_off, 1);
        corn_count_launch(ctx);
        CORN_LAUNCH_CHECK(ctx);
    }
    k_sdust_item_rec_first<<<(n_rec + 1 + 255) / 256, 256, 0, st>>>(it_rec, out_off, d_tot + 1, d_tot, n_rec, d_first);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));

    out->n_iv = n_iv;
    if (n_iv) {
        out->iv = (uint64_t *)corn_host_alloc((size_t)n_iv * sizeof(uint64_t));
        if (!out->iv) return corn_set_err(ctx, CORN_E_NOMEM, "pinned alloc of %u intervals", n_iv);
        CORN_CUDA(ctx, cudaMemcpyAsync(out->iv, gp.out, (size_t)n_iv * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    }
    CORN_CUDA(ctx, cudaMemcpyAsync(h_first, d_first, ((size_t)n_rec + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[6], st));
    CORN_CUDA(ctx, cudaStreamSynchronize(st));
    float a = 0, b = 0;
    cudaEventElapsedTime(&ctx->timing.scan_ms, ctx->ev[3], ctx->ev[4]);      // scout + item table + item phase
    if (getenv("CORNETTO_TRACE")) {
        float t_scout = 0, t_items = 0, t_dense = 0, t_sparse = 0;
        cudaEventElapsedTime(&t_scout, ctx->ev[3], ctx->ev[12]);
        cudaEventElapsedTime(&t_items, ctx->ev[12], ctx->ev[13]);
        cudaEventElapsedTime(&t_dense, ctx->ev[13], ctx->ev[14]);
        cudaEventElapsedTime(&t_sparse, ctx->ev[14], ctx->ev[4]);
        fprintf(stderr, "[sdust] scout %.3f ms (%u chunks), item table %.3f ms (%u items of %u blocks), dense items %.3f ms, sparse items %.3f ms, %u intervals\n",
                t_scout, n_chunks, t_items, n_items, n_blk, t_dense, t_sparse, n_iv);
    }
    cudaEventElapsedTime(&a, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&b, ctx->ev[4], ctx->ev[5]);
    ctx->timing.post_ms = a + b;
    cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[5], ctx->ev[6]);
    ctx->timing.out_bytes = (uint64_t)n_iv * 8;
    return CORN_OK;
}

static int sdust_run(corn_ctx *ctx, const corn_dbatch *db, int T, int W, corn_intervals_t *out)
{
    // Window range.  The reference accepts whatever atoi() returns (src/sdust/sdust.c:186-189):
    //   W >= 3   works (W - 2 triplets per window);
    //   W <  3   its window can never hold a triplet and kdq_shift() of the empty deque is dereferenced (:69-70) --
    //            it crashes on the first triplet without printing anything.  Here: no interval, status OK.
    //   W > 1024 its own 32-bit score products (r * l, :113-117) leave the int range inside long low-complexity
    //            runs, so there is no defined result to reproduce: refused.
    if (W > CORN_SDUST_MAX_W) return corn_set_err(ctx, CORN_E_ARG, "sdust window %d above %d (the reference's 32-bit score products overflow there)", W, CORN_SDUST_MAX_W);
    const bool wide = W > SD_MAX_W;
    CORN_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->pending) { CORN_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); CORN_TRY(corn_telofind_resolve(ctx)); }
    cudaStream_t st = ctx->stream;
    const float keep_h2d = ctx->timing.h2d_ms;
    memset(&ctx->timing, 0, sizeof ctx->timing);
    ctx->timing.h2d_ms = keep_h2d;
    out->iv = NULL; out->rec_first = NULL; out->n_iv = 0; out->n_rec = db->n_rec; out->_owner = NULL;
    // the defaults (and anything with W <= 64, 20 <= T <= 24) take the two-phase path; $CORNETTO_SDUST_CLASSIC=1 forces
    // the chunk-grid kernels, which serve every other (T, W)
    if (sd_scout_supported(T, W) && !(getenv("CORNETTO_SDUST_CLASSIC") && atoi(getenv("CORNETTO_SDUST_CLASSIC")) > 0))
    {
        const int r = sdust_run_fast(ctx, db, T, W, out);
        if (r != CORN_E_STATE) return r;
    }

    // chunk length: 4096 bases for large batches (5 % warm-up overhead).  A chunk is a serial chain, and
    // the chain of a chunk inside a tandem repeat is ~4x slower than the rest, so on a small batch the
    // kernel lasts as long as that one chain: balance "all work / machine rate" against "C x slow-step
    // time", which on a B200 puts C near n_bases / 270 000.  Results do not depend on it.
    const size_t smem = SdLayout(wide ? SD_MAX_W : W).bytes();
    typedef void (*sd_kernel_t)(const SdParams);
    const sd_kernel_t kern = W <= 66 ? (T >= 5 ? k_sdust_scan<2, true> : k_sdust_scan<2, false>)
                                     : (T >= 5 ? k_sdust_scan<4, true> : k_sdust_scan<4, false>);
    int blocks_per_sm = 1;
    if (!wide) {
        CORN_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, SD_BLOCK, smem) != cudaSuccess || blocks_per_sm < 1) { cudaGetLastError(); blocks_per_sm = 1; }
    }
    // Chunk length.  Every chunk pays ~3W warm-up positions, which favours long chunks; but the kernel cannot end
    // before its most expensive warp-task does (one with a lane inside a tandem repeat runs 3-4x as long), which
    // favours many short tasks when the batch is small.  Measured optimum (scripts/chunk_sweep.sh, Gbases/s):
    // 200 Mb: 512 -> 42, 768 -> 32, 2048 -> 17;  1 Gb: 1024..1536 -> 60, 4096 -> 47;  3 Gb: 3072 -> 69, 512 -> 57.
    int C = 3072;
    {
        const uint64_t want = db->n_bases / 1000000u;
        if (want < 3072) C = (int)(want < 512 ? 512 : want / 64 * 64);
    }
    if (wide && C < 16 * W) C = 16 * W;          // the ~3W warm-up positions of a chunk stay below a fifth of it
    if (const char *e = getenv("CORNETTO_SDUST_CHUNK")) { int v = atoi(e); if (v >= 16 && v <= (1 << 20)) C = v; }
    const uint32_t n_rec = db->n_rec;
    const uint32_t cap = (uint32_t)(C + 2 * W) / 4 + 2;

    // tables: nch[n_rec+1] | chunk_base[n_rec+1] | (later) cnt[n_chunks] | out_cnt | out_off
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->misc, 4096));
    uint32_t *d_tot = (uint32_t *)((uint8_t *)ctx->misc.p + 2048);
    uint32_t *d_err = d_tot + 4;
    CORN_CUDA(ctx, cudaMemsetAsync(d_tot, 0, 64, st));
    // host-side chunk count (lengths are known on the host): sizes the tables without a readback
    uint64_t n_chunks64 = 0;
    for (uint32_t r = 0; r < n_rec; ++r) n_chunks64 += ((uint64_t)db->h_rec_len[r] + C - 1) / C;
    if (n_chunks64 > 0xFFFFFFF0ull) return corn_set_err(ctx, CORN_E_TOOBIG, "too many sdust chunks");
    const uint32_t n_chunks = (uint32_t)n_chunks64;
    const uint32_t n_tasks = (n_chunks + 31u) / 32u;
    const size_t tab_words = 2 * ((size_t)n_rec + 1) + 3 * ((size_t)n_chunks + 1) + (size_t)n_tasks + ((size_t)n_chunks + 3) / 4 + 32;
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->sd_tab, tab_words * sizeof(uint32_t)));
    uint32_t *nch = (uint32_t *)ctx->sd_tab.p, *chunk_base = nch + n_rec + 1;
    uint32_t *cnt = chunk_base + n_rec + 1, *out_cnt = cnt + n_chunks + 1, *out_off = out_cnt + n_chunks + 1;
    uint32_t *task_list = out_off + n_chunks + 1;
    uint8_t *heavy_flag = (uint8_t *)(task_list + n_tasks + 1);

    uint64_t *h_first = (uint64_t *)corn_host_alloc(sizeof(uint64_t) * ((size_t)n_rec + 1));
    if (!h_first) return corn_set_err(ctx, CORN_E_NOMEM, "pinned alloc");
    out->rec_first = h_first;
    if (n_chunks == 0 || W < 3) {
        for (uint32_t r = 0; r <= n_rec; ++r) h_first[r] = 0;
        return CORN_OK;
    }
    const size_t slot_words = (size_t)(W | 1);
    const size_t iv_bytes = ((size_t)n_chunks * cap * sizeof(uint64_t) + 255) & ~(size_t)255;
    const size_t state_bytes = wide ? (size_t)n_chunks * corn_sdust_wide_state_stride(W) : (size_t)n_chunks * slot_words * sizeof(uint32_t);
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->sd_slots, iv_bytes + state_bytes));

    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    k_sdust_nchunks<<<(n_rec + 255) / 256, 256, 0, st>>>(db->d_rec_len, nch, n_rec, (uint32_t)C);
    corn_count_launch(ctx);
    CORN_TRY(corn_scan_u32(ctx, nch, chunk_base, n_rec, chunk_base + n_rec));

    SdParams sp;
    memset(&sp, 0, sizeof sp);                        // (chunk mode: no item tables)
    sp.seq = db->d_seq; sp.rec_off = db->d_rec_off; sp.rec_len = db->d_rec_len; sp.chunk_base = chunk_base;
    sp.n_rec = n_rec; sp.n_chunks = n_chunks; sp.T = T; sp.W = W; sp.C = C; sp.cap = cap;
    sp.slots = (uint64_t *)ctx->sd_slots.p; sp.cnt = cnt; sp.err = d_err; sp.task_counter = d_tot + 8;
    sp.gslots = (uint32_t *)((uint8_t *)ctx->sd_slots.p + iv_bytes);
    sp.task_list = task_list;
    if (wide) {
        corn_sdust_wide_params wp;
        wp.seq = sp.seq; wp.rec_off = sp.rec_off; wp.rec_len = sp.rec_len; wp.chunk_base = chunk_base;
        wp.n_rec = n_rec; wp.n_chunks = n_chunks; wp.T = T; wp.W = W; wp.C = C; wp.cap = cap;
        wp.slots = sp.slots; wp.cnt = cnt; wp.err = d_err;
        wp.state = (uint8_t *)sp.gslots; wp.state_stride = corn_sdust_wide_state_stride(W);
        CORN_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
        CORN_TRY(corn_sdust_wide_scan(ctx, wp));
    } else {
        CORN_CUDA(ctx, cudaMemsetAsync(sp.gslots, 0, (size_t)n_chunks * slot_words * sizeof(uint32_t), st));
        k_sdust_probe<<<(n_chunks + 255) / 256, 256, 0, st>>>(sp, heavy_flag);
        k_sdust_order<<<(n_tasks + 255) / 256, 256, 0, st>>>(heavy_flag, n_chunks, n_tasks, task_list, d_tot + 9);
        corn_count_launch(ctx, 2);
        CORN_LAUNCH_CHECK(ctx);
        CORN_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
        const unsigned want = (n_chunks + SD_BLOCK - 1) / SD_BLOCK, resident = (unsigned)(ctx->sm_count * blocks_per_sm);
        kern<<<want < resident ? want : resident, SD_BLOCK, smem, st>>>(sp);
        corn_count_launch(ctx);
        CORN_LAUNCH_CHECK(ctx);
    }
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));

    GatherParams gp;
    gp.slots = sp.slots; gp.cnt = cnt; gp.chunk_base = chunk_base; gp.rec_len = db->d_rec_len;
    gp.n_rec = n_rec; gp.n_chunks = n_chunks; gp.cap = cap; gp.C = C; gp.W = W;
    gp.out_cnt = out_cnt; gp.out_off = out_off; gp.out = NULL;
    k_sdust_gather<<<(n_chunks + 255) / 256, 256, 0, st>>>(gp, 0);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    CORN_TRY(corn_scan_u32(ctx, out_cnt, out_off, n_chunks, d_tot));
    uint32_t hv[8];
    CORN_TRY(corn_read_small(ctx, hv, d_tot, 32));
    const uint32_t n_iv = hv[0];
    if (hv[4]) return corn_set_err(ctx, CORN_E_INTERNAL, "sdust: %u chunks overflowed their interval slot", hv[4]);
    uint64_t *d_first = NULL;
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->sd_out, ((size_t)n_iv + 1) * sizeof(uint64_t) + ((size_t)n_rec + 1) * sizeof(uint64_t)));
    gp.out = (uint64_t *)ctx->sd_out.p;
    d_first = gp.out + n_iv + 1;
    if (n_iv) {
        k_sdust_gather<<<(n_chunks + 255) / 256, 256, 0, st>>>(gp, 1);
        corn_count_launch(ctx);
        CORN_LAUNCH_CHECK(ctx);
    }
    k_sdust_rec_first<<<(n_rec + 1 + 255) / 256, 256, 0, st>>>(chunk_base, out_off, d_tot, n_rec, n_chunks, d_first);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));

    out->n_iv = n_iv;
    if (n_iv) {
        out->iv = (uint64_t *)corn_host_alloc((size_t)n_iv * sizeof(uint64_t));
        if (!out->iv) return corn_set_err(ctx, CORN_E_NOMEM, "pinned alloc of %u intervals", n_iv);
        CORN_CUDA(ctx, cudaMemcpyAsync(out->iv, gp.out, (size_t)n_iv * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    }
    CORN_CUDA(ctx, cudaMemcpyAsync(h_first, d_first, ((size_t)n_rec + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CORN_CUDA(ctx, cudaEventRecord(ctx->ev[6], st));
    CORN_CUDA(ctx, cudaStreamSynchronize(st));
    float a = 0, b = 0;
    cudaEventElapsedTime(&ctx->timing.scan_ms, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&a, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&b, ctx->ev[4], ctx->ev[5]);
    ctx->timing.post_ms = a + b;
    cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[5], ctx->ev[6]);
    ctx->timing.out_bytes = (uint64_t)n_iv * 8;
    return CORN_OK;
}

extern "C" int corn_gpu_sdust_dev(corn_ctx_t *ctx, const corn_dbatch_t *db, int T, int W, corn_intervals_t *out)
{
    if (!ctx || !db || !out) return CORN_E_ARG;
    int r = sdust_run(ctx, db, T, W, out);
    if (r != CORN_OK) corn_gpu_intervals_free(out);
    return r;
}

extern "C" int corn_gpu_sdust(corn_ctx_t *ctx, const corn_batch_t *batch, int T, int W, corn_intervals_t *out)
{
    if (!ctx || !batch || !out) return CORN_E_ARG;
    corn_dbatch_t *db = NULL;
    corn_ctx_adopt(ctx, NULL);   // retire the previous resident batch first: its buffer is reused by the upload
    CORN_TRY(corn_gpu_upload(ctx, batch, &db));
    const float h2d = ctx->timing.h2d_ms;
    corn_ctx_adopt(ctx, db);
    ctx->last_db = NULL;            // the telofind state of this context no longer matches the resident batch
    ctx->timing.h2d_ms = h2d;
    return corn_gpu_sdust_dev(ctx, db, T, W, out);
}

extern "C" void corn_gpu_intervals_free(corn_intervals_t *iv)
{
    if (!iv) return;
    corn_host_free(iv->iv);
    corn_host_free(iv->rec_first);
    iv->iv = NULL; iv->rec_first = NULL; iv->n_iv = 0; iv->_owner = NULL;
}
