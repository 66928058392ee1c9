// cornetto_b200/csrc/sdust_core.cuh -- symmetric DUST state machine for one chunk of one record.
//
// Restates sdust_core() of src/sdust/sdust.c:130-160 (lh3/sdust 0.1-r2) in a form that fits a GPU
// thread: bounded state, no heap, and a provably exact way to start in the middle of a record.
// The same code compiles for the device (sdust.cu: one thread per chunk, state in shared memory)
// and for the host (tests/sim/: a CPU simulation of the chunked execution, checked against the
// oracle without a GPU).
//
// 1. Bounded perfect-interval list.  The reference keeps a list P of perfect intervals sorted
//    by descending start (sdust.c:104-128) that can hold ~W^2/2 entries inside a long
//    homopolymer.  Only two things are ever read from it:
//      - the best score ratio among entries with start >= s (the j-loop, :113-117), and
//      - for each start s, the entry inserted last (largest finish), which is the one saved
//        when s leaves the window (:88-102); the others with the same start are contained in it.
//    A later insertion with the same start always has a ratio >= the earlier ones (it had to
//    beat them to be inserted) and a larger finish, so ONE slot per start position -- a ring of
//    W slots holding (r, l, finish-start) -- reproduces every decision and every saved interval.
//
// 2. Starting mid-record.  The window state (w, cw, rw, L, cv, rv) is a pure function of the last
//    W-2 emitted triplets, the slots only matter while their start is inside the window, and at
//    any non-ACGT byte the reference flushes all of P (:152-156) but keeps the window (the "stale
//    window" quirk).  Let a run be started fresh (empty window, no slots) at p0 < c0 and let tau0
//    be the position at which it has emitted W-2 triplets.  From tau0 on its window equals the
//    true one; any slot that differs from the true execution has start <= ws(tau0) + W-3 and
//    finish <= start + W, and a differing decision can only create a slot whose start is <= that
//    of a differing slot it contains (and sd_save may save/drop the slot one start to its right
//    differently).  Hence every interval touching positions >= ws(tau0) + 2W - 2 is identical, and since ws(tau0) <= tau0 it suffices that
//        c0 >= tau0 + 2W.
//    sdust_warm_start() walks back from c0 - 2W until W triplet positions have been seen (W-2
//    plus two for the run-length difference at p0) or the record start is reached (exact).
//    For N-free sequence this is a 3W+2 = 194 byte warm-up; with Ns it stretches as needed.
//    A chunk owns the save events that happen while positions [c0, c1) are processed; it stops at
//    c1 (the next chunk re-creates the state there), so there is no right halo at all.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SD_HD __host__ __device__ __forceinline__
#else
#define SD_HD static inline
#endif

// Two instances of this header exist.  The default ("narrow") one is the hot path: W <= 128, byte counters,
// 32-bit slots, state in shared memory.  With SD_WIDE defined before inclusion the same code is compiled with
// 16-bit counters and 64-bit slots for 128 < W <= SD_MAX_W = 1024 (csrc/sdust_wide.cu; `sdust -w 200`): the
// reference takes any W (src/sdust/sdust.c:186-189), and up to there its own 32-bit arithmetic (score products
// r * l in find_perfect, :113-117) is well defined.  Everything lives in a namespace per instance, so the two
// can be linked into one library.
#if defined(SD_WIDE)
#define SD_MAX_W 1024
#define SD_NS sd_wide
typedef uint16_t sd_cnt_t;
typedef uint64_t sd_slot_t;
#else
#define SD_MAX_W 128
#define SD_NS sd_narrow
typedef uint8_t  sd_cnt_t;
typedef uint32_t sd_slot_t;
#endif
namespace SD_NS {

// seq_nt4_table, src/sdust/sdust.c:23-40: A/a C/c G/g T/t -> 0..3, bytes 0..3 -> 0..3, rest 4
// Branch-free on purpose: written as an if/else chain the four bases became four divergent paths that
// only rejoined at the end of the step, i.e. the whole per-base body ran four times per warp.
SD_HD int sd_nt4(uint8_t c)
{
    const uint32_t u = c & 0xDFu;                          // fold case
    const uint32_t k = (c >> 1) & 3u;                      // A 0, C 1, T 2, G 3 for valid letters
    const uint32_t expect = (0x47544341u >> (8u * k)) & 0xFFu;   // 'A','C','T','G'
    const uint32_t code = (0x2310u >> (4u * k)) & 3u;      // -> A 0, C 1, G 2, T 3
    const uint32_t r = (u == expect) ? code : 4u;
    return (int)(c < 4 ? (uint32_t)c : r);
}

// Per-thread state lives in shared memory on the device.
//   cw, cv   are indexed by data (the triplet): one 4-byte wide COLUMN per thread, so a thread always
//            hits its own bank whatever index it uses.  pitch = bytes between consecutive 4-byte rows
//            of a column; on the host pitch = 4 (a plain array).
//   ring, slot  are indexed by window position, which is (nearly) the same for all lanes of a warp:
//            plain per-thread arrays with an odd word stride between threads, conflict-free both
//            for that access and for the warp-cooperative routines, where 32 lanes read 32
//            consecutive entries of ONE thread's arrays.
struct sd_mem {
    uint8_t   *ring;   // W entries: triplet codes of the window deque
    sd_cnt_t  *cw;     // 64 window counters (narrow: a column, see above)
    sd_cnt_t  *cv;     // 64 suffix counters
    sd_slot_t *slot;   // W slots: narrow valid:1 | flen:8 | l:8 | r:15, wide valid:1 | flen:11 | l:11 | r:32
    uint32_t   pitch;  // bytes, for cw / cv (narrow only)
};

#if defined(SD_WIDE)
#define SD_U8(base, i)  ((base)[(i)])                         /* plain per-thread arrays */
#else
#define SD_U8(base, i)  (*((base) + ((uint32_t)(i) >> 2) * m.pitch + ((uint32_t)(i) & 3u)))
#endif
#define SD_RING(i)      (m.ring[(i)])
#define SD_SLOT(i)      (m.slot[(i)])

#if defined(SD_WIDE)
#define SD_SLOT_VALID 0x8000000000000000ull
SD_HD sd_slot_t sd_slot_pack(int r, int l, int flen) { return SD_SLOT_VALID | ((uint64_t)flen << 43) | ((uint64_t)l << 32) | (uint32_t)r; }
SD_HD int sd_slot_r(sd_slot_t s) { return (int)(uint32_t)s; }
SD_HD int sd_slot_l(sd_slot_t s) { return (int)((s >> 32) & 0x7FFu); }
SD_HD int sd_slot_flen(sd_slot_t s) { return (int)((s >> 43) & 0x7FFu); }
#else
#define SD_SLOT_VALID 0x80000000u
SD_HD sd_slot_t sd_slot_pack(int r, int l, int flen) { return SD_SLOT_VALID | ((uint32_t)flen << 23) | ((uint32_t)l << 15) | (uint32_t)r; }
SD_HD int sd_slot_r(sd_slot_t s) { return (int)(s & 0x7FFFu); }
SD_HD int sd_slot_l(sd_slot_t s) { return (int)((s >> 15) & 0xFFu); }
SD_HD int sd_slot_flen(sd_slot_t s) { return (int)((s >> 23) & 0xFFu); }
#endif

// interval sink = the reference's `res` vector for the save events this chunk owns.
// save_masked_regions() appends an interval unless it starts at or before the finish of the
// LAST saved one, in which case only that finish may grow (sdust.c:94-98).  After an N the
// stale window can produce saves that are not in ascending order, and then this rule is not a
// plain union (an earlier start is swallowed) -- so chunks own save EVENTS BY TIME (the input
// position being processed when the save happens), fold them with this very rule, and the
// per-chunk lists are folded again across seams (sd_gather_*) with the same rule.
struct sd_sink {
    uint64_t *out;
    uint32_t  n, cap;
    int       on;              // events before the chunk's first owned position are warm-up: ignored
    int       cur_s, cur_f;    // last saved interval (cur_f < 0: none); written out when superseded
    uint32_t  overflow;
};

SD_HD void sd_sink_init(sd_sink &k, uint64_t *out, uint32_t cap)
{
    k.out = out; k.n = 0; k.cap = cap; k.on = 0; k.cur_s = 0; k.cur_f = -1; k.overflow = 0;
}
SD_HD void sd_sink_close(sd_sink &k)
{
    if (k.cur_f >= 0) {
        if (k.n < k.cap) k.out[k.n] = (uint64_t)(uint32_t)k.cur_s << 32 | (uint32_t)k.cur_f;
        else k.overflow = 1;
        ++k.n;
        k.cur_f = -1;
    }
}
SD_HD void sd_sink_put(sd_sink &k, int s, int f)
{
    if (!k.on) return;
    if (k.cur_f >= 0 && s <= k.cur_f) { if (f > k.cur_f) k.cur_f = f; return; }
    sd_sink_close(k);
    k.cur_s = s; k.cur_f = f;
}

struct sd_state {
    int wn, whead;       // deque size / index of its oldest element in the ring
    int L, rw, rv;
    int nslot;           // valid slots
    int pstart;          // every valid slot has start >= pstart
    int pslot;           // pstart mod W, kept incrementally (no integer division in the per-base path)
    int l;               // length of the current A/C/G/T run
    unsigned t;          // current triplet
    int slack;           // >= 0: a lower bound of T*new_l - 10*new_r over every suffix find_perfect would examine,
                         //       i.e. it would find no candidate and may be skipped; < 0: unknown (see sd_slack_*)
};

SD_HD void sd_reset_counters(sd_state &s, const sd_mem &m)
{
    s.wn = s.whead = s.L = s.rw = s.rv = s.nslot = s.pstart = s.pslot = s.l = 0;
    s.t = 0;
    s.slack = -1;
    for (int i = 0; i < 64; ++i) { SD_U8(m.cw, i) = 0; SD_U8(m.cv, i) = 0; }
}

SD_HD void sd_reset(sd_state &s, const sd_mem &m, int W)
{
    sd_reset_counters(s, m);
    for (int i = 0; i < W; ++i) SD_SLOT(i) = 0;
}

SD_HD int sd_ring_idx(const sd_state &s, int i, int W)
{
    int k = s.whead + i;
    return k >= W ? k - W : k;
}

// shift_window(): sdust.c:66-86, in two halves so that the device can run the second one
// (a data-dependent loop) cooperatively.  _push does everything up to and including the counter
// updates for the new triplet and reports whether the suffix v must shrink; _pop is that loop.
// ---- skipping find_perfect calls that cannot find anything ------------------------------------------
// find_perfect (:104-128) examines the window suffixes longer than L and inserts one only if
// new_r*10 > T*new_l (:112); a call without such a candidate changes nothing.  Let
//   f(a) = T*new_l(a) - 10*new_r(a)      for the suffix starting at window element a.
// Pushing triplet t makes every suffix one element longer (new_l + 1) and adds to new_r the number of
// t already in it, at most cw[t] before the push:  f'(a) >= f(a) + T - 10*cw[t].  The set of examined
// starts only loses elements (the head that leaves the window) as long as L grows with the pushes; it
// gains some only when the shrink loop shortens L.  So after a call that found no candidate,
// slack = min f >= 0 is kept per chunk, updated by T - 10*cw[t] at every push, and reset to "unknown"
// by a shrink or by a call that did find candidates; while it stays >= 0 the reference's own trigger
// (rw*10 > L*T, :149) may fire but the call is skipped.  (Most triggers come in bursts after a shrink:
// the first one evaluates, the following ones are usually covered.)
#if defined(SD_SLACK_STATS)
static unsigned long long sd_stat_eval, sd_stat_skip;     // host-only counters for tests/sim (how often the skip applies)
#endif
SD_HD void sd_slack_push(sd_state &s, int T, int cw_before) { if (s.slack >= 0) s.slack += T - 10 * cw_before; }
SD_HD void sd_slack_unknown(sd_state &s) { s.slack = -1; }

// cv_max = floor(2T / 10): the largest count a triplet may have inside the suffix v (:79)
SD_HD bool sd_shift_window_push(sd_state &s, const sd_mem &m, int t, int T, int cv_max, int W)
{
    if (s.wn >= W - 2) {
        const int x = SD_RING(s.whead);
        s.whead = s.whead + 1 >= W ? 0 : s.whead + 1;
        --s.wn;
        const int c = SD_U8(m.cw, x) - 1;
        SD_U8(m.cw, x) = (sd_cnt_t)c;
        s.rw -= c;
        if (s.L > s.wn) {
            --s.L;
            const int d = SD_U8(m.cv, x) - 1;
            SD_U8(m.cv, x) = (sd_cnt_t)d;
            s.rv -= d;
        }
    }
    SD_RING(sd_ring_idx(s, s.wn, W)) = (uint8_t)t;
    ++s.wn;
    ++s.L;
    const int c = SD_U8(m.cw, t);
    s.rw += c;
    sd_slack_push(s, T, c);
    SD_U8(m.cw, t) = (sd_cnt_t)(c + 1);
    const int d = SD_U8(m.cv, t);
    s.rv += d;
    SD_U8(m.cv, t) = (sd_cnt_t)(d + 1);
    return d + 1 > cv_max;                           // (d + 1) * 10 > 2 * T
}

SD_HD void sd_shift_window_pop(sd_state &s, const sd_mem &m, int t, int W)
{
    int x;
    do {
        x = SD_RING(sd_ring_idx(s, s.wn - s.L, W));
        const int e = SD_U8(m.cv, x) - 1;
        SD_U8(m.cv, x) = (sd_cnt_t)e;
        s.rv -= e;
        --s.L;
    } while (x != t);
}

SD_HD void sd_shift_window(sd_state &s, const sd_mem &m, int t, int T, int W)
{
    if (sd_shift_window_push(s, m, t, T, (T << 1) / 10, W)) { sd_shift_window_pop(s, m, t, W); sd_slack_unknown(s); }
}

// save_masked_regions(): sdust.c:88-102.  If the smallest start is below `start`, that ONE slot
// is saved and every slot below `start` is dropped -- including, when `start` advanced by more
// than one (it does by two at a flush with l >= W), slots that were never saved.  That loss is
// reference behaviour and is reproduced here.
// (pstart, pslot) are only meaningful while some slot is valid: with none, they are left stale
// (always pslot == pstart mod W) and re-anchored by the insertion that creates the first slot
// (sd_anchor_pstart) -- the common per-base path then has nothing to maintain.
SD_HD void sd_anchor_pstart(sd_state &s, int start, int base)
{
    s.pstart = start; s.pslot = base;            // base == start mod W, computed by the caller
}

SD_HD void sd_save(sd_state &s, const sd_mem &m, sd_sink &k, int start, int W)
{
    if (s.nslot == 0) return;
    while (s.pstart < start && !(SD_SLOT(s.pslot) & SD_SLOT_VALID)) { ++s.pstart; if (++s.pslot == W) s.pslot = 0; }
    if (s.pstart >= start) return;                       // smallest start >= start: nothing to do
    sd_sink_put(k, s.pstart, s.pstart + sd_slot_flen(SD_SLOT(s.pslot)));
    while (s.pstart < start) {
        if (SD_SLOT(s.pslot) & SD_SLOT_VALID) {
            SD_SLOT(s.pslot) = 0;
            if (--s.nslot == 0) return;
        }
        ++s.pstart; if (++s.pslot == W) s.pslot = 0;
    }
}

// flush at a non-ACGT byte or at the end of the sequence: sdust.c:152-154
//   start = max(l-W+1,0) + (i+1-l);  while (P.n) save_masked_regions(start++)
SD_HD void sd_flush(sd_state &s, const sd_mem &m, sd_sink &k, int start, int W)
{
    while (s.nslot) sd_save(s, m, k, start++, W);
}

// find_perfect(): sdust.c:104-128 with the slot ring in place of P
SD_HD void sd_find_perfect(sd_state &s, const sd_mem &m, int T, int start, int W)
{
    int r = s.rv, max_r = 0, max_l = 0;
    const int i0 = s.wn - s.L - 1;
    // slot of window index i is (start + i) mod W; every valid slot has start >= pstart >= ... so the
    // base is derived from the tracked pstart/pslot pair instead of a division
    int base = s.pslot + (start - s.pstart);
    if (base >= W || base < 0) base = (int)((uint32_t)start % (uint32_t)W);
    // entries whose start lies right of the first candidate also count for the maximum (the
    // reference's j-loop always starts from the largest start, :113)
    if (s.nslot)
        for (int i = s.wn - 1; i > i0; --i) {
            int si = base + i; if (si >= W) si -= W;
            const sd_slot_t v = SD_SLOT(si);
            if (v & SD_SLOT_VALID) {
                const int pr = sd_slot_r(v), pl = sd_slot_l(v);
                if (max_r == 0 || pr * max_l > max_r * pl) { max_r = pr; max_l = pl; }
            }
        }
    int ri = sd_ring_idx(s, i0 < 0 ? 0 : i0, W);          // ring position of window index i, walked downwards
    for (int i = i0; i >= 0; --i) {
        const int t = SD_RING(ri);
        if (--ri < 0) ri = W - 1;
        const int c = SD_U8(m.cv, t);
        r += c;
        SD_U8(m.cv, t) = (sd_cnt_t)(c + 1);          // temporary; undone below (the reference copies cv)
        const int new_r = r, new_l = s.wn - i - 1;
        int si = base + i; if (si >= W) si -= W;
        const sd_slot_t v = SD_SLOT(si);
        if (v & SD_SLOT_VALID) {                      // entries with this start join the running maximum
            const int pr = sd_slot_r(v), pl = sd_slot_l(v);
            if (max_r == 0 || pr * max_l > max_r * pl) { max_r = pr; max_l = pl; }
        }
        if (new_r * 10 > T * new_l) {
            if (max_r == 0 || new_r * max_l >= max_r * new_l) {
                max_r = new_r; max_l = new_l;
                if (!(v & SD_SLOT_VALID)) { if (s.nslot == 0) sd_anchor_pstart(s, start, base); ++s.nslot; }
                SD_SLOT(si) = sd_slot_pack(new_r, new_l, s.wn + 2 - i);
            }
        }
    }
    ri = sd_ring_idx(s, i0 < 0 ? 0 : i0, W);
    for (int i = i0; i >= 0; --i) {
        const int t = SD_RING(ri);
        if (--ri < 0) ri = W - 1;
        SD_U8(m.cv, t) = (sd_cnt_t)(SD_U8(m.cv, t) - 1);
    }
}

// ---- data-parallel form of find_perfect ------------------------------------------------------
// The serial loop above hides three independent scans.  With t_i the triplet at window index i,
// i0 = wn-L-1 and rank(i) = #{k <= i : t_k == t_i}:
//   c_i      = cw[t_i] - rank(i)                  (what "r += c[t]++" adds at index i)
//   new_r(i) = rv + sum_{k=i..i0} c_k             (suffix sum)
//   M_i      = best ratio among the slots of indices >= i and the candidates of indices > i
//              (a rejected candidate is below the running maximum, so it may be included)
//   insert i <=> new_r*10 > T*new_l  and  new_r/new_l >= M_i
// This is what the warp-cooperative device version (sdust.cu) evaluates with match/scan
// primitives; sd_find_perfect_vec is the same arithmetic with plain loops so that the CPU
// simulation can check it against the serial form (SD_USE_VEC).  Requires T >= 5 (so that every
// candidate has new_l >= 1), which the callers check.
SD_HD void sd_fracmax(int &ar, int &al, int br, int bl) { if (br * al > ar * bl) { ar = br; al = bl; } }

SD_HD void sd_find_perfect_vec(sd_state &s, const sd_mem &m, int T, int start, int W)
{
    const int wn = s.wn, i0 = wn - s.L - 1;
    int base = s.pslot + (start - s.pstart);
    if (base >= W || base < 0) base = (int)((uint32_t)start % (uint32_t)W);
    int tt[SD_MAX_W], c[SD_MAX_W], nr[SD_MAX_W], er[SD_MAX_W], el[SD_MAX_W], pr[SD_MAX_W], pl[SD_MAX_W], sv[SD_MAX_W];
    int cnt[64];
    for (int t = 0; t < 64; ++t) cnt[t] = 0;
    for (int i = 0; i < wn; ++i) {                        // ranks (ascending)
        tt[i] = SD_RING(sd_ring_idx(s, i, W));
        const int rank = ++cnt[tt[i]];
        c[i] = i <= i0 ? (int)SD_U8(m.cw, tt[i]) - rank : 0;
    }
    int acc = 0;
    for (int i = wn - 1; i >= 0; --i) { acc += c[i]; nr[i] = s.rv + acc; }      // suffix sums
    {
        int fmin = 0x7fffffff;
        for (int i = 0; i <= i0; ++i) { const int f = T * (wn - i - 1) - 10 * nr[i]; if (f < fmin) fmin = f; }
        if (fmin >= 0) { s.slack = fmin < (1 << 20) ? fmin : (1 << 20); return; }   // no candidate: nothing to insert
        sd_slack_unknown(s);
    }
    for (int i = 0; i < wn; ++i) {                        // elements: slot and candidate
        int si = base + i; if (si >= W) si -= W;
        const sd_slot_t v = SD_SLOT(si);
        sv[i] = (v & SD_SLOT_VALID) != 0;
        pr[i] = sv[i] ? sd_slot_r(v) : 0; pl[i] = sv[i] ? sd_slot_l(v) : 1;
        er[i] = pr[i]; el[i] = pl[i];
        const int new_l = wn - i - 1;
        if (i <= i0 && nr[i] * 10 > T * new_l) sd_fracmax(er[i], el[i], nr[i], new_l);
    }
    int mr = 0, ml = 1;                                    // exclusive running maximum, descending i
    for (int i = wn - 1; i >= 0; --i) {
        const int new_l = wn - i - 1;
        if (i <= i0 && nr[i] * 10 > T * new_l) {
            int Mr = mr, Ml = ml;
            sd_fracmax(Mr, Ml, pr[i], pl[i]);
            if (nr[i] * Ml >= Mr * new_l) {
                int si = base + i; if (si >= W) si -= W;
                if (!sv[i]) { if (s.nslot == 0) sd_anchor_pstart(s, start, base); ++s.nslot; }
                SD_SLOT(si) = sd_slot_pack(nr[i], new_l, wn + 2 - i);
            }
        }
        sd_fracmax(mr, ml, er[i], el[i]);
    }
}

// one input position (i == l_seq is the virtual end-of-sequence byte, b = 4).
// Returns the window start of an emitting step, INT32_MAX after a flush (every slot resolved),
// or -1 for an A/C/G/T byte that does not complete a triplet.
SD_HD int sd_step(sd_state &s, const sd_mem &m, sd_sink &k, int i, int b, int T, int W)
{
    if (b < 4) {
        ++s.l;
        s.t = (s.t << 2 | (unsigned)b) & 63u;
        if (s.l < 3) return -1;
        const int start = (s.l - W > 0 ? s.l - W : 0) + (i + 1 - s.l);
        sd_save(s, m, k, start, W);
        sd_shift_window(s, m, (int)s.t, T, W);
#if defined(SD_USE_VEC)
        if (s.rw * 10 > s.L * T) {
#if defined(SD_SLACK_STATS)
            if (T >= 5 && s.wn - s.L - 1 >= 0) { if (s.slack < 0) ++sd_stat_eval; else ++sd_stat_skip; }
#endif
            if (T >= 5) { if (s.slack < 0 && s.wn - s.L - 1 >= 0) sd_find_perfect_vec(s, m, T, start, W); }
            else sd_find_perfect(s, m, T, start, W);
        }
#else
        if (s.rw * 10 > s.L * T) sd_find_perfect(s, m, T, start, W);
#endif
        return start;
    }
    sd_flush(s, m, k, (s.l - W + 1 > 0 ? s.l - W + 1 : 0) + (i + 1 - s.l), W);
    s.l = 0; s.t = 0;
    return 0x7fffffff;
}

// Where to start so that everything from c0 on is exact (see the header comment).
// fetch(i) returns byte i of the record.
template <class Fetch>
SD_HD int sd_warm_start(Fetch &fetch, int c0, int W)
{
    int p = c0 - 2 * W;
    if (p <= 0) return 0;
    int need = W, run = 0;
    // walk left until W triplet-completing positions have been passed: a maximal ACGT run of
    // length n holds n-2 of them (two more than the W-2 needed, for the run cut at p0)
    while (p > 0 && need > 0) {
        --p;
        if (sd_nt4(fetch(p)) < 4) { if (++run >= 3) --need; }
        else run = 0;
    }
    return p;
}

// Runs the state machine from the warm start and records the save events that happen while
// processing positions [c0, c1) of a record of length l_seq (c1 >= l_seq: last chunk, which also
// owns the reference's final i == l_seq flush).
template <class Fetch>
SD_HD void sd_run_chunk(Fetch &fetch, int l_seq, int c0, int c1, int T, int W, const sd_mem &m, sd_sink &k)
{
    sd_state s;
    sd_reset(s, m, W);
    const int p0 = sd_warm_start(fetch, c0, W);
    s.pstart = p0; s.pslot = (int)((uint32_t)p0 % (uint32_t)W);
    const int stop = c1 < l_seq ? c1 : l_seq;
    for (int i = p0; i < stop; ++i) {
        if (i == c0) k.on = 1;
        sd_step(s, m, k, i, sd_nt4(fetch(i)), T, W);
    }
    if (c1 >= l_seq) { k.on = 1; sd_step(s, m, k, l_seq, 4, T, W); }
    sd_sink_close(k);
}

// ---- items of the two-phase execution (see below) ---------------------------------------------------
// warm start of an item whose preceding block is quiet: P is known to be empty at c0, so only the window has to be
// rebuilt -- W triplet positions back from c0 itself (W-2 plus two for the run cut at p0), window half only
template <class Fetch>
SD_HD int sd_warm_quiet(Fetch &fetch, int c0, int W)
{
    int p = c0;
    if (p <= 0) return 0;
    int need = W, run = 0;
    while (p > 0 && need > 0) {
        --p;
        if (sd_nt4(fetch(p)) < 4) { if (++run >= 3) --need; }
        else run = 0;
    }
    return p;
}

// the window half of sd_step alone (no save, no find_perfect, no flush: P is empty and stays empty)
SD_HD void sd_step_window(sd_state &s, const sd_mem &m, int b, int T, int W)
{
    if (b < 4) {
        ++s.l;
        s.t = (s.t << 2 | (unsigned)b) & 63u;
        if (s.l >= 3) sd_shift_window(s, m, (int)s.t, T, W);
    } else { s.l = 0; s.t = 0; }
}

#define SD_ITEM_QUIET 1u     /* the block before c0 is quiet: window-only warm-up from sd_warm_quiet() */
#define SD_ITEM_CHAIN 2u     /* the previous item of the record ends exactly at c0 (same run of blocks): fold across the seam */

template <class Fetch>
SD_HD void sd_run_item(Fetch &fetch, int l_seq, int c0, int c1, uint32_t flags, int T, int W, const sd_mem &m, sd_sink &k)
{
    sd_state s;
    sd_reset(s, m, W);
    const bool quiet = (flags & SD_ITEM_QUIET) != 0;
    const int p0 = quiet ? sd_warm_quiet(fetch, c0, W) : sd_warm_start(fetch, c0, W);
    s.pstart = p0; s.pslot = (int)((uint32_t)p0 % (uint32_t)W);
    const int stop = c1 < l_seq ? c1 : l_seq;
    for (int i = p0; i < stop; ++i) {
        if (i == c0) k.on = 1;
        if (quiet && i < c0) sd_step_window(s, m, sd_nt4(fetch(i)), T, W);
        else sd_step(s, m, k, i, sd_nt4(fetch(i)), T, W);
    }
    if (c1 >= l_seq) { k.on = 1; sd_step(s, m, k, l_seq, 4, T, W); }
    sd_sink_close(k);
}

// ---- the machine in AGE ORDER, for items without a non-ACGT byte ------------------------------------------
// Inside low-complexity sequence find_perfect runs at every step and costs O(W) dependent iterations: served per lane
// it is latency bound, served cooperatively it occupies a whole warp for one lane.  For such DENSE items one warp
// runs ONE item, with the window spread over its lanes.  This is the plain-array statement of what that kernel
// (csrc/sdust.cu: sdust_dense_item) computes with shuffles; tests/sim runs it against the oracle.
//
// On a stretch [p0, c1) without non-ACGT bytes the fresh run is regular: every position from p0 + 2 on emits a triplet,
// the element pushed at position i has the interval start i - 2 for as long as it is in the window, and the window
// start advances by exactly one per step once it is full.  So the perfect-interval slot of a start position can live
// WITH the window element of that start, and everything is indexed by AGE j (0 = newest element):
//   x[j]     triplet;   slot[j]  the perfect interval that starts at this element (sd_slot_pack, 0 = none);
//   R[j]     score of the suffix that starts at this element = pairs of equal triplets among ages j..0.  Pushing t
//            adds to R[j] the number of t among ages j..1 -- so the scores find_perfect recomputes with its c[t]++ loop
//            (:108-112) are simply there, and the suffix counts cv[] / rv are never needed;
//   L        as in the scout: min(L + 1, wn, age of the 4th previous occurrence of t when t is then >= 5 times in the window).
// find_perfect in age order (cf. sd_find_perfect_vec): candidate ages j >= L with R[j]*10 > T*j; with
//   X[j] = best ratio among {slots and candidates of ages < j}  and  M[j] = max(X[j], slot[j]),
// age j is (re)inserted iff it is a candidate and R[j] / j >= M[j]; the interval is (r = R[j], l = j, flen = j + 3).
// save_masked_regions(): the only slot that can fall left of the window start is the one of the element that leaves.
struct sd_dense {
    int wn, L, rw;
    uint8_t  x[64];
    uint16_t R[64];
    uint32_t slot[64];
};

SD_HD void sd_dense_find_perfect(sd_dense &d, int T)
{
    int xr = 0, xl = 1;                               // X[j]: running maximum over ages < j
    for (int j = 0; j < d.wn; ++j) {
        const uint32_t v = d.slot[j];
        const bool sv = (v & SD_SLOT_VALID) != 0;
        const int pr = sv ? sd_slot_r(v) : 0, pl = sv ? sd_slot_l(v) : 1;
        const bool cand = j >= d.L && (int)d.R[j] * 10 > T * j;
        int er = pr, el = pl;                         // this age's contribution to the running maximum
        if (cand) {
            int mr = xr, ml = xl;
            sd_fracmax(mr, ml, pr, pl);               // M[j]
            if ((int)d.R[j] * ml >= mr * j) d.slot[j] = sd_slot_pack(d.R[j], j, j + 3);
            sd_fracmax(er, el, d.R[j], j);
        }
        sd_fracmax(xr, xl, er, el);
    }
}

// one emitting step: `start` = window start of this step, t = the new triplet
SD_HD void sd_dense_step(sd_dense &d, sd_sink &k, int start, int t, int T, int W, bool fp_enabled)
{
    if (d.wn >= W - 2) {                              // the oldest element leaves; its interval, if any, is saved (:88-102)
        const int o = d.wn - 1;
        if (d.slot[o] & SD_SLOT_VALID) sd_sink_put(k, start - 1, start - 1 + sd_slot_flen(d.slot[o]));
        int c = 0;
        for (int j = 0; j < o; ++j) c += d.x[j] == d.x[o];
        d.rw -= c;
        --d.wn;
    }
    if (d.L > d.wn) d.L = d.wn;
    for (int j = d.wn; j > 0; --j) { d.x[j] = d.x[j - 1]; d.R[j] = d.R[j - 1]; d.slot[j] = d.slot[j - 1]; }
    int cnt = 0, d4 = 64;
    for (int j = 1; j <= d.wn; ++j)
        if (d.x[j] == t) { ++cnt; if (cnt == 4) d4 = j; d.R[j] = (uint16_t)(d.R[j] + cnt); }
        else d.R[j] = (uint16_t)(d.R[j] + cnt);
    d.x[0] = (uint8_t)t; d.R[0] = 0; d.slot[0] = 0;
    d.rw += cnt;
    ++d.wn;
    ++d.L;
    if (cnt >= 4 && d4 < d.L) d.L = d4;
    if (fp_enabled && d.rw * 10 > d.L * T && d.L < d.wn) sd_dense_find_perfect(d, T);
}

// flush at the end of the record (:152-154): start counts up from start0; every round saves the valid slot with the
// smallest start if that is below `start`, and drops every slot below `start`.  a0 = interval start of the newest element.
SD_HD void sd_dense_flush(sd_dense &d, sd_sink &k, int start0, int a0)
{
    for (int start = start0;; ++start) {
        int oldest = -1;
        for (int j = d.wn - 1; j >= 0; --j) if (d.slot[j] & SD_SLOT_VALID) { oldest = j; break; }
        if (oldest < 0) break;
        if (a0 - oldest >= start) continue;
        sd_sink_put(k, a0 - oldest, a0 - oldest + sd_slot_flen(d.slot[oldest]));
        for (int j = 0; j < d.wn; ++j) if (a0 - j < start) d.slot[j] = 0;
    }
}

// an item whose bytes [p0, c1) are all A/C/G/T (the caller checks): same events as sd_run_item
template <class Fetch>
SD_HD void sd_run_item_dense(Fetch &fetch, int l_seq, int c0, int c1, uint32_t flags, int T, int W, sd_dense &d, sd_sink &k)
{
    const bool quiet = (flags & SD_ITEM_QUIET) != 0;
    int p0 = quiet ? c0 - (W + 2) : c0 - 2 * W - (W + 2);       // = sd_warm_quiet / sd_warm_start on ACGT-only sequence
    if (p0 < 0 || (!quiet && c0 - 2 * W <= 0)) p0 = 0;
    d.wn = d.L = d.rw = 0;
    unsigned t = 0;
    const int stop = c1 < l_seq ? c1 : l_seq;
    for (int i = p0; i < stop; ++i) {
        if (i == c0) k.on = 1;
        t = (t << 2 | (unsigned)sd_nt4(fetch(i))) & 63u;
        const int l = i - p0 + 1;
        if (l >= 3) sd_dense_step(d, k, p0 + (l - W > 0 ? l - W : 0), (int)t, T, W, !(quiet && i < c0));
    }
    if (c1 >= l_seq) {
        k.on = 1;
        const int l = l_seq - p0;
        if (l >= 3) sd_dense_flush(d, k, (l - W + 1 > 0 ? l - W + 1 : 0) + (l_seq + 1 - l), l_seq - 3);
    }
    sd_sink_close(k);
}

// --------------------------------------------------------------------------------------------
// Two-phase execution (W <= 64, floor(2T/10) == 4: the defaults).
//
// shift_window() does not depend on the perfect-interval list P at all: the window w, its counts, rw and L evolve from
// the triplets alone, and find_perfect() is only ever called at steps where  rw*10 > L*T  (:149) -- on ordinary
// sequence well under 1 % of the steps, in short bursts.  Between two such steps more than a window apart P is empty
// (an entry lives at most W-2 emitting steps, and a non-ACGT byte empties P at once), so nothing there can be saved.
//
// Phase 1, the SCOUT, therefore runs the window half alone over every base, in a form without a single
// data-dependent loop, and records WHERE the trigger fires.  Instead of the suffix counts cv[] and the shrink loop
// (:79-85) it keeps, per triplet value, the window count and the ring positions of the last four occurrences in one
// 32-bit word:  v is the longest suffix in which no triplet occurs more than four times, so pushing t shortens it
// exactly when t then occurs five times inside it, and the new v starts right after the fourth-previous occurrence:
//        L <- min(L + 1, wn, distance to the 4th previous occurrence of t  [if t is now >= 5 times in the window]).
// rv and cv[] are not needed for the trigger.  (sd_scout_* below; the test is bit-identical to :149, checked
// against the full machine in tests/sim.)
//
// How long an entry lives: inserted at step j it has a start in [ws(j), ws(j) + W-3] and leaves when the window start
// passes it.  In the normal phase (run length l >= W) the window start advances with every step: gone by j + W-2.  In the
// STALE phase -- fewer than W bases after a non-ACGT byte, where ws is still pinned to the first base after it while the
// (stale) window already holds W-2 triplets (:146) -- the start can lie up to W-3 right of ws and ws only starts moving
// when l reaches W: gone by j + (W - l) + W-2 < j + 2W.  (This is the quirk that lets intervals cover bases that are not
// low complexity at all.)  So a trigger at i keeps the blocks of i and i + 64 busy, and of i + 128 too in the stale phase.
//
// Phase 2 runs the full machine only over ITEMS: maximal runs of 64-base blocks that hold a trigger position i or lie
// within its drain, cut every SD_ITEM_MAX bases.  An item whose preceding block is quiet
// starts with P empty for certain, so its warm-up only has to rebuild the window (W-2 emitted triplets, window half
// only); an item cut out of a longer run uses the general exact start of section 2 above.  Items own their save events
// by time as before, and since starts and finishes of different runs of blocks are more than a window apart, only
// items of the same run are folded across their seams.
// --------------------------------------------------------------------------------------------
#if !defined(SD_WIDE)
#define SD_BLK 64                 /* activity granularity (bases) */
#define SD_ITEM_MAX 512           /* longest item (bases); >= 4W so that a seam fold never looks past one item.  Short items keep
                                     the serial chain of a task inside a long low-complexity run short (it bounds the kernel's tail) */

struct sd_scout {
    uint32_t wn, L, rw, e;        // window size, suffix length, window score, emitted-triplet counter (ring position = e & 63)
    int l;                        // current ACGT run length
    unsigned t;
};

// word of triplet value x: count:8 << 24 | p4:6 << 18 | p3:6 << 12 | p2:6 << 6 | p1:6   (p1 = ring position of the
// most recent occurrence).  SD_SW(x) / SD_SR(i) are the storage accessors: plain arrays on the host, shared-memory
// columns on the device.
template <class Words, class Ring>
SD_HD void sd_scout_reset(sd_scout &s, Words &words, Ring &ring)
{
    s.wn = s.L = s.rw = s.e = 0; s.l = 0; s.t = 0;
    for (int i = 0; i < 64; ++i) words(i) = 0;
    (void)ring;
}

// one emitted triplet; returns whether the reference would call find_perfect with something to examine
template <class Words, class Ring>
SD_HD bool sd_scout_push(sd_scout &s, Words &words, Ring &ring, uint32_t t, int T, int W)
{
    if (s.wn >= (uint32_t)(W - 2)) {
        const uint32_t x = ring((s.e - s.wn) & 63u);
        const uint32_t wx = words(x) - (1u << 24);
        words(x) = wx;
        s.rw -= wx >> 24;
        --s.wn;
    }
    if (s.L > s.wn) s.L = s.wn;
    ring(s.e & 63u) = (uint8_t)t;
    ++s.wn;
    const uint32_t w = words(t);
    const uint32_t c = w >> 24;
    s.rw += c;
    const uint32_t d4 = (s.e - (w >> 18)) & 63u;
    ++s.L;
    if (c >= 4u && d4 < s.L) s.L = d4;
    words(t) = ((c + 1u) << 24) | ((w << 6) & 0xFFFFC0u) | (s.e & 63u);
    ++s.e;
    return s.rw * 10u > s.L * (uint32_t)T && s.L < s.wn;
}

SD_HD bool sd_scout_supported(int T, int W) { return W >= 3 && W <= 64 && T > 0 && (T << 1) / 10 == 4; }
#endif

// --------------------------------------------------------------------------------------------
// Seam fold.  Chunk j (k-th chunk of its record, covering positions [k*C, ...)) left its list
// R_j in slots[j*cap .. j*cap+n[j]).  The record's result is the fold of R_0, R_1, ... with the
// reference's rule (start <= last.finish: only the finish grows; else append).
//
// A save at time i has start in [i-W, i+W) and finish <= i+W.  So (a) the finish of the interval
// that is `last` when chunk j begins matters only if it is >= c0-W, which is decided by saves in
// [c0-2W, c0), and (b) those saves fold the same way whatever happened before c0-4W.  Folding
// the lists of the chunks that start in [c0-4W, c0) from an empty state therefore yields the
// finish that chunk j's leading intervals are tested against.
// --------------------------------------------------------------------------------------------
#define SD_IV_START(x)  ((int)((x) >> 32))
#define SD_IV_FINISH(x) ((int)(uint32_t)(x))

// finish of the reference's last saved interval when chunk j (k-th of its record) begins; -1 = none that matters
SD_HD int sd_incoming_finish(const uint64_t *slots, const uint32_t *n, uint32_t cap, uint32_t j, uint32_t k, int C, int W)
{
    uint32_t back = (uint32_t)((4 * W + C - 1) / C);
    if (back > k) back = k;
    int F = -1;
    for (uint32_t jj = j - back; jj < j; ++jj)
        for (uint32_t a = 0; a < n[jj]; ++a) {
            const uint64_t iv = slots[(uint64_t)jj * cap + a];
            if (F >= 0 && SD_IV_START(iv) <= F) { if (SD_IV_FINISH(iv) > F) F = SD_IV_FINISH(iv); }
            else F = SD_IV_FINISH(iv);
        }
    return F;
}

// number of leading intervals of R_j swallowed by the incoming last interval
SD_HD uint32_t sd_absorbed(const uint64_t *slots, const uint32_t *n, uint32_t cap, uint32_t j, uint32_t k, int C, int W)
{
    int F = sd_incoming_finish(slots, n, cap, j, k, C, W);
    uint32_t a = 0;
    while (a < n[j] && F >= 0 && SD_IV_START(slots[(uint64_t)j * cap + a]) <= F) {
        const int f = SD_IV_FINISH(slots[(uint64_t)j * cap + a]);
        if (f > F) F = f;
        ++a;
    }
    return a;
}

SD_HD uint32_t sd_gather_count(const uint64_t *slots, const uint32_t *n, uint32_t cap, uint32_t j, uint32_t k, int C, int W)
{
    return n[j] - sd_absorbed(slots, n, cap, j, k, C, W);
}

// writes the intervals chunk j contributes; its last one keeps growing through the following
// chunks of the record for as long as their leading intervals are swallowed by it.
SD_HD void sd_gather_write(const uint64_t *slots, const uint32_t *n, uint32_t cap, uint32_t j, uint32_t k,
                           uint32_t n_chunks_rec, int C, int W, uint64_t *dst)
{
    const uint32_t a0 = sd_absorbed(slots, n, cap, j, k, C, W);
    for (uint32_t a = a0; a < n[j]; ++a) {
        uint64_t iv = slots[(uint64_t)j * cap + a];
        if (a == n[j] - 1) {
            int F = SD_IV_FINISH(iv);
            int open = 1;
            for (uint32_t kk = k + 1; open && kk < n_chunks_rec; ++kk) {
                if (F < (int)kk * C - W) break;                 // nothing later can start at or before F
                const uint32_t jj = j + (kk - k);
                for (uint32_t b = 0; b < n[jj]; ++b) {
                    const uint64_t nx = slots[(uint64_t)jj * cap + b];
                    if (SD_IV_START(nx) <= F) { if (SD_IV_FINISH(nx) > F) F = SD_IV_FINISH(nx); }
                    else { open = 0; break; }
                }
            }
            iv = (uint64_t)(uint32_t)SD_IV_START(iv) << 32 | (uint32_t)F;
        }
        *dst++ = iv;
    }
}

// ---- the same fold over ITEMS (two-phase execution): item j covers [c0[j], c1[j]) of its record, its list is
// slots[off[j] .. off[j] + n[j]), and flags[j] & SD_ITEM_CHAIN says that item j-1 ends exactly where j begins.
// Items are at most SD_ITEM_MAX >= 4W long but the first one of a run may be short, so the look-back walks the
// chain until it has covered [c0 - 4W, c0).
SD_HD int sd_item_incoming_finish(const uint64_t *slots, const uint32_t *off, const uint32_t *n, const uint32_t *c0, const uint32_t *flags,
                                  uint32_t j, int W)
{
    uint32_t jj = j;
    while ((flags[jj] & SD_ITEM_CHAIN) && (int)c0[jj] > (int)c0[j] - 4 * W) --jj;      // (jj > 0 whenever its CHAIN bit is set)
    int F = -1;
    for (; jj < j; ++jj)
        for (uint32_t a = 0; a < n[jj]; ++a) {
            const uint64_t iv = slots[(uint64_t)off[jj] + a];
            if (F >= 0 && SD_IV_START(iv) <= F) { if (SD_IV_FINISH(iv) > F) F = SD_IV_FINISH(iv); }
            else F = SD_IV_FINISH(iv);
        }
    return F;
}

SD_HD uint32_t sd_item_absorbed(const uint64_t *slots, const uint32_t *off, const uint32_t *n, const uint32_t *c0, const uint32_t *flags,
                                uint32_t j, int W)
{
    if (!(flags[j] & SD_ITEM_CHAIN)) return 0;
    int F = sd_item_incoming_finish(slots, off, n, c0, flags, j, W);
    uint32_t a = 0;
    while (a < n[j] && F >= 0 && SD_IV_START(slots[(uint64_t)off[j] + a]) <= F) {
        const int f = SD_IV_FINISH(slots[(uint64_t)off[j] + a]);
        if (f > F) F = f;
        ++a;
    }
    return a;
}

SD_HD uint32_t sd_item_gather_count(const uint64_t *slots, const uint32_t *off, const uint32_t *n, const uint32_t *c0, const uint32_t *flags,
                                    uint32_t j, int W)
{
    return n[j] - sd_item_absorbed(slots, off, n, c0, flags, j, W);
}

SD_HD void sd_item_gather_write(const uint64_t *slots, const uint32_t *off, const uint32_t *n, const uint32_t *c0, const uint32_t *flags,
                                uint32_t j, uint32_t n_items, int W, uint64_t *dst)
{
    const uint32_t a0 = sd_item_absorbed(slots, off, n, c0, flags, j, W);
    for (uint32_t a = a0; a < n[j]; ++a) {
        uint64_t iv = slots[(uint64_t)off[j] + a];
        if (a == n[j] - 1) {
            int F = SD_IV_FINISH(iv);
            int open = 1;
            for (uint32_t jj = j + 1; open && jj < n_items && (flags[jj] & SD_ITEM_CHAIN); ++jj) {
                if (F < (int)c0[jj] - W) break;                 // nothing later can start at or before F
                for (uint32_t b = 0; b < n[jj]; ++b) {
                    const uint64_t nx = slots[(uint64_t)off[jj] + b];
                    if (SD_IV_START(nx) <= F) { if (SD_IV_FINISH(nx) > F) F = SD_IV_FINISH(nx); }
                    else { open = 0; break; }
                }
            }
            iv = (uint64_t)(uint32_t)SD_IV_START(iv) << 32 | (uint32_t)F;
        }
        *dst++ = iv;
    }
}

}  // namespace SD_NS
