// tests/sim/sim_main.cpp -- CPU simulation of the device-side algorithms (TEST INFRASTRUCTURE).
//
// The kernels' arithmetic lives in host/device headers (cornetto_b200/csrc/*_core.cuh).  This
// program compiles those headers for the host and drives them the way the kernels do (lane
// chunks, thread-per-chunk sdust with warm start and seam merge), so the logic can be checked
// against the oracle on a machine without a GPU.  It is not part of the product and is never
// used as a fallback: the shipped library only contains the CUDA path.
//
//   sim coretest                       exhaustive/brute-force checks of the bit-plane primitives
//   sim telofind <fastx> [motif]       candidate masks -> verify -> start/end classification -> runs
//   sim sdust [-w W] [-t T] [-c C] <fastx>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../cornetto_b200/csrc/sdust_core.cuh"
#include "../../cornetto_b200/csrc/telofind_core.cuh"
#include "../../oracle/oracle.h"

using namespace SD_NS;      // sd_narrow, or sd_wide when built with -DSD_WIDE (16-bit counters, 64-bit slots, W <= 1024)

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

static int coretest()
{
    // 1. gather8 over every pair of bytes patterns (random words, exhaustive on the code bits)
    for (int it = 0; it < 2000000; ++it) {
        uint32_t w[8];
        for (int i = 0; i < 8; ++i) w[i] = (uint32_t)rnd();
        uint32_t p1, p2;
        corn_planes32(w, p1, p2);
        const uint8_t *b = (const uint8_t *)w;
        for (int j = 0; j < 32; ++j) {
            if (((p1 >> j) & 1) != ((b[j] >> 1) & 1u) || ((p2 >> j) & 1) != ((b[j] >> 2) & 1u)) {
                fprintf(stderr, "planes mismatch at iteration %d base %d\n", it, j);
                return 1;
            }
        }
    }
    // 2. match32 against a byte-wise code comparison, several motifs
    const char *motifs[] = { "TTAGGG", "CCCTAA", "A", "ACGT", "TTTAGGG", "AAAAAA", "TATATA", "GATTACAGATTACAGATTACAGATTACAGATT" };
    for (const char *mo : motifs) {
        corn_motif_info mi;
        corn_analyse_motif(mo, &mi);
        if (!mi.acgt) { fprintf(stderr, "motif %s not acgt?\n", mo); return 1; }
        for (int it = 0; it < 20000; ++it) {
            uint8_t buf[96];
            for (int i = 0; i < 96; ++i) {
                uint64_t r = rnd();
                // mostly ACGT in both cases, sometimes the motif itself, sometimes junk
                buf[i] = (r & 0xF00) == 0 ? (uint8_t)(r >> 16) : (uint8_t)("ACGTacgt"[(r >> 3) & 7]);
            }
            if (it & 1) for (int k = 0; k < 3; ++k) { int p = (int)(rnd() % (96 - mi.m)); memcpy(buf + p, (k & 1) ? mi.rev : mi.fwd, mi.m); }
            uint32_t p1, p2, n1, n2, mf, mr, w[16];
            memcpy(w, buf, 64);
            corn_planes32(w, p1, p2);
            corn_planes32(w + 8, n1, n2);
            corn_match32<0>(p1, p2, n1, n2, mi.fc, mi.rc, mi.m, mf, mr);
            for (int j = 0; j < 32; ++j) {
                int ef = 1, er = 1;
                for (int d = 0; d < mi.m; ++d) {
                    int code = (buf[j + d] >> 1) & 3;
                    ef &= code == corn_code_of((char)mi.fwd[d]);
                    er &= code == corn_code_of((char)mi.rev[d]);
                }
                if ((int)((mf >> j) & 1) != ef || (int)((mr >> j) & 1) != er) { fprintf(stderr, "match32 mismatch motif %s\n", mo); return 1; }
                // a verified candidate is a true occurrence and every true occurrence is a candidate
                bool tf = corn_occ_at(buf + j, mi.fwd, mi.m), tr = corn_occ_at(buf + j, mi.rev, mi.m);
                if ((tf && !ef) || (tr && !er)) { fprintf(stderr, "occurrence without candidate, motif %s\n", mo); return 1; }
            }
            if (mi.m == 6) {
                uint32_t mf6, mr6;
                corn_match32<6>(p1, p2, n1, n2, mi.fc, mi.rc, 6, mf6, mr6);
                if (mf6 != mf || mr6 != mr) { fprintf(stderr, "match32<6> differs\n"); return 1; }
            }
        }
    }
    printf("coretest ok\n");
    return 0;
}

// ---- telofind simulation: per record, exactly the kernels' logic but position by position ----
static int sim_telofind(const char *path, const char *motif)
{
    orc_rec_t *recs; size_t n;
    if (orc_read_fastx(path, &recs, &n) < 0) return 1;
    corn_motif_info mi;
    corn_analyse_motif(motif, &mi);
    for (size_t r = 0; r < n; ++r) {
        const size_t len = recs[r].len;
        // padded copy: guard before, zeros after (the HBM layout)
        std::vector<uint8_t> buf(1024 + len + 4096 + 64, 0);
        uint8_t *seq = buf.data() + 1024;
        memcpy(seq, recs[r].seq, len);
        for (int strand = 0; strand < 2; ++strand) {
            const uint8_t *pat = strand ? mi.rev : mi.fwd;
            const uint64_t codes = strand ? mi.rc : mi.fc;
            std::vector<uint32_t> starts, ends, occ;
            const size_t n_chunks = (len + 31) / 32 + 1;
            for (size_t c = 0; c < n_chunks; ++c) {
                uint32_t mask = 0;
                if (mi.acgt && mi.m <= CORN_MAX_FAST_MOTIF) {
                    uint32_t w[16], p1, p2, n1, n2, mf, mr;
                    memcpy(w, seq + c * 32, 64);
                    corn_planes32(w, p1, p2);
                    corn_planes32(w + 8, n1, n2);
                    corn_match32<0>(p1, p2, n1, n2, codes, codes, mi.m, mf, mr);
                    mask = mf;
                } else {
                    for (int b = 0; b < 32; ++b) if (corn_occ_at(seq + c * 32 + b, pat, mi.m)) mask |= 1u << b;
                }
                for (int b = 0; b < 32; ++b) {
                    if (!((mask >> b) & 1)) continue;
                    const uint8_t *p = seq + c * 32 + b;
                    if (!corn_occ_at(p, pat, mi.m)) continue;
                    uint32_t pos = (uint32_t)(c * 32 + b);
                    if (mi.bordered) { occ.push_back(pos); continue; }
                    if (!corn_occ_at(p - mi.m, pat, mi.m)) starts.push_back(pos);
                    if (!corn_occ_at(p + mi.m, pat, mi.m)) ends.push_back(pos + mi.m);
                }
            }
            if (!mi.bordered) {
                if (starts.size() != ends.size()) { fprintf(stderr, "start/end mismatch\n"); return 1; }
                for (size_t k = 0; k < starts.size(); ++k)
                    printf("%s\t%zu\t%d\t%u\t%u\t%u\n", recs[r].name, len, strand, starts[k], ends[k], ends[k] - starts[k]);
            } else {
                size_t i = 0, hi = occ.size();
                while (i < hi) {
                    uint32_t p = occ[i], q = p + mi.m; size_t j = i + 1;
                    for (;;) { while (j < hi && occ[j] < q) ++j; if (j < hi && occ[j] == q) { q += mi.m; ++j; } else break; }
                    printf("%s\t%zu\t%d\t%u\t%u\t%u\n", recs[r].name, len, strand, p, q, q - p);
                    i = j; while (i < hi && occ[i] <= q) ++i;
                }
            }
        }
    }
    orc_free_recs(recs, n);
    return 0;
}

// ---- sdust simulation: thread-per-chunk + seam merge ------------------------------------------
static int sim_sdust(const char *path, int T, int W, int C)
{
    orc_rec_t *recs; size_t n;
    if (orc_read_fastx(path, &recs, &n) < 0) return 1;
    const uint32_t cap = (uint32_t)(C + 2 * W) / 4 + 2;
    for (size_t r = 0; r < n; ++r) {
        const int len = (int)recs[r].len;
        const uint32_t nch = (uint32_t)((len + C - 1) / C);
        std::vector<uint64_t> slots((size_t)nch * cap + 1);
        std::vector<uint32_t> cnt(nch + 1, 0);
        for (uint32_t k = 0; k < nch; ++k) {
            uint8_t ring[SD_MAX_W];
            sd_cnt_t cw[64], cv[64];
            sd_slot_t slot[SD_MAX_W];
            sd_mem m = { ring, cw, cv, slot, 4 };
            const int c0 = (int)k * C, c1 = std::min(len, (int)(k + 1) * C);
            sd_sink sink;
            sd_sink_init(sink, slots.data() + (size_t)k * cap, cap);
            struct { const uint8_t *p; uint8_t operator()(int i) const { return p[i]; } } fetch = { (const uint8_t *)recs[r].seq };
            sd_run_chunk(fetch, len, c0, c1, T, W, m, sink);
            if (sink.overflow) { fprintf(stderr, "slot overflow\n"); return 1; }
            cnt[k] = sink.n;
        }
        for (uint32_t k = 0; k < nch; ++k) {
            uint32_t c = sd_gather_count(slots.data(), cnt.data(), cap, k, k, C, W);
            std::vector<uint64_t> out(c + 1);
            sd_gather_write(slots.data(), cnt.data(), cap, k, k, nch, C, W, out.data());
            for (uint32_t a = 0; a < c; ++a) printf("%s\t%d\t%d\n", recs[r].name, SD_IV_START(out[a]), SD_IV_FINISH(out[a]));
        }
    }
    orc_free_recs(recs, n);
#if defined(SD_SLACK_STATS)
    fprintf(stderr, "find_perfect triggers with an index to examine: %llu evaluated, %llu skipped by the slack bound\n", sd_stat_eval, sd_stat_skip);
#endif
    return 0;
}

#if !defined(SD_WIDE)
// ---- two-phase sdust simulation: scout (window half only, per chunk) -> active 64-base blocks -> items -> full
//      machine per item -> fold across item seams.  Mirrors csrc/sdust.cu's fast path step by step.
struct HostWords { uint32_t w[64]; uint32_t &operator()(uint32_t i) { return w[i]; } };
struct HostRing { uint8_t r[64]; uint8_t &operator()(uint32_t i) { return r[i]; } };

static unsigned long long st_pos, st_trig, st_active_blk, st_blk, st_items, st_item_pos, st_dense_items;

static int sim_sdust2(const char *path, int T, int W, int C)
{
    if (!sd_scout_supported(T, W)) { fprintf(stderr, "two-phase path needs W <= 64 and floor(2T/10) == 4\n"); return 2; }
    orc_rec_t *recs; size_t n;
    if (orc_read_fastx(path, &recs, &n) < 0) return 1;
    for (size_t r = 0; r < n; ++r) {
        const int len = (int)recs[r].len;
        const uint8_t *seq = (const uint8_t *)recs[r].seq;
        struct { const uint8_t *p; uint8_t operator()(int i) const { return p[i]; } } fetch = { seq };
        const int n_blk = (len + SD_BLK - 1) / SD_BLK;
        std::vector<uint8_t> active(n_blk + 2, 0);
        // phase 1: scout per chunk of C bases
        for (int c0 = 0; c0 < len; c0 += C) {
            const int c1 = std::min(len, c0 + C);
            HostWords words; HostRing ring;
            sd_scout sc;
            sd_scout_reset(sc, words, ring);
            const int p0 = sd_warm_quiet(fetch, c0, W);
            for (int i = p0; i < c1; ++i) {
                const int b = sd_nt4(seq[i]);
                if (b < 4) {
                    ++sc.l; sc.t = (sc.t << 2 | (unsigned)b) & 63u;
                    if (sc.l >= 3 && sd_scout_push(sc, words, ring, sc.t, T, W) && i >= c0) {
                        ++st_trig;
                        active[i / SD_BLK] = 1;
                        if (i / SD_BLK + 1 < n_blk) active[i / SD_BLK + 1] = 1;
                        if (sc.l < W && i / SD_BLK + 2 < n_blk) active[i / SD_BLK + 2] = 1;    // stale phase: longer drain (sdust_core.cuh)
                    }
                } else { sc.l = 0; sc.t = 0; }
            }
        }
        st_pos += len; st_blk += n_blk;
        // items
        std::vector<uint32_t> ic0, ic1, iflags;
        for (int j = 0; j < n_blk; ++j) {
            if (!active[j]) continue;
            ++st_active_blk;
            const bool prev = j > 0 && active[j - 1];
            if (prev && (j * SD_BLK) % SD_ITEM_MAX != 0) { ic1.back() = (uint32_t)std::min(len, (j + 1) * SD_BLK); continue; }
            ic0.push_back((uint32_t)(j * SD_BLK));
            ic1.push_back((uint32_t)std::min(len, (j + 1) * SD_BLK));
            iflags.push_back((prev ? SD_ITEM_CHAIN : 0u) | ((j > 0 && !prev) ? SD_ITEM_QUIET : 0u));
        }
        const uint32_t ni = (uint32_t)ic0.size();
        st_items += ni;
        // phase 2: full machine per item
        std::vector<uint32_t> off(ni + 1, 0), cnt(ni + 1, 0);
        for (uint32_t j = 0; j < ni; ++j) off[j + 1] = off[j] + (ic1[j] - ic0[j] + 2 * W) / 4 + 2;
        std::vector<uint64_t> slots(off[ni] + 1);
        for (uint32_t j = 0; j < ni; ++j) {
            uint8_t ringb[SD_MAX_W]; sd_cnt_t cw[64], cv[64]; sd_slot_t slot[SD_MAX_W];
            sd_mem m = { ringb, cw, cv, slot, 4 };
            sd_sink sink;
            sd_sink_init(sink, slots.data() + off[j], off[j + 1] - off[j]);
            // an item that reaches the last block of the record also owns the final flush (c1 >= len)
            // items without a non-ACGT byte in [p0, c1) may take the age-ordered machine (what the dense kernel runs);
            // SIM_DENSE=1 sends every such item there, so that the whole corpus checks it
            const int ic1j = (int)ic1[j] >= len ? len : (int)ic1[j];
            int lo = (iflags[j] & SD_ITEM_QUIET) ? (int)ic0[j] - (W + 2) : (int)ic0[j] - 2 * W - (W + 2);
            if (lo < 0) lo = 0;
            bool acgt = getenv("SIM_DENSE") != NULL;
            for (int q = lo; acgt && q < std::min(ic1j, len); ++q) acgt = sd_nt4(seq[q]) < 4;
            if (acgt) {
                sd_dense dm;
                sd_run_item_dense(fetch, len, (int)ic0[j], ic1j, iflags[j], T, W, dm, sink);
                ++st_dense_items;
            } else
            sd_run_item(fetch, len, (int)ic0[j], ic1j, iflags[j], T, W, m, sink);
            if (sink.overflow) { fprintf(stderr, "slot overflow\n"); return 1; }
            cnt[j] = sink.n;
            st_item_pos += ic1[j] - ic0[j];
        }
        for (uint32_t j = 0; j < ni; ++j) {
            const uint32_t c = sd_item_gather_count(slots.data(), off.data(), cnt.data(), ic0.data(), iflags.data(), j, W);
            std::vector<uint64_t> out(c + 1);
            sd_item_gather_write(slots.data(), off.data(), cnt.data(), ic0.data(), iflags.data(), j, ni, W, out.data());
            for (uint32_t a = 0; a < c; ++a) printf("%s\t%d\t%d\n", recs[r].name, SD_IV_START(out[a]), SD_IV_FINISH(out[a]));
        }
    }
    orc_free_recs(recs, n);
    fprintf(stderr, "positions %llu triggers %llu (%.3f%%) active blocks %llu of %llu (%.1f%%) items %llu item positions %llu (%.1f%%)\n",
            st_pos, st_trig, 100.0 * st_trig / (st_pos ? st_pos : 1), st_active_blk, st_blk, 100.0 * st_active_blk / (st_blk ? st_blk : 1),
            st_items, st_item_pos, 100.0 * st_item_pos / (st_pos ? st_pos : 1));
    if (st_dense_items) fprintf(stderr, "items run by the age-ordered machine: %llu\n", st_dense_items);
    return 0;
}
#endif

int main(int argc, char **argv)
{
    if (argc < 2) return 1;
    if (!strcmp(argv[1], "coretest")) return coretest();
    if (!strcmp(argv[1], "telofind") && argc >= 3) return sim_telofind(argv[2], argc >= 4 ? argv[3] : "TTAGGG");
    if (!strcmp(argv[1], "sdust")) {
        int W = 64, T = 20, C = 4096; const char *f = NULL;
        for (int i = 2; i < argc; ++i) {
            if (!strcmp(argv[i], "-w") && i + 1 < argc) W = atoi(argv[++i]);
            else if (!strcmp(argv[i], "-t") && i + 1 < argc) T = atoi(argv[++i]);
            else if (!strcmp(argv[i], "-c") && i + 1 < argc) C = atoi(argv[++i]);
            else f = argv[i];
        }
        if (!f) return 1;
        return sim_sdust(f, T, W, C);
    }
#if !defined(SD_WIDE)
    if (!strcmp(argv[1], "sdust2")) {
        int W = 64, T = 20, C = 4096; const char *f = NULL;
        for (int i = 2; i < argc; ++i) {
            if (!strcmp(argv[i], "-w") && i + 1 < argc) W = atoi(argv[++i]);
            else if (!strcmp(argv[i], "-t") && i + 1 < argc) T = atoi(argv[++i]);
            else if (!strcmp(argv[i], "-c") && i + 1 < argc) C = atoi(argv[++i]);
            else f = argv[i];
        }
        if (!f) return 1;
        return sim_sdust2(f, T, W, C);
    }
#endif
    return 1;
}
