// cornetto_b200/csrc/shard.cu -- splitting a set of records over several GPUs and putting the results back
// into file order (host code; no CUDA call).
//
// The reference has one thread and one address space: records are scanned in file order and printed as they
// come (src/find_telomere.c:101-105, src/sdust/sdust.c:196-203).  On several B200s the unit of distribution is
// the RECORD: the scans never look across a record boundary, so shards are independent, need no halo and no
// collective, and the only thing left to do afterwards is to interleave the per-shard result lists back into
// file order.  (A record is never cut: 48 contigs of a diploid human assembly balance over 8 GPUs to within a
// few percent, see corn_shard_plan(); a single-contig input would stay on one GPU.)
#include <algorithm>
#include <vector>

#include "corn_internal.cuh"

// Longest-processing-time-first: records by decreasing length (ties: file order), each onto the least loaded
// shard (ties: lowest shard).  Deterministic, so every process of a job computes the same plan on its own.
extern "C" int corn_shard_plan(const uint32_t *length, uint32_t n_rec, uint32_t n_shards, uint32_t *shard_of)
{
    if (n_shards == 0 || (n_rec && (!length || !shard_of))) return CORN_E_ARG;
    std::vector<uint32_t> order(n_rec);
    for (uint32_t i = 0; i < n_rec; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return length[a] > length[b]; });
    std::vector<uint64_t> load(n_shards, 0);
    for (uint32_t i = 0; i < n_rec; ++i) {
        const uint32_t r = order[i];
        uint32_t best = 0;
        for (uint32_t s = 1; s < n_shards; ++s) if (load[s] < load[best]) best = s;
        shard_of[r] = best;
        load[best] += (uint64_t)length[r] + 2 * CORN_ALIGN;       // what the record occupies in a batch, padding included
    }
    return CORN_OK;
}

// local[r] = index of record r inside its shard's batch (records of a shard keep their file order);
// count[s] = records of shard s
extern "C" int corn_shard_local_index(const uint32_t *shard_of, uint32_t n_rec, uint32_t n_shards, uint32_t *local, uint32_t *count)
{
    if (n_shards == 0 || (n_rec && (!shard_of || !local)) || !count) return CORN_E_ARG;
    for (uint32_t s = 0; s < n_shards; ++s) count[s] = 0;
    for (uint32_t r = 0; r < n_rec; ++r) {
        if (shard_of[r] >= n_shards) return CORN_E_ARG;
        local[r] = count[shard_of[r]]++;
    }
    return CORN_OK;
}

// runs[s][0 .. n_run[s]) = result of shard s (rec = index inside the shard, in the reference's print order)
// -> out[0 .. sum n_run) in file order with rec = global record index
extern "C" int corn_shard_merge_runs(const corn_run_t *const *runs, const uint64_t *n_run, const uint32_t *shard_of,
                                     uint32_t n_rec, uint32_t n_shards, corn_run_t *out)
{
    if (n_shards == 0 || !n_run || !runs) return CORN_E_ARG;
    std::vector<uint32_t> local(n_rec ? n_rec : 1), count(n_shards);
    CORN_TRY(corn_shard_local_index(shard_of, n_rec, n_shards, local.data(), count.data()));
    std::vector<uint64_t> cur(n_shards, 0);
    uint64_t o = 0;
    for (uint32_t r = 0; r < n_rec; ++r) {
        const uint32_t s = shard_of[r];
        while (cur[s] < n_run[s] && runs[s][cur[s]].rec == local[r]) {
            out[o] = runs[s][cur[s]++];
            out[o++].rec = r;
        }
    }
    for (uint32_t s = 0; s < n_shards; ++s) if (cur[s] != n_run[s]) return CORN_E_LAYOUT;     // a list was not in record order
    return CORN_OK;
}

// iv[s], rec_first[s][0 .. count[s]] = sdust result of shard s -> out_iv in file order, out_first[0 .. n_rec]
extern "C" int corn_shard_merge_intervals(const uint64_t *const *iv, const uint64_t *const *rec_first, const uint32_t *shard_of,
                                          uint32_t n_rec, uint32_t n_shards, uint64_t *out_iv, uint64_t *out_first)
{
    if (n_shards == 0 || !iv || !rec_first || !out_first) return CORN_E_ARG;
    std::vector<uint32_t> local(n_rec ? n_rec : 1), count(n_shards);
    CORN_TRY(corn_shard_local_index(shard_of, n_rec, n_shards, local.data(), count.data()));
    uint64_t o = 0;
    for (uint32_t r = 0; r < n_rec; ++r) {
        const uint32_t s = shard_of[r];
        out_first[r] = o;
        for (uint64_t k = rec_first[s][local[r]]; k < rec_first[s][local[r] + 1]; ++k) out_iv[o++] = iv[s][k];
    }
    out_first[n_rec] = o;
    return CORN_OK;
}
