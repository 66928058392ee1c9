// cornetto_b200/csrc/synth.cu -- in-HBM synthetic assemblies for bench.py (include/corn_bench.h).
// Not on the product path.
#include "corn_internal.cuh"
#include "../../include/corn_bench.h"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// one thread per 16-byte block of the batch
// rec_id == NULL: keyed by (seed, block index in the batch).  rec_id != NULL: keyed by (seed, rec_id[record], block
// index inside the record), so a record has the same bytes whichever batch -- whichever GPU's shard -- it is in.
__global__ void __launch_bounds__(256) k_fill_random(uint8_t *seq, uint64_t n_blocks16, const uint32_t *__restrict__ rec_off,
                                                     const uint32_t *__restrict__ rec_len, uint32_t n_rec, uint64_t seed,
                                                     const uint32_t *__restrict__ rec_id)
{
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_blocks16) return;
    const uint32_t pos = (uint32_t)(g << 4);
    const uint32_t rec = corn_upper_bound(rec_off, n_rec, pos) - 1;
    const uint32_t r0 = rec_off[rec], r1 = r0 + rec_len[rec];
    uint64_t h = rec_id ? mix64(mix64(seed ^ ((uint64_t)rec_id[rec] << 40)) ^ ((uint64_t)((pos - r0) >> 4) * 0xD1342543DE82EF95ull))
                        : mix64(seed ^ (g * 0xD1342543DE82EF95ull));
    uint32_t w[4] = { 0, 0, 0, 0 };
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t p = pos + i;
        const uint32_t c = (uint32_t)(h >> (2 * i)) & 3u;
        const uint32_t b = (p >= r0 && p < r1) ? (0x54474341u >> (8 * c)) & 0xFFu : 0u;   // "ACGT"
        w[i >> 2] |= b << (8 * (i & 3));
    }
    *(uint4 *)(seq + pos) = make_uint4(w[0], w[1], w[2], w[3]);
}

// one block per feature
__global__ void __launch_bounds__(128) k_apply_features(uint8_t *seq, const uint32_t *__restrict__ rec_off,
                                                        const uint32_t *__restrict__ rec_len, uint32_t n_rec,
                                                        const corn_feature_t *__restrict__ feats)
{
    const corn_feature_t f = feats[blockIdx.x];
    if (f.rec >= n_rec) return;
    const uint32_t len = rec_len[f.rec];
    if (f.start >= len) return;
    const uint32_t n = min(f.len, len - f.start);
    uint8_t *p = seq + rec_off[f.rec] + f.start;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        if (f.kind == CORN_FEAT_NGAP) p[i] = 'N';
        else if (f.kind == CORN_FEAT_LOWER) { uint8_t b = p[i]; if (b >= 'A' && b <= 'Z') p[i] = b + 32; }
        else {
            const uint32_t copy = i / f.period, k = i % f.period;
            uint8_t b = f.unit[k];
            const uint64_t h = mix64(((uint64_t)f.seed << 32) ^ copy);
            if ((float)(h & 0xFFFFFF) * (1.0f / 16777216.0f) < f.p_variant && (uint32_t)((h >> 24) % f.period) == k)
                b = (uint8_t)((0x54474341u >> (8 * ((h >> 40) & 3))) & 0xFF);
            p[i] = b;
        }
    }
}

__global__ void k_flush(uint32_t *p, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}

}  // namespace

extern "C" int corn_bench_fill_random(corn_ctx_t *ctx, corn_dbatch_t *db, uint64_t seed)
{
    if (!ctx || !db) return CORN_E_ARG;
    CORN_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t nb = db->total_bytes / 16;
    if (nb == 0 || db->n_rec == 0) return CORN_OK;
    k_fill_random<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>(db->d_seq, nb, db->d_rec_off, db->d_rec_len, db->n_rec, seed, NULL);
    CORN_LAUNCH_CHECK(ctx);
    CORN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CORN_OK;
}

extern "C" int corn_bench_fill_random_rec(corn_ctx_t *ctx, corn_dbatch_t *db, uint64_t seed, const uint32_t *rec_id)
{
    if (!ctx || !db || !rec_id) return CORN_E_ARG;
    CORN_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t nb = db->total_bytes / 16;
    if (nb == 0 || db->n_rec == 0) return CORN_OK;
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->wins, (size_t)db->n_rec * sizeof(uint32_t)));
    CORN_CUDA(ctx, cudaMemcpyAsync(ctx->wins.p, rec_id, (size_t)db->n_rec * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    k_fill_random<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>(db->d_seq, nb, db->d_rec_off, db->d_rec_len, db->n_rec, seed,
                                                                         (const uint32_t *)ctx->wins.p);
    CORN_LAUNCH_CHECK(ctx);
    CORN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CORN_OK;
}

extern "C" int corn_bench_apply_features(corn_ctx_t *ctx, corn_dbatch_t *db, const corn_feature_t *feat, uint32_t n_feat)
{
    if (!ctx || !db || (n_feat && !feat)) return CORN_E_ARG;
    if (n_feat == 0) return CORN_OK;
    CORN_CUDA(ctx, cudaSetDevice(ctx->device));
    for (uint32_t i = 0; i < n_feat; ++i)
        if (feat[i].kind == CORN_FEAT_TANDEM && (feat[i].period == 0 || feat[i].period > 8))
            return corn_set_err(ctx, CORN_E_ARG, "feature %u: period %u", i, feat[i].period);
    // features of one call may run concurrently: callers pass non-overlapping ones per call
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->wins, (size_t)n_feat * sizeof(corn_feature_t)));
    CORN_CUDA(ctx, cudaMemcpyAsync(ctx->wins.p, feat, (size_t)n_feat * sizeof(corn_feature_t), cudaMemcpyHostToDevice, ctx->stream));
    k_apply_features<<<n_feat, 128, 0, ctx->stream>>>(db->d_seq, db->d_rec_off, db->d_rec_len, db->n_rec, (const corn_feature_t *)ctx->wins.p);
    CORN_LAUNCH_CHECK(ctx);
    CORN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CORN_OK;
}

extern "C" int corn_bench_download_all(corn_ctx_t *ctx, const corn_dbatch_t *db, uint8_t *dst)
{
    if (!ctx || !db || !dst) return CORN_E_ARG;
    CORN_CUDA(ctx, cudaSetDevice(ctx->device));
    CORN_CUDA(ctx, cudaMemcpyAsync(dst, db->d_seq, db->total_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CORN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CORN_OK;
}

extern "C" int corn_bench_flush_l2(corn_ctx_t *ctx)
{
    if (!ctx) return CORN_E_ARG;
    CORN_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)384 << 20;
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->bitmap, bytes));
    k_flush<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((uint32_t *)ctx->bitmap.p, bytes / 4);
    CORN_LAUNCH_CHECK(ctx);
    return CORN_OK;
}
