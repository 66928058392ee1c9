/* include/corn_bench.h -- synthetic-input helpers for bench.py and the full-size GPU tests.
 *
 * NOT part of the drop-in ABI (include/corn_gpu.h): nothing here is on the product path.  The
 * reference has no equivalent; BASELINE.json asks for synthetic assemblies of 3.1 / 6.2 Gb, which
 * are generated in place in HBM (seeded, counter based) instead of being shipped as files.
 */
#ifndef CORN_BENCH_H
#define CORN_BENCH_H

#include "corn_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { CORN_FEAT_TANDEM = 0, CORN_FEAT_NGAP = 1, CORN_FEAT_LOWER = 2 };

typedef struct corn_feature {
    uint32_t rec, start, len;   /* [start, start+len) on record rec (clipped to the record) */
    uint32_t kind;              /* CORN_FEAT_* */
    uint32_t period;            /* tandem: unit length 1..8 */
    uint32_t seed;              /* tandem: per-feature seed for the variant copies */
    float    p_variant;         /* tandem: probability that a copy carries one substituted base */
    uint8_t  unit[8];
    uint32_t _pad;
} corn_feature_t;

/* every base of every record <- uniform A/C/G/T from a counter-based generator keyed by
 * (seed, byte position in the batch); padding stays 0x00. */
int corn_bench_fill_random(corn_ctx_t *ctx, corn_dbatch_t *db, uint64_t seed);
/* same, keyed by (seed, rec_id[record], position inside the record): a record gets the same bytes whichever batch
 * (whichever GPU's shard) holds it.  rec_id: [n_rec] host array of global record numbers. */
int corn_bench_fill_random_rec(corn_ctx_t *ctx, corn_dbatch_t *db, uint64_t seed, const uint32_t *rec_id);
/* overlays features; the features of ONE call are applied concurrently (overlapping ones race,
 * any outcome being a valid sequence); successive calls are ordered. */
int corn_bench_apply_features(corn_ctx_t *ctx, corn_dbatch_t *db, const corn_feature_t *feat, uint32_t n_feat);
/* copies the whole padded sequence area (corn_gpu_dbatch_bytes() bytes) to host memory */
int corn_bench_download_all(corn_ctx_t *ctx, const corn_dbatch_t *db, uint8_t *dst);
/* writes a buffer larger than L2 (bytes >= 256 MiB) so the next kernel starts cold */
int corn_bench_flush_l2(corn_ctx_t *ctx);

#ifdef __cplusplus
}
#endif
#endif
