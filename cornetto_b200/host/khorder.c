/* cornetto_b200/host/khorder.c -- iteration order of a klib khash string map.
 *
 * telobreaks prints its contigs by walking final_scaffold_map from bucket 0 upwards
 * (src/telomere_breaks.c:133), so the output order is the khash bucket order of the lens-file
 * names.  This reproduces the table geometry only -- which key ends in which bucket:
 *   hash        X31 over the (signed) chars                          src/khash.h:395-400
 *   table       power of two, at least 4 buckets, grown when n_occupied reaches
 *               (int)(n_buckets * 0.77 + 0.5)                         src/khash.h:191,303-311,344
 *   probing     i = (i + ++step) & mask                               src/khash.h:238,327
 *   growth      keys are re-inserted scanning the old buckets upwards; an insertion that lands
 *               on an old bucket still holding an un-moved key evicts that key, which is
 *               re-inserted next (the "kick-out" loop)                src/khash.h:268-294
 */
#include "cornetto.h"

typedef struct {
    uint32_t n_buckets, size, n_occupied, upper;
    long    *key;          /* index into names[], -1 = empty bucket */
} table_t;

static uint32_t hash_x31(const char *s)
{
    uint32_t h = (uint32_t)(int)(signed char)*s;
    if (h) for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)(int)(signed char)*s;
    return h;
}

static uint32_t round_up_pow2(uint32_t x)
{
    --x; x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16;
    return x + 1;
}

static uint32_t probe_free(const long *key, uint32_t mask, uint32_t h)
{
    uint32_t i = h & mask, step = 0;
    while (key[i] >= 0) i = (i + (++step)) & mask;
    return i;
}

static void table_resize(table_t *t, uint32_t wanted, const char *const *names)
{
    uint32_t nb = round_up_pow2(wanted);
    if (nb < 4) nb = 4;
    if (t->size >= (uint32_t)(nb * 0.77 + 0.5)) return;
    long *fresh = (long *)malloc(sizeof(long) * nb);
    unsigned char *gone = (unsigned char *)calloc(t->n_buckets ? t->n_buckets : 1, 1);
    CORN_MALLOC_CHK(fresh); CORN_MALLOC_CHK(gone);
    for (uint32_t i = 0; i < nb; ++i) fresh[i] = -1;
    for (uint32_t j = 0; j < t->n_buckets; ++j) {
        if (t->key[j] < 0 || gone[j]) continue;
        long k = t->key[j];
        gone[j] = 1;
        for (;;) {
            const uint32_t i = probe_free(fresh, nb - 1, hash_x31(names[k]));
            fresh[i] = k;
            if (i < t->n_buckets && t->key[i] >= 0 && !gone[i]) { k = t->key[i]; gone[i] = 1; }
            else break;
        }
    }
    free(gone); free(t->key);
    t->key = fresh;
    t->n_buckets = nb;
    t->n_occupied = t->size;
    t->upper = (uint32_t)(nb * 0.77 + 0.5);
}

size_t khash_str_order(const char *const *names, size_t n, size_t *order)
{
    table_t t;
    memset(&t, 0, sizeof t);
    for (size_t k = 0; k < n; ++k) {
        if (t.n_occupied >= t.upper)
            table_resize(&t, t.n_buckets > (t.size << 1) ? t.n_buckets - 1 : t.n_buckets + 1, names);
        const uint32_t mask = t.n_buckets - 1;
        uint32_t i = hash_x31(names[k]) & mask, step = 0;
        const uint32_t first = i;
        int found = 0;
        while (t.key[i] >= 0) {
            if (strcmp(names[t.key[i]], names[k]) == 0) { found = 1; break; }
            i = (i + (++step)) & mask;
            if (i == first) break;
        }
        if (!found && t.key[i] < 0) { t.key[i] = (long)k; ++t.size; ++t.n_occupied; }
    }
    size_t m = 0;
    for (uint32_t i = 0; i < t.n_buckets; ++i) if (t.key[i] >= 0) order[m++] = (size_t)t.key[i];
    free(t.key);
    return m;
}
