/* cornetto_b200/host/telobreaks_main.c -- `cornetto telobreaks <lens> <sdust.bed> <telofind.tsv>`.
 *
 * Same contract as telomere_breaks_main(), src/telomere_breaks.c:47-172.  The reference keeps two
 * bitsets per contig and walks them bit by bit; the same answer is obtained here from interval
 * arithmetic on the (tiny) text inputs:
 *   mask[c]  = union of the sdust intervals of contig c, as maximal runs        (:76-91)
 *   a telomere line with matched length >= 24 (MIN_TEL, :10) qualifies when
 *   [max(start-100,0), min(end+100,len)) lies inside ONE run of mask[c]         (:99-112);
 *   it then contributes the part of the mask run(s) reached by the reference's left/right
 *   extension from start and end (:116-126);
 *   final[c] = union of the contributions; every maximal run [i,e) prints as
 *   "Found telomere positions max(i-1,0) to e-1 is a telomere in NAME of length LEN" (:136-144)
 *   contigs in khash bucket order of the lens-file insertions (:133, khorder.c).
 * This step touches a few thousand intervals and no sequence bytes, so it stays on the host. */
#include <ctype.h>

#include "cornetto.h"

#define MIN_TEL 24
#define LINE_CAP 2048

typedef struct { int s, e; } iv_t;
typedef struct { iv_t *a; size_t n, m; } ivv_t;

static void iv_push(ivv_t *v, int s, int e)
{
    if (v->n == v->m) { v->m = v->m ? v->m * 2 : 16; v->a = (iv_t *)realloc(v->a, v->m * sizeof(iv_t)); CORN_MALLOC_CHK(v->a); }
    v->a[v->n].s = s; v->a[v->n].e = e; ++v->n;
}
static int iv_cmp(const void *x, const void *y)
{
    const iv_t *a = (const iv_t *)x, *b = (const iv_t *)y;
    return a->s < b->s ? -1 : a->s > b->s ? 1 : (a->e > b->e) - (a->e < b->e);
}
/* sort + fuse overlapping or adjacent intervals: maximal runs of set bits */
static void iv_normalise(ivv_t *v)
{
    if (v->n == 0) return;
    qsort(v->a, v->n, sizeof(iv_t), iv_cmp);
    size_t k = 0;
    for (size_t i = 1; i < v->n; ++i) {
        if (v->a[i].s <= v->a[k].e) { if (v->a[i].e > v->a[k].e) v->a[k].e = v->a[i].e; }
        else v->a[++k] = v->a[i];
    }
    v->n = k + 1;
}
/* run containing position p (bit p set), or NULL */
static const iv_t *iv_find(const ivv_t *v, int p)
{
    size_t lo = 0, hi = v->n;
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (v->a[mid].s <= p) lo = mid + 1; else hi = mid; }
    if (lo == 0) return NULL;
    const iv_t *r = &v->a[lo - 1];
    return p < r->e ? r : NULL;
}

static int split_ws(char *line, char **tok, int maxtok)
{
    int n = 0;
    char *p = line;
    while (n < maxtok) {
        while (*p && isspace((unsigned char)*p)) ++p;
        if (!*p) break;
        tok[n++] = p;
        while (*p && !isspace((unsigned char)*p)) ++p;
        if (*p) *p++ = 0;
    }
    return n;
}

typedef struct { char *name; int length; ivv_t mask, fin; } contig_t;

static int name_cmp(const void *a, const void *b) { return strcmp((*(contig_t *const *)a)->name, (*(contig_t *const *)b)->name); }

static contig_t *lookup(contig_t **sorted, size_t n, const char *name)
{
    size_t lo = 0, hi = n;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        int c = strcmp(sorted[mid]->name, name);
        if (c == 0) return sorted[mid];
        if (c < 0) lo = mid + 1; else hi = mid;
    }
    return NULL;
}

int telomere_breaks_main(int argc, char *argv[])
{
    if (argc < 4) {
        fprintf(stderr, "Usage: telobreaks <lens_file> <sdust_file> <telomere_file>\n");
        return EXIT_FAILURE;
    }
    FILE *fp = fopen(argv[1], "r");
    CORN_F_CHK(fp, argv[1]);
    char line[LINE_CAP];
    char **names = NULL; int *lengths = NULL; size_t n = 0, m = 0;
    while (fgets(line, sizeof line, fp)) {
        char *t[2];
        if (split_ws(line, t, 2) < 2) continue;
        if (n == m) { m = m ? m * 2 : 64; names = (char **)realloc(names, m * sizeof(char *)); lengths = (int *)realloc(lengths, m * sizeof(int)); CORN_MALLOC_CHK(names); CORN_MALLOC_CHK(lengths); }
        names[n] = strdup(t[0]); lengths[n] = atoi(t[1]); ++n;
    }
    fclose(fp);

    size_t *order = (size_t *)malloc(sizeof(size_t) * (n ? n : 1));
    CORN_MALLOC_CHK(order);
    const size_t nd = khash_str_order((const char *const *)names, n, order);
    contig_t *ct = (contig_t *)calloc(nd ? nd : 1, sizeof(contig_t));
    contig_t **sorted = (contig_t **)malloc(sizeof(contig_t *) * (nd ? nd : 1));
    CORN_MALLOC_CHK(ct); CORN_MALLOC_CHK(sorted);
    for (size_t k = 0; k < nd; ++k) { ct[k].name = names[order[k]]; ct[k].length = lengths[order[k]]; sorted[k] = &ct[k]; }
    qsort(sorted, nd, sizeof(contig_t *), name_cmp);
    /* a repeated name overwrites the map value (:66-71): the last length wins */
    for (size_t j = 0; j < n; ++j) { contig_t *c = lookup(sorted, nd, names[j]); if (c) c->length = lengths[j]; }

    fp = fopen(argv[2], "r");
    CORN_F_CHK(fp, argv[2]);
    while (fgets(line, sizeof line, fp)) {
        char *t[3];
        if (split_ws(line, t, 3) < 3) continue;
        contig_t *c = lookup(sorted, nd, t[0]);
        if (!c) continue;
        int s = atoi(t[1]), e = atoi(t[2]);
        if (s < 0) s = 0;
        if (e > c->length) e = c->length;      /* the reference's bitset ends at length (beyond is out of bounds there) */
        if (s < e) iv_push(&c->mask, s, e);
    }
    fclose(fp);
    for (size_t k = 0; k < nd; ++k) iv_normalise(&ct[k].mask);

    fp = fopen(argv[3], "r");
    CORN_F_CHK(fp, argv[3]);
    while (fgets(line, sizeof line, fp)) {
        char *t[6];
        if (split_ws(line, t, 6) < 6) continue;
        const int start = atoi(t[3]), end = atoi(t[4]), mlen = atoi(t[5]);
        if (mlen < MIN_TEL) continue;
        contig_t *c = lookup(sorted, nd, t[0]);
        if (!c) continue;
        const int rs = start - 100 < 0 ? 0 : start - 100;
        const int re = end + 100 > c->length ? c->length : end + 100;
        if (rs < re) {                         /* every bit of [rs,re) set <=> one mask run covers it */
            const iv_t *r = iv_find(&c->mask, rs);
            if (!r || r->e < re) continue;
        }
        /* extension (:116-123): left from start while bit(start-1), right from end while bit(end) */
        int xs = start, xe = end;
        if (xs > 0) { const iv_t *r = iv_find(&c->mask, xs - 1); if (r) xs = r->s; }
        if (xe < c->length && xe >= 0) { const iv_t *r = iv_find(&c->mask, xe); if (r) xe = r->e; }
        if (xs < 0) xs = 0;
        if (xe > c->length) xe = c->length;
        if (xs < xe) iv_push(&c->fin, xs, xe);
    }
    fclose(fp);

    for (size_t k = 0; k < nd; ++k) {          /* khash bucket order */
        contig_t *c = &ct[k];
        iv_normalise(&c->fin);
        for (size_t i = 0; i < c->fin.n; ++i) {
            const int s = c->fin.a[i].s, e = c->fin.a[i].e;
            printf("Found telomere positions %d to %d is a telomere in %s of length %d\n", s - 1 < 0 ? 0 : s - 1, e - 1, c->name, c->length);
        }
    }
    for (size_t k = 0; k < nd; ++k) { free(ct[k].mask.a); free(ct[k].fin.a); }
    for (size_t j = 0; j < n; ++j) free(names[j]);
    free(names); free(lengths); free(order); free(ct); free(sorted);
    return EXIT_SUCCESS;
}
