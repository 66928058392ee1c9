/* oracle/oracle.c -- plain-C restatement of the reference algorithms (see oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into the product. */
#define _GNU_SOURCE
#include "oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

/* ========================================================================================
 * FASTA/FASTQ reader -- follows kseq_read(), src/kseq.h:184-224, and ks_getuntil2(),
 * src/kseq.h:91-141, restated over an in-memory copy of the (decompressed) file.
 * ====================================================================================== */

/* zero_read: kseq learns that the input has ended only from a read that returns fewer than its 16384 buffer bytes
 * (src/kseq.h:72-73,107-108); when the input length is a multiple of 16384 it finds out one read later.  Until then
 * ks_getuntil2 at the end of the data does not return -1 (:100) but reads nothing, falls through and finishes the
 * string -- which gives one more record with an empty name after a final '>' / '@', and strips a final lone '\r'.
 * [checked against the compiled reference on files of 16384, 32768 and 49152 bytes] */
typedef struct { const uint8_t *b; size_t n, p; int zero_read; } cur_t;

static int cur_knows_eof(const cur_t *c) { return c->zero_read || (c->n % 16384u) != 0; }

static int cur_getc(cur_t *c)
{
    if (c->p < c->n) return (int)c->b[c->p++];
    c->zero_read = 1;
    return -1;
}

typedef struct { char *s; size_t l, m; } str_t;

static void str_need(str_t *s, size_t extra)
{
    if (s->l + extra + 1 > s->m) {
        s->m = (s->l + extra + 1) * 2;
        s->s = (char *)realloc(s->s, s->m);
    }
}

/* Append bytes up to (not including) the next '\n'; consume the '\n'.  Mirrors the
 * KS_SEP_LINE branch incl. the single trailing '\r' strip "str->l > 1" (src/kseq.h:138).
 * Returns -1 when called at end of data (src/kseq.h:95), else the string length. */
static long getline_append(cur_t *c, str_t *s)
{
    if (c->p >= c->n) {
        if (cur_knows_eof(c)) return -1;
        c->zero_read = 1;                           /* reads nothing, then the CR rule and the terminator as usual */
        str_need(s, 0);
        if (s->l > 1 && s->s[s->l - 1] == '\r') --s->l;
        s->s[s->l] = 0;
        return (long)s->l;
    }
    const uint8_t *nl = (const uint8_t *)memchr(c->b + c->p, '\n', c->n - c->p);
    size_t k = nl ? (size_t)(nl - (c->b + c->p)) : c->n - c->p;
    str_need(s, k);
    memcpy(s->s + s->l, c->b + c->p, k);
    s->l += k;
    c->p += k + (nl ? 1 : 0);
    if (s->l > 1 && s->s[s->l - 1] == '\r') --s->l;
    s->s[s->l] = 0;
    return (long)s->l;
}

int orc_parse_fastx_mem(const uint8_t *buf, size_t n, orc_rec_t **recs_out, size_t *n_out)
{
    cur_t c = { buf, n, 0, 0 };
    orc_rec_t *recs = NULL;
    size_t nr = 0, mr = 0;
    int last_char = 0, ch;

    for (;;) {
        str_t name = {0}, seq = {0}, qual = {0}, comment = {0};
        if (last_char == 0) {                       /* src/kseq.h:189-193 */
            while ((ch = cur_getc(&c)) != -1 && ch != '>' && ch != '@') {}
            if (ch == -1) break;
            last_char = ch;
        }
        /* name: up to the first isspace() (src/kseq.h:195) */
        if (c.p >= c.n) {                           /* ks_getuntil() < 0 -> EOF, if kseq knows; else an empty name */
            if (cur_knows_eof(&c)) break;
            c.zero_read = 1;
        }
        str_need(&name, 0);
        ch = 0;
        while (c.p < c.n) {
            int b = c.b[c.p++];
            if (isspace(b)) { ch = b; break; }
            str_need(&name, 1);
            name.s[name.l++] = (char)b;
        }
        name.s[name.l] = 0;
        if (ch != '\n') getline_append(&c, &comment); /* src/kseq.h:196 */
        str_need(&seq, 0);
        seq.s[0] = 0;
        /* sequence lines (src/kseq.h:201-205) */
        while ((ch = cur_getc(&c)) != -1 && ch != '>' && ch != '+' && ch != '@') {
            if (ch == '\n') continue;
            str_need(&seq, 1);
            seq.s[seq.l++] = (char)ch;
            seq.s[seq.l] = 0;
            getline_append(&c, &seq);
        }
        if (ch == '>' || ch == '@') last_char = ch;
        int ok = 1;
        if (ch == '+') {                             /* FASTQ (src/kseq.h:214-223) */
            while ((ch = cur_getc(&c)) != -1 && ch != '\n') {}
            if (ch == -1) ok = 0;                    /* -2: no quality string */
            else {
                str_need(&qual, 0);
                while (getline_append(&c, &qual) >= 0 && qual.l < seq.l) {}
                last_char = 0;
                if (seq.l != qual.l) ok = 0;         /* -2: truncated quality */
            }
        }
        free(qual.s); free(comment.s);
        if (!ok) { free(name.s); free(seq.s); break; }
        if (nr == mr) { mr = mr ? mr * 2 : 16; recs = (orc_rec_t *)realloc(recs, mr * sizeof(*recs)); }
        recs[nr].name = name.s; recs[nr].seq = seq.s; recs[nr].len = seq.l;
        ++nr;
    }
    *recs_out = recs; *n_out = nr;
    return 0;
}

static uint8_t *slurp(const char *path, size_t *n_out)
{
    gzFile fp = strcmp(path, "-") ? gzopen(path, "r") : gzdopen(fileno(stdin), "r");
    if (!fp) return NULL;
    gzbuffer(fp, 1 << 20);
    size_t n = 0, m = 1 << 22;
    uint8_t *b = (uint8_t *)malloc(m);
    for (;;) {
        if (m - n < (1 << 20)) { m *= 2; b = (uint8_t *)realloc(b, m); }
        int k = gzread(fp, b + n, (unsigned)(m - n > (1u << 30) ? (1u << 30) : m - n));
        if (k <= 0) break;
        n += (size_t)k;
    }
    gzclose(fp);
    *n_out = n;
    return b;
}

int orc_read_fastx(const char *path, orc_rec_t **recs, size_t *n_recs)
{
    size_t n;
    uint8_t *b = slurp(path, &n);
    if (!b) return -1;
    int r = orc_parse_fastx_mem(b, n, recs, n_recs);
    free(b);
    return r;
}

void orc_free_recs(orc_rec_t *recs, size_t n)
{
    for (size_t i = 0; i < n; ++i) { free(recs[i].name); free(recs[i].seq); }
    free(recs);
}

/* ========================================================================================
 * telofind -- src/find_telomere.c
 * ====================================================================================== */

void orc_revcomp_motif(const char *motif, char *out)   /* rc(): src/find_telomere.c:24-42 */
{
    size_t m = strlen(motif);
    for (size_t i = 0; i < m; ++i) {
        char c = motif[m - 1 - i];
        out[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
    }
    out[m] = 0;
}

/* One strand pass of find(): src/find_telomere.c:49-58 (and :63-72). */
static void scan_strand(const char *q, size_t n, const char *pat, uint32_t strand,
                        orc_run_t **runs, size_t *nr, size_t *mr)
{
    size_t m = strlen(pat), pos = 0;
    if (m == 0) return;              /* reference would loop forever on an empty motif */
    while (pos <= n) {
        const char *hit = strstr(q + pos, pat);
        if (!hit) break;
        size_t p = (size_t)(hit - q), e = p;
        while (strncmp(q + e, pat, m) == 0) e += m;
        if (*nr == *mr) { *mr = *mr ? *mr * 2 : 64; *runs = (orc_run_t *)realloc(*runs, *mr * sizeof(**runs)); }
        (*runs)[*nr].strand = strand; (*runs)[*nr].start = p; (*runs)[*nr].end = e; ++*nr;
        pos = e + 1;                 /* may step one past the NUL exactly as the reference does; guarded by pos <= n */
    }
}

size_t orc_telofind(const char *seq, size_t len, const char *motif, orc_run_t **runs)
{
    char *q = (char *)malloc(len + 2), *rev = (char *)malloc(strlen(motif) + 1);
    for (size_t i = 0; i < len; ++i) q[i] = (char)toupper((unsigned char)seq[i]);  /* disambiguate(): :76-81 */
    q[len] = q[len + 1] = 0;
    size_t nr = 0, mr = 0;
    *runs = NULL;
    scan_strand(q, len, motif, 0, runs, &nr, &mr);
    orc_revcomp_motif(motif, rev);
    scan_strand(q, len, rev, 1, runs, &nr, &mr);
    free(q); free(rev);
    return nr;
}

int orc_telofind_file(const char *path, const char *motif, FILE *out)
{
    orc_rec_t *recs; size_t n;
    if (orc_read_fastx(path, &recs, &n) < 0) return -1;
    for (size_t i = 0; i < n; ++i) {
        orc_run_t *r; size_t k = orc_telofind(recs[i].seq, recs[i].len, motif, &r);
        for (size_t j = 0; j < k; ++j)      /* src/find_telomere.c:51,56 */
            fprintf(out, "%s\t%zu\t%u\t%zu\t%zu\t%zu\n", recs[i].name, recs[i].len, r[j].strand,
                    (size_t)r[j].start, (size_t)r[j].end, (size_t)(r[j].end - r[j].start));
        free(r);
    }
    orc_free_recs(recs, n);
    return 0;
}

/* ========================================================================================
 * telowin -- src/telomere_windows.c
 * ====================================================================================== */

double orc_telowin_threshold(double thr, double identity_percent)
{
    double identity = identity_percent / 100;        /* :53 */
    return thr * pow(identity, 6);                   /* :54 */
}

size_t orc_telowin_contig(const uint8_t *b, int length, double thr, orc_win_t **wins)
{
    size_t nw = 0, mw = 0;
    *wins = NULL;
    if (!b) return 0;
    for (int i = 0; i <= length; i += 200) {          /* :31, WINDOW_SIZE/5 */
        int car = 0;
        for (int j = i; j < i + 1000 && j < length; ++j) if (b[j]) ++car;
        int den = (i + 1000 < length) ? 1000 : length - i;
        if ((double)car / den >= thr) {
            if (nw == mw) { mw = mw ? mw * 2 : 64; *wins = (orc_win_t *)realloc(*wins, mw * sizeof(**wins)); }
            (*wins)[nw].start = i; (*wins)[nw].end = i + den; (*wins)[nw].car = car; ++nw;
        }
        if (i + 1000 >= length) break;
    }
    return nw;
}

static void flush_contig(const char *name, uint8_t *b, int length, double thr, FILE *out)
{
    orc_win_t *w; size_t n = orc_telowin_contig(b, length, thr, &w);
    for (size_t i = 0; i < n; ++i) {
        int den = w[i].end - w[i].start;
        fprintf(out, "Window\t%s\t%d\t%d\t%d\t%.3g\n", name, length, w[i].start, w[i].end, (double)w[i].car / den);
    }
    free(w);
}

/* whitespace tokenizer equivalent to sscanf("%s %s ...") on one line */
static int split_ws(char *line, char **tok, int maxtok)
{
    int n = 0; char *p = line;
    while (n < maxtok) {
        while (*p && isspace((unsigned char)*p)) ++p;
        if (!*p) break;
        tok[n++] = p;
        while (*p && !isspace((unsigned char)*p)) ++p;
        if (*p) *p++ = 0;
    }
    return n;
}

int orc_telowin_file(const char *path, double identity_percent, int have_thr, double thr_in, FILE *out)
{
    double thr = orc_telowin_threshold(have_thr ? thr_in : 0.4, identity_percent);
    FILE *fp = fopen(path, "r");
    if (!fp) return -1;
    char line[2048], name[2048] = {0};
    uint8_t *b = NULL; int length = 0;
    while (fgets(line, sizeof line, fp)) {
        char *t[6];
        if (split_ws(line, t, 6) < 6) continue;      /* malformed lines are undefined in the reference */
        if (!b || strcmp(t[0], name) != 0) {         /* :69-74 */
            flush_contig(name, b, length, thr, out);
            free(b);
            length = atoi(t[1]);
            b = (uint8_t *)calloc(length > 0 ? (size_t)length : 1, 1);
            strcpy(name, t[0]);
        }
        int s = atoi(t[3]), e = atoi(t[4]);
        for (int i = s; i < e; ++i) b[i] = 1;        /* :75-79 */
    }
    flush_contig(name, b, length, thr, out);
    free(b);
    fclose(fp);
    return 0;
}

/* ========================================================================================
 * sdust -- src/sdust/sdust.c (lh3/sdust 0.1-r2)
 * ====================================================================================== */

typedef struct { int start, finish, r, l; } pint_t;      /* perf_intv_t, :13-16 */

typedef struct {
    int *w; int wcap, whead, wn;                          /* the triplet deque (kdq) */
    pint_t *P; int nP, mP;                                /* descending start */
    uint64_t *res; int nres, mres;
    int cw[64], cv[64], rw, rv, L;
} sd_t;

static int  dq_at(const sd_t *s, int i) { return s->w[(s->whead + i) % s->wcap]; }
static void dq_push(sd_t *s, int t)
{
    if (s->wn == s->wcap) {                               /* grow, keeping logical order */
        int ncap = s->wcap * 2, *nw = (int *)malloc(sizeof(int) * ncap);
        for (int i = 0; i < s->wn; ++i) nw[i] = dq_at(s, i);
        free(s->w); s->w = nw; s->wcap = ncap; s->whead = 0;
    }
    s->w[(s->whead + s->wn) % s->wcap] = t; ++s->wn;
}
static int  dq_shift(sd_t *s) { int t = s->w[s->whead]; s->whead = (s->whead + 1) % s->wcap; --s->wn; return t; }

/* shift_window(): :66-86 */
static void sd_shift_window(sd_t *s, int t, int T, int W)
{
    if (s->wn >= W - 3 + 1) {
        int x = dq_shift(s);
        s->rw -= --s->cw[x];
        if (s->L > s->wn) { --s->L; s->rv -= --s->cv[x]; }
    }
    dq_push(s, t);
    ++s->L;
    s->rw += s->cw[t]++;
    s->rv += s->cv[t]++;
    if (s->cv[t] * 10 > T << 1) {
        int x;
        do {
            x = dq_at(s, s->wn - s->L);
            s->rv -= --s->cv[x];
            --s->L;
        } while (x != t);
    }
}

/* save_masked_regions(): :88-102 */
static void sd_save(sd_t *s, int start)
{
    if (s->nP == 0 || s->P[s->nP - 1].start >= start) return;
    pint_t *p = &s->P[s->nP - 1];
    int saved = 0;
    if (s->nres) {
        int a = (int)(s->res[s->nres - 1] >> 32), f = (int)(uint32_t)s->res[s->nres - 1];
        if (p->start <= f) {
            saved = 1;
            s->res[s->nres - 1] = (uint64_t)a << 32 | (uint32_t)(f > p->finish ? f : p->finish);
        }
    }
    if (!saved) {
        if (s->nres == s->mres) { s->mres = s->mres ? s->mres * 2 : 64; s->res = (uint64_t *)realloc(s->res, 8 * (size_t)s->mres); }
        s->res[s->nres++] = (uint64_t)p->start << 32 | (uint32_t)p->finish;
    }
    int i;
    for (i = s->nP - 1; i >= 0 && s->P[i].start < start; --i) {}
    s->nP = i + 1;
}

/* find_perfect(): :104-128 */
static void sd_find_perfect(sd_t *s, int T, int start)
{
    int c[64], r = s->rv, max_r = 0, max_l = 0;
    memcpy(c, s->cv, sizeof c);
    for (int i = s->wn - s->L - 1; i >= 0; --i) {
        int t = dq_at(s, i), j;
        r += c[t]++;
        int new_r = r, new_l = s->wn - i - 1;
        if (new_r * 10 > T * new_l) {
            for (j = 0; j < s->nP && s->P[j].start >= i + start; ++j) {
                pint_t *p = &s->P[j];
                if (max_r == 0 || p->r * max_l > max_r * p->l) { max_r = p->r; max_l = p->l; }
            }
            if (max_r == 0 || new_r * max_l >= max_r * new_l) {
                max_r = new_r; max_l = new_l;
                if (s->nP == s->mP) { s->mP = s->mP ? s->mP * 2 : 16; s->P = (pint_t *)realloc(s->P, sizeof(pint_t) * s->mP); }
                memmove(&s->P[j + 1], &s->P[j], (size_t)(s->nP - j) * sizeof(pint_t));
                ++s->nP;
                s->P[j].start = i + start; s->P[j].finish = s->wn + 2 + start;
                s->P[j].r = new_r; s->P[j].l = new_l;
            }
        }
    }
}

static int nt4(uint8_t c)                                   /* seq_nt4_table: :23-40 */
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    case 0: return 0; case 1: return 1; case 2: return 2; case 3: return 3;  /* table rows 0..3 */
    default: return 4;
    }
}

/* sdust_core()+sdust(): :130-171 */
uint64_t *orc_sdust(const uint8_t *seq, int l_seq, int T, int W, int *n)
{
    sd_t s; memset(&s, 0, sizeof s);
    s.wcap = 8; s.w = (int *)malloc(sizeof(int) * s.wcap);
    if (l_seq < 0) l_seq = (int)strlen((const char *)seq);
    int l = 0, start; unsigned t = 0;
    for (int i = 0; i <= l_seq; ++i) {
        int b = i < l_seq ? nt4(seq[i]) : 4;
        if (b < 4) {
            ++l; t = (t << 2 | (unsigned)b) & 63;
            if (l >= 3) {
                start = (l - W > 0 ? l - W : 0) + (i + 1 - l);
                sd_save(&s, start);
                sd_shift_window(&s, (int)t, T, W);
                if (s.rw * 10 > s.L * T) sd_find_perfect(&s, T, start);
            }
        } else {
            start = (l - W + 1 > 0 ? l - W + 1 : 0) + (i + 1 - l);
            while (s.nP) sd_save(&s, start++);
            l = 0; t = 0;
        }
    }
    free(s.w); free(s.P);
    *n = s.nres;
    return s.res;
}

int orc_sdust_file(const char *path, int T, int W, FILE *out)
{
    orc_rec_t *recs; size_t n;
    if (orc_read_fastx(path, &recs, &n) < 0) return -1;
    for (size_t i = 0; i < n; ++i) {
        int k; uint64_t *r = orc_sdust((const uint8_t *)recs[i].seq, -1, T, W, &k);
        for (int j = 0; j < k; ++j)          /* :200-201 */
            fprintf(out, "%s\t%d\t%d\n", recs[i].name, (int)(r[j] >> 32), (int)r[j]);
        free(r);
    }
    orc_free_recs(recs, n);
    return 0;
}

/* ========================================================================================
 * khash iteration order -- src/khash.h:230-348 (resize, put), :395-400 (X31 hash)
 * ====================================================================================== */

static uint32_t x31(const char *s)
{
    uint32_t h = (uint32_t)(int)*s;                       /* char is signed on the reference's targets */
    if (h) for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)(int)*s;
    return h;
}

typedef struct { uint32_t nb, size, nocc, upper; long *slot; /* -1 empty, else name index */ } kh_t;

static void kh_grow(kh_t *h, uint32_t want, const char *const *names)
{
    uint32_t nn = want;
    --nn; nn |= nn >> 1; nn |= nn >> 2; nn |= nn >> 4; nn |= nn >> 8; nn |= nn >> 16; ++nn;   /* kroundup32 */
    if (nn < 4) nn = 4;
    if (h->size >= (uint32_t)(nn * 0.77 + 0.5)) return;   /* requested size too small: no change */
    /* Re-insert in old-bucket order.  The reference does this in place with a kick-out loop
     * (:268-294); because a kicked-out element is immediately re-inserted before the scan
     * continues, the sequence of insertions into the new table is: for each old bucket j in
     * ascending order that is still un-moved, insert its key, then the key it displaced from the
     * old slot it landed on (if that slot held an un-moved key), and so on. */
    long *ns = (long *)malloc(sizeof(long) * nn);
    for (uint32_t i = 0; i < nn; ++i) ns[i] = -1;
    uint8_t *moved = (uint8_t *)calloc(h->nb ? h->nb : 1, 1);
    uint32_t mask = nn - 1;
    for (uint32_t j = 0; j < h->nb; ++j) {
        if (h->slot[j] < 0 || moved[j]) continue;
        long key = h->slot[j];
        moved[j] = 1;
        for (;;) {
            uint32_t i = x31(names[key]) & mask, step = 0;
            while (ns[i] >= 0) i = (i + (++step)) & mask;
            ns[i] = key;
            if (i < h->nb && h->slot[i] >= 0 && !moved[i]) { key = h->slot[i]; moved[i] = 1; }
            else break;
        }
    }
    free(moved); free(h->slot);
    h->slot = ns; h->nb = nn; h->nocc = h->size; h->upper = (uint32_t)(nn * 0.77 + 0.5);
}

size_t orc_khash_order(const char *const *names, size_t n, size_t *order)
{
    kh_t h; memset(&h, 0, sizeof h);
    for (size_t k = 0; k < n; ++k) {
        if (h.nocc >= h.upper) {                           /* kh_put: :301-311 */
            if (h.nb > (h.size << 1)) kh_grow(&h, h.nb - 1, names);
            else kh_grow(&h, h.nb + 1, names);
        }
        uint32_t mask = h.nb - 1, i = x31(names[k]) & mask, step = 0, last = i;
        int present = 0;
        while (h.slot[i] >= 0) {
            if (strcmp(names[h.slot[i]], names[k]) == 0) { present = 1; break; }
            i = (i + (++step)) & mask;
            if (i == last) break;
        }
        if (!present && h.slot[i] < 0) { h.slot[i] = (long)k; ++h.size; ++h.nocc; }
    }
    size_t m = 0;
    for (uint32_t i = 0; i < h.nb; ++i) if (h.slot[i] >= 0) order[m++] = (size_t)h.slot[i];
    free(h.slot);
    return m;
}

/* ========================================================================================
 * telobreaks -- src/telomere_breaks.c:47-172 (literal bitset form)
 * ====================================================================================== */

typedef struct { char *name; int length; uint8_t *bits, *fin; } scaf_t;

static void bset(uint8_t *b, int i) { b[i / 8] |= (uint8_t)(1 << (i % 8)); }   /* :33-35 */
static int  bget(const uint8_t *b, int i) { return b[i / 8] & (1 << (i % 8)); } /* :37-39 */

int orc_telobreaks_files(const char *lens, const char *sdust, const char *telomere, FILE *out)
{
    FILE *fp = fopen(lens, "r");
    if (!fp) return -1;
    char line[2048];
    char **names = NULL; int *lengths = NULL; size_t n = 0, m = 0;
    while (fgets(line, sizeof line, fp)) {                 /* :61-72 */
        char *t[2];
        if (split_ws(line, t, 2) < 2) continue;
        if (n == m) { m = m ? m * 2 : 64; names = (char **)realloc(names, m * sizeof(char *)); lengths = (int *)realloc(lengths, m * sizeof(int)); }
        names[n] = strdup(t[0]); lengths[n] = atoi(t[1]); ++n;
    }
    fclose(fp);
    size_t *order = (size_t *)malloc(sizeof(size_t) * (n ? n : 1));
    size_t nd = orc_khash_order((const char *const *)names, n, order);
    /* one scaffold per distinct key; a repeated name replaces the value (:66-71) so the LAST length wins */
    scaf_t *sc = (scaf_t *)calloc(nd ? nd : 1, sizeof(scaf_t));
    for (size_t k = 0; k < nd; ++k) {
        sc[k].name = names[order[k]];
        for (size_t j = 0; j < n; ++j) if (strcmp(names[j], sc[k].name) == 0) sc[k].length = lengths[j];
        size_t nb = (size_t)ceil(sc[k].length / 8.0);
        sc[k].bits = (uint8_t *)calloc(nb ? nb : 1, 1);
        sc[k].fin  = (uint8_t *)calloc(nb ? nb : 1, 1);
    }
#define FIND(nm, idx) do { idx = -1; for (size_t q_ = 0; q_ < nd; ++q_) if (strcmp(sc[q_].name, nm) == 0) { idx = (long)q_; break; } } while (0)
    fp = fopen(sdust, "r");
    if (!fp) return -1;
    while (fgets(line, sizeof line, fp)) {                 /* :79-89 */
        char *t[3]; long k;
        if (split_ws(line, t, 3) < 3) continue;
        FIND(t[0], k);
        if (k < 0) continue;
        int s = atoi(t[1]), e = atoi(t[2]);
        /* the reference writes bits >= length out of bounds (sdust intervals may pass the record end after
         * an N); it never READS a bit >= length (:104-123,:136-138), so dropping them is equivalent */
        if (s < 0) s = 0;
        if (e > sc[k].length) e = sc[k].length;
        for (int j = s; j < e; ++j) bset(sc[k].bits, j);
    }
    fclose(fp);
    fp = fopen(telomere, "r");
    if (!fp) return -1;
    while (fgets(line, sizeof line, fp)) {                 /* :96-129 */
        char *t[6]; long k;
        if (split_ws(line, t, 6) < 6) continue;
        int start = atoi(t[3]), end = atoi(t[4]), mlen = atoi(t[5]);
        if (mlen < 24) continue;                           /* MIN_TEL :10 */
        FIND(t[0], k);
        if (k < 0) continue;
        int rs = start - 100 < 0 ? 0 : start - 100;
        int re = end + 100 > sc[k].length ? sc[k].length : end + 100;
        int all = 1;
        for (int j = rs; j < re; ++j) if (!bget(sc[k].bits, j)) { all = 0; break; }
        if (!all) continue;
        rs = start; while (rs > 0 && bget(sc[k].bits, rs - 1)) --rs;
        re = end;   while (re < sc[k].length && bget(sc[k].bits, re)) ++re;
        for (int j = rs; j < re; ++j) bset(sc[k].fin, j);
    }
    fclose(fp);
    for (size_t k = 0; k < nd; ++k) {                      /* :133-148, bucket order */
        for (int i = 0; i < sc[k].length; ++i) {
            if (!bget(sc[k].fin, i)) continue;
            int end = i;
            while (end < sc[k].length && bget(sc[k].fin, end)) ++end;
            i = i - 1 < 0 ? 0 : i - 1;
            fprintf(out, "Found telomere positions %d to %d is a telomere in %s of length %d\n", i, end - 1, sc[k].name, sc[k].length);
            i = end;
        }
    }
    for (size_t k = 0; k < nd; ++k) { free(sc[k].bits); free(sc[k].fin); }
    for (size_t j = 0; j < n; ++j) free(names[j]);
    free(sc); free(names); free(lengths); free(order);
    return 0;
}

/* ========================================================================================
 * fa2bed -- src/assbed.c:97-100
 * ====================================================================================== */
int orc_fa2bed_file(const char *path, FILE *out)
{
    orc_rec_t *recs; size_t n;
    if (orc_read_fastx(path, &recs, &n) < 0) return -1;
    for (size_t i = 0; i < n; ++i) fprintf(out, "%s\t%d\t%d\n", recs[i].name, 0, (int)recs[i].len);
    orc_free_recs(recs, n);
    return 0;
}


/* ========================================================================================
 * noboringbits / boringbits -- follows get_depths() src/boringbits_main.c:179-293, get_regs() :315-372,
 * print_fun_bits() :425-446, print_boring_bits() :465-485, the_boring_bits() :487-539.
 * ====================================================================================== */
void orc_bits_defaults(orc_bits_opt_t *o)
{   /* init_optp(), :543-559 */
    o->window_size = 2500; o->window_inc = 50;
    o->low_cov_thresh = 0.4f; o->high_cov_thresh = 2.5f; o->low_mq_cov_thresh = 0.4f;
    o->min_ctg_len = 1000000; o->edge_len = 100000; o->boring = 1;
}

typedef struct { char *name; int len, cap; uint16_t *d, *q; } bits_ctg_t;

int orc_bits_files(const char *cov_total, const char *cov_mq, const orc_bits_opt_t *o, FILE *out, FILE *err)
{
    FILE *f1 = fopen(cov_total, "r"), *f2 = fopen(cov_mq, "r");
    if (!f1 || !f2) { if (err) fprintf(err, "cannot open input\n"); if (f1) fclose(f1); if (f2) fclose(f2); return 1; }
    bits_ctg_t *c = NULL;
    int n = 0, m = 0, rc = 0;
    char b1[10000], b2[10000], prev[10000] = "";
    int st1, st2, e1, e2, d1, d2, ret, prev_pos = 0;
    double tot_len = 0, tot_d = 0, tot_q = 0;
    for (;;) {                                                   /* :201-279 */
        if ((ret = fscanf(f1, "%s\t%d\t%d\t%d\n", b1, &st1, &e1, &d1)) == EOF) break;
        if (ret != 4) { if (err) fprintf(err, "The depth files should have 4 columns. Had %d.\n", ret); rc = 1; break; }
        if ((ret = fscanf(f2, "%s\t%d\t%d\t%d\n", b2, &st2, &e2, &d2)) == EOF) { if (err) fprintf(err, "The two files are not in the same order\n"); rc = 1; break; }
        if (ret != 4) { if (err) fprintf(err, "The depth files should have 4 columns. Had %d.\n", ret); rc = 1; break; }
        if (strcmp(b1, b2) != 0 || st1 != st2 || e1 != e2) { if (err) fprintf(err, "The two files are not in the same order\n"); rc = 1; break; }
        if (strcmp(b1, prev) != 0) {
            strcpy(prev, b1);
            if (n == m) { m = m ? m * 2 : 4; c = (bits_ctg_t *)realloc(c, (size_t)m * sizeof *c); }
            c[n].name = strdup(b1); c[n].len = 0; c[n].cap = 100;
            c[n].d = (uint16_t *)calloc(100, 2); c[n].q = (uint16_t *)calloc(100, 2);
            ++n;
            prev_pos = 0;                                        /* (the first line of a contig is not checked against 0) */
        } else {
            if (prev_pos + 1 != st1) { if (err) fprintf(err, "The depth files should be incremantal at one base resolution. Found %d to %d\n", prev_pos, st1); rc = 1; break; }
            ++prev_pos;
        }
        if (st1 + 1 != e1) { if (err) fprintf(err, "The depth files should have end=start+1. Found %d to %d\n", st1, e1); rc = 1; break; }
        if (d1 > 65535) d1 = 65535;                              /* (+ a WARNING on stderr) */
        if (d2 > 65535) d2 = 65535;
        bits_ctg_t *x = &c[n - 1];
        if (x->len == x->cap) { x->cap *= 2; x->d = (uint16_t *)realloc(x->d, (size_t)x->cap * 2); x->q = (uint16_t *)realloc(x->q, (size_t)x->cap * 2); }
        x->d[x->len] = (uint16_t)d1; x->q[x->len] = (uint16_t)d2; ++x->len;   /* (a negative depth wraps, as in the reference's store) */
        tot_d += d1; tot_q += d2; tot_len++;
    }
    fclose(f1); fclose(f2);
    if (!rc) {
        const int mean_depth = (int)round(tot_d / tot_len);
        const int w = o->window_size, inc = o->window_inc;
        const int lo = (int)round(o->low_cov_thresh * mean_depth), hi = (int)round(o->high_cov_thresh * mean_depth);    /* :524-525 */
        for (int i = 0; i < n; ++i) {
            const bits_ctg_t *x = &c[i];
            const int length = x->len;
            int n_reg = (length - w + inc - 1) / inc + 1;        /* :331-332 */
            if (n_reg < 1) n_reg = 1;
            if (!o->boring) {                                    /* print_fun_bits, :425-446 */
                if (length < o->min_ctg_len) { fprintf(out, "%s\t%d\t%d\t.\t.\n", x->name, 0, o->min_ctg_len); continue; }
                fprintf(out, "%s\t%d\t%d\t.\t.\n", x->name, 0, o->edge_len);
                fprintf(out, "%s\t%d\t%d\t.\t.\n", x->name, length - o->edge_len, length);
            } else if (!(length > o->min_ctg_len)) continue;     /* print_boring_bits, :468 */
            for (int j = 0; j < n_reg; ++j) {                    /* get_regs, :340-363 */
                const int st = j * inc;
                int end = st + w;
                if (end > length) end = length;
                int depth = 0, mq = 0;
                for (int k = st; k < end; ++k) { depth += x->d[k]; mq += x->q[k]; }
                depth /= (end - st); mq /= (end - st);
                const int fun = depth < lo || depth > hi || (mq / (double)depth) < o->low_mq_cov_thresh;
                if (!o->boring) { if (fun) fprintf(out, "%s\t%d\t%d\t%d\t%d\n", x->name, st, end, depth, mq); }
                else if (st > o->edge_len && end < length - o->edge_len && !fun) fprintf(out, "%s\t%d\t%d\t%d\t%d\n", x->name, st, end, depth, mq);
            }
        }
    }
    for (int i = 0; i < n; ++i) { free(c[i].name); free(c[i].d); free(c[i].q); }
    free(c);
    return rc;
}
