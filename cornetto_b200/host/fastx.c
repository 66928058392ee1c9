/* cornetto_b200/host/fastx.c -- streaming FASTA/FASTQ reader with kseq_read() semantics.
 *
 * Reproduces what the reference's parser delivers to the hot path (src/kseq.h:184-224 with
 * ks_getuntil2, src/kseq.h:91-141), but writes the sequence bytes straight into the pinned
 * batch buffer instead of a realloc'ed kstring:
 *   - a record starts at the next '>' or '@' (anywhere after a FASTQ record, src/kseq.h:189-193);
 *   - name = header bytes up to the first isspace(), the rest of the line is the comment (:195-196);
 *   - sequence = the bytes of the following lines without '\n'; a line whose first byte is
 *     '>', '+' or '@' ends it; empty lines are skipped (:201-205);
 *   - one trailing '\r' per line is dropped, unless it would be the first byte of the sequence
 *     ("str->l > 1", :138);
 *   - '+' introduces a FASTQ quality block, read by length; a missing or mismatching block ends
 *     the input (-2, :214-223).
 */
#include <ctype.h>
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include "cornetto.h"

struct fastx {
    gzFile   fp;          /* gzip input (or stdin of unknown kind) */
    int      fd;          /* plain file: read(2) straight into buf, no zlib copy */
    uint8_t *buf;
    size_t   cap, beg, end;
    int      eof;
    int      last_char;
    char    *name;
    size_t   name_cap;
    int      at_line_start, pending_cr, seq_done, term;
    size_t   line_len;
    uint64_t seq_len;
    /* kseq only learns that the input has ended from a read that returns fewer than its 16384 buffer bytes
     * (src/kseq.h:72-73,107-108): when the input length is a multiple of 16384 it finds out one read later, and two
     * decisions fall the other way at that moment (fastx_next, fastx_seq).  total = bytes delivered so far;
     * zero_read = kseq would already have made the read that returned nothing. */
    uint64_t total;
    int      zero_read;
};

static int kseq_knows_eof(const fastx_t *fx) { return fx->zero_read || (fx->total % 16384u) != 0; }

fastx_t *fastx_open(const char *path)
{
    gzFile fp = NULL;
    int fd = -1;
    if (strcmp(path, "-") == 0) fp = gzdopen(fileno(stdin), "r");
    else {
        fd = open(path, O_RDONLY);
        if (fd < 0) return NULL;
        unsigned char magic[2] = { 0, 0 };
        const ssize_t k = pread(fd, magic, 2, 0);
        if (k == 2 && magic[0] == 0x1f && magic[1] == 0x8b) { fp = gzdopen(fd, "r"); fd = -1; }   /* gzip: through zlib */
    }
    if (!fp && fd < 0) return NULL;
    if (fp) gzbuffer(fp, 1 << 20);
    fastx_t *fx = (fastx_t *)calloc(1, sizeof *fx);
    CORN_MALLOC_CHK(fx);
    fx->fp = fp;
    fx->fd = fd;
    fx->cap = 4u << 20;
    fx->buf = (uint8_t *)malloc(fx->cap);
    fx->name_cap = 256;
    fx->name = (char *)malloc(fx->name_cap);
    CORN_MALLOC_CHK(fx->buf);
    CORN_MALLOC_CHK(fx->name);
    fx->name[0] = 0;
    return fx;
}

fastx_t *fastx_open_at(const char *path, uint64_t offset)
{
    fastx_t *fx = fastx_open(path);
    if (fx && offset) {
        /* plain file: seek; gzip: zlib inflates and discards up to the (decompressed) offset */
        if (fx->fp ? gzseek(fx->fp, (z_off_t)offset, SEEK_SET) < 0 : lseek(fx->fd, (off_t)offset, SEEK_SET) < 0) { fastx_close(fx); return NULL; }
        fx->total = offset;                          /* (kseq_knows_eof: position in the file, not in this reader) */
    }
    return fx;
}

void fastx_close(fastx_t *fx)
{
    if (!fx) return;
    if (fx->fp) gzclose(fx->fp);
    if (fx->fd >= 0) close(fx->fd);
    free(fx->buf); free(fx->name); free(fx);
}

static int fill(fastx_t *fx)
{
    if (fx->beg < fx->end) return 1;
    if (fx->eof) return 0;
    long n;
    if (fx->fp) n = gzread(fx->fp, fx->buf, (unsigned)fx->cap);
    else do { n = (long)read(fx->fd, fx->buf, fx->cap); } while (n < 0 && errno == EINTR);
    fx->beg = 0;
    fx->end = n > 0 ? (size_t)n : 0;
    if (n <= 0) { fx->eof = 1; return 0; }
    fx->total += (uint64_t)n;
    return 1;
}

const char *fastx_name(const fastx_t *fx) { return fx->name; }

int fastx_next(fastx_t *fx)
{
    if (fx->last_char == 0) {                       /* jump to the next header character */
        for (;;) {
            if (!fill(fx)) { fx->zero_read = 1; return 0; }
            const uint8_t *p = fx->buf + fx->beg, *e = fx->buf + fx->end;
            while (p < e && *p != '>' && *p != '@') ++p;
            if (p < e) { fx->beg = (size_t)(p - fx->buf) + 1; break; }
            fx->beg = fx->end;
        }
    }
    if (!fill(fx)) {
        /* the header character was the last byte of the input.  If kseq already knows that (ks_getuntil returns -1,
         * :194) there is no record; if the input length is a multiple of its buffer it reads an empty name instead
         * and delivers one more record: name "", length 0 */
        if (kseq_knows_eof(fx)) return 0;
        fx->zero_read = 1;
        fx->name[0] = 0;
        fx->last_char = 0;
        fx->at_line_start = 1; fx->pending_cr = 0; fx->seq_done = 0; fx->term = 0;
        fx->line_len = 0; fx->seq_len = 0;
        return 1;
    }
    size_t nl = 0;
    int c = 0;
    for (;;) {
        if (!fill(fx)) break;
        int b = fx->buf[fx->beg++];
        if (isspace(b)) { c = b; break; }
        if (nl + 2 > fx->name_cap) { fx->name_cap *= 2; fx->name = (char *)realloc(fx->name, fx->name_cap); CORN_MALLOC_CHK(fx->name); }
        fx->name[nl++] = (char)b;
    }
    fx->name[nl] = 0;
    if (c != '\n') {                                /* comment: rest of the line */
        while (fill(fx)) {
            const uint8_t *p = fx->buf + fx->beg;
            const uint8_t *q = (const uint8_t *)memchr(p, '\n', fx->end - fx->beg);
            if (q) { fx->beg = (size_t)(q - fx->buf) + 1; break; }
            fx->beg = fx->end;
        }
    }
    fx->last_char = 0;
    fx->at_line_start = 1; fx->pending_cr = 0; fx->seq_done = 0; fx->term = 0;
    fx->line_len = 0; fx->seq_len = 0;
    return 1;
}

size_t fastx_seq(fastx_t *fx, uint8_t *dst, size_t cap, int *done)
{
    size_t w = 0;
    *done = 0;
    if (fx->seq_done) { *done = 1; return 0; }
    for (;;) {
        if (!fill(fx)) {                             /* end of input */
            /* a line that is a lone '\r', read right at the end of the input: kseq has stored it and calls
             * ks_getuntil2 for the rest of the line.  Knowing the input has ended, that call returns at once and the
             * byte stays; not knowing it yet (length a multiple of 16384), it runs into the CR rule (:138) and drops
             * the byte unless it is the whole sequence so far */
            if (fx->pending_cr && fx->line_len == 1 && (kseq_knows_eof(fx) || fx->seq_len == 0)) {
                if (w == cap) return w;
                dst[w++] = '\r'; fx->seq_len++;
            }
            fx->zero_read = 1;
            fx->pending_cr = 0;
            fx->term = -1; fx->seq_done = 1; *done = 1;
            return w;
        }
        if (fx->at_line_start) {
            const int c = fx->buf[fx->beg];
            if (c == '>' || c == '@' || c == '+') { fx->beg++; fx->term = c; fx->seq_done = 1; *done = 1; return w; }
            if (c == '\n') { fx->beg++; continue; }
            fx->at_line_start = 0; fx->line_len = 0;
        }
        const uint8_t *p = fx->buf + fx->beg;
        const size_t avail = fx->end - fx->beg;
        const uint8_t *q = (const uint8_t *)memchr(p, '\n', avail);
        const size_t seg = q ? (size_t)(q - p) : avail;
        if (q && !fx->pending_cr) {
            /* common case: the rest of the line, newline included, is in the buffer -- one step.
             * A trailing '\r' goes unless it would be the only byte of the sequence so far (:138). */
            size_t take = seg;
            if (take && p[take - 1] == '\r' && fx->seq_len + seg > 1) --take;
            if (take <= cap - w) {
                memcpy(dst + w, p, take);
                w += take; fx->seq_len += take;
                fx->beg += seg + 1;
                fx->at_line_start = 1;
                continue;
            }
        }
        if (seg == 0) {                              /* the line ends here */
            if (fx->pending_cr) {
                if (fx->seq_len >= 1) fx->pending_cr = 0;            /* stripped */
                else {                                               /* would be the first byte: kept */
                    if (w == cap) return w;
                    dst[w++] = '\r'; fx->seq_len++; fx->pending_cr = 0;
                }
            }
            fx->beg++; fx->at_line_start = 1;
            continue;
        }
        if (fx->pending_cr) {                        /* the held '\r' was not at the end of its line */
            if (w == cap) return w;
            dst[w++] = '\r'; fx->seq_len++; fx->pending_cr = 0;
        }
        size_t take = seg;
        int hold = 0;
        if (p[seg - 1] == '\r') { take = seg - 1; hold = 1; }
        const size_t room = cap - w;
        if (take > room) {
            memcpy(dst + w, p, room);
            fx->beg += room; fx->line_len += room; fx->seq_len += room;
            return w + room;                         /* destination full, record continues */
        }
        memcpy(dst + w, p, take);
        w += take; fx->seq_len += take; fx->line_len += take; fx->beg += take;
        if (hold) { fx->pending_cr = 1; fx->line_len++; fx->beg++; }
    }
}

/* length of the next line (without '\n'); -1 if no byte is left.  last/prev: its final two bytes */
static long line_len(fastx_t *fx, int *last, int *prev)
{
    if (!fill(fx)) return -1;
    long L = 0;
    for (;;) {
        if (!fill(fx)) return L;
        const uint8_t *p = fx->buf + fx->beg;
        const size_t avail = fx->end - fx->beg;
        const uint8_t *q = (const uint8_t *)memchr(p, '\n', avail);
        const size_t seg = q ? (size_t)(q - p) : avail;
        if (seg >= 2) { *prev = p[seg - 2]; *last = p[seg - 1]; }
        else if (seg == 1) { *prev = *last; *last = p[0]; }
        L += (long)seg;
        fx->beg += seg + (q ? 1 : 0);
        if (q) return L;
    }
}

int fastx_finish(fastx_t *fx)
{
    if (fx->term == '>' || fx->term == '@') { fx->last_char = fx->term; return 0; }
    if (fx->term != '+') { fx->last_char = 0; return 0; }        /* FASTA record ended by EOF */
    /* skip the rest of the '+' line (src/kseq.h:218-219) */
    for (;;) {
        if (!fill(fx)) return -2;
        const uint8_t *p = fx->buf + fx->beg;
        const uint8_t *q = (const uint8_t *)memchr(p, '\n', fx->end - fx->beg);
        if (q) { fx->beg = (size_t)(q - fx->buf) + 1; break; }
        fx->beg = fx->end;
    }
    uint64_t ql = 0;
    int last = 0, prev = 0;
    for (;;) {                                                   /* :220 */
        long L = line_len(fx, &last, &prev);
        if (L < 0) break;
        ql += (uint64_t)L;
        if (ql > 1 && last == '\r') { --ql; last = prev; prev = 0; }
        if (ql >= fx->seq_len) break;
    }
    fx->last_char = 0;
    return ql == fx->seq_len ? 0 : -2;
}

/* ------------------------------------------------------------------------------------------------ */
rec_batch_t *rec_batch_create(uint64_t capacity_bytes, uint32_t max_rec)
{
    rec_batch_t *b = (rec_batch_t *)calloc(1, sizeof *b);
    CORN_MALLOC_CHK(b);
    int r = corn_hbatch_create(capacity_bytes, max_rec, &b->hb);
    if (r != CORN_OK) { CORN_ERROR("cannot allocate a %llu byte batch: %s", (unsigned long long)capacity_bytes, corn_gpu_strerror(r)); exit(EXIT_FAILURE); }
    b->name = (char **)calloc(max_rec, sizeof(char *));
    CORN_MALLOC_CHK(b->name);
    b->max_rec = max_rec;
    return b;
}

void rec_batch_destroy(rec_batch_t *b)
{
    if (!b) return;
    for (uint32_t i = 0; i < b->n; ++i) free(b->name[i]);
    free(b->name);
    corn_hbatch_destroy(b->hb);
    free(b);
}

/* partial record carried from a full batch into the next one */
static uint8_t *g_carry;
static size_t   g_carry_len, g_carry_cap;
static char    *g_carry_name;
static int      g_have_carry;

static void grow_batch(rec_batch_t *b, uint64_t need_bytes, const uint8_t *keep, size_t keep_len, const char *name)
{
    uint64_t cap2 = (need_bytes + 4096) * 2;
    if (cap2 > CORN_MAX_BATCH_BYTES) cap2 = CORN_MAX_BATCH_BYTES;
    if (cap2 <= need_bytes + CORN_ALIGN) { CORN_ERROR("record %s is too large for one GPU batch", name); exit(EXIT_FAILURE); }
    corn_hbatch_t *nb = NULL;
    int r = corn_hbatch_create(cap2, b->max_rec, &nb);
    if (r != CORN_OK) { CORN_ERROR("cannot grow the batch to %llu bytes: %s", (unsigned long long)cap2, corn_gpu_strerror(r)); exit(EXIT_FAILURE); }
    if (keep_len) memcpy(corn_hbatch_cursor(nb), keep, keep_len);
    corn_hbatch_destroy(b->hb);
    b->hb = nb;
}

uint32_t rec_batch_fill(rec_batch_t *b, fastx_t *fx)
{
    for (uint32_t i = 0; i < b->n; ++i) free(b->name[i]);
    b->n = 0;
    corn_hbatch_reset(b->hb);
    if (b->eof) return 0;
    for (;;) {
        char *name;
        size_t have = 0;
        if (g_have_carry) {
            name = g_carry_name;
        } else {
            if (!fastx_next(fx)) { b->eof = 1; break; }
            name = strdup(fastx_name(fx));
            CORN_MALLOC_CHK(name);
        }
        int done = 0;
        for (;;) {
            uint64_t room = corn_hbatch_room(b->hb);
            if (g_have_carry) {
                if (room >= g_carry_len) {
                    memcpy(corn_hbatch_cursor(b->hb), g_carry, g_carry_len);
                    have = g_carry_len;
                    g_have_carry = 0;
                } else if (b->n > 0) {
                    return b->n;                      /* flush the finished records first */
                } else {
                    grow_batch(b, g_carry_len, NULL, 0, name);
                    continue;
                }
            }
            if (room > have) {
                have += fastx_seq(fx, corn_hbatch_cursor(b->hb) + have, (size_t)(room - have), &done);
                if (done) break;
                if (have < room) continue;
            }
            /* the batch is full and the record continues */
            if (b->n > 0) {                           /* hand over the finished records, carry this one */
                if (have > g_carry_cap) { g_carry_cap = have * 2 + 4096; g_carry = (uint8_t *)realloc(g_carry, g_carry_cap); CORN_MALLOC_CHK(g_carry); }
                memcpy(g_carry, corn_hbatch_cursor(b->hb), have);
                g_carry_len = have; g_carry_name = name; g_have_carry = 1;
                return b->n;
            }
            /* a single record larger than the whole (empty) batch: grow it, keeping the bytes read so far */
            {
                uint8_t *tmp = (uint8_t *)malloc(have ? have : 1);
                CORN_MALLOC_CHK(tmp);
                memcpy(tmp, corn_hbatch_cursor(b->hb), have);
                grow_batch(b, have, tmp, have, name);
                free(tmp);
            }
        }
        if (fastx_finish(fx) != 0) { free(name); b->eof = 1; break; }   /* kseq_read() returned -2: the caller's loop ends */
        int r = corn_hbatch_commit(b->hb, have);
        if (r != CORN_OK) { CORN_ERROR("batch commit failed: %s", corn_gpu_strerror(r)); exit(EXIT_FAILURE); }
        b->name[b->n++] = name;
        if (b->n == b->max_rec || corn_hbatch_room(b->hb) == 0) break;
    }
    return b->n;
}
