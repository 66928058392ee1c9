/* cornetto_b200/host/seq_main.c -- `cornetto seq [-m INT] <reads.fastq>`: length filter for reads.
 *
 * A by-product of the reader, host only: re-prints every record of at least
 * min-len bases as  @name[\tcomment] / sequence / + / quality  and reports the totals on stderr.
 * Contract restated from seq_main(), src/seq.c:52-138.  Unlike the scan commands this one needs the
 * comment and the quality string of a record, which the batch reader (fastx.c) never materialises, so
 * it carries its own small record reader with the full kseq_read() semantics (src/kseq.h:184-224,
 * ks_getuntil2 :102-141), including what the reference's buffers do between records: the quality
 * buffer is only (re)filled by records that have a '+' line, so a FASTA record prints whatever it
 * held before -- "(null)" if nothing ever did. */
#include <ctype.h>
#include <getopt.h>
#include <zlib.h>

#include "cornetto.h"

typedef struct { char *s; size_t l, m; } str_t;

typedef struct {
    gzFile fp;
    unsigned char *buf;
    int beg, end, eof;
    uint64_t total;          /* bytes delivered so far; zero_read: kseq would already have made the read that returned */
    int zero_read;           /* nothing -- it learns of the end of the input only from a short read (src/kseq.h:107-108) */
    int last_char;
    str_t name, comment, seq, qual;
} rd_t;

enum { RD_BUF = 1 << 16, SEP_SPACE = 0, SEP_LINE = 2 };

static void str_reserve(str_t *s, size_t need)
{
    if (need <= s->m) return;
    size_t m = s->m ? s->m : 64;
    while (m < need) m *= 2;
    s->s = (char *)realloc(s->s, m);
    CORN_MALLOC_CHK(s->s);
    s->m = m;
}

static int rd_fill(rd_t *r)
{
    if (r->beg < r->end) return 1;
    if (r->eof) return 0;
    r->beg = 0;
    r->end = gzread(r->fp, r->buf, RD_BUF);
    if (r->end <= 0) { r->end = 0; r->eof = 1; return 0; }
    r->total += (uint64_t)r->end;
    return 1;
}

static int rd_knows_eof(const rd_t *r) { return r->zero_read || (r->total % 16384u) != 0; }

static int rd_getc(rd_t *r)
{
    if (rd_fill(r)) return r->buf[r->beg++];
    r->zero_read = 1;
    return -1;
}

/* reads up to (and consumes) the next delimiter; returns the string length or -1 when no byte was left */
static long rd_until(rd_t *r, int sep, str_t *s, int *dret, int append)
{
    int got = 0;
    if (dret) *dret = 0;
    if (!append) s->l = 0;
    while (rd_fill(r)) {
        int i = r->beg;
        if (sep == SEP_LINE) { unsigned char *q = (unsigned char *)memchr(r->buf + i, '\n', (size_t)(r->end - i)); i = q ? (int)(q - r->buf) : r->end; }
        else while (i < r->end && !isspace(r->buf[i])) ++i;
        str_reserve(s, s->l + (size_t)(i - r->beg) + 1);
        memcpy(s->s + s->l, r->buf + r->beg, (size_t)(i - r->beg));
        s->l += (size_t)(i - r->beg);
        got = 1;
        r->beg = i + 1;
        if (i < r->end) { if (dret) *dret = r->buf[i]; break; }
    }
    if (!got) {
        /* no byte left.  kseq returns -1 only if it already KNOWS the input has ended (:100); when the length is a
         * multiple of its 16384-byte buffer it reads nothing, falls through and finishes the string (empty name;
         * CR rule applied to what is there) */
        if (rd_knows_eof(r)) return -1;
        r->zero_read = 1;
    }
    str_reserve(s, s->l + 1);
    if (sep == SEP_LINE && s->l > 1 && s->s[s->l - 1] == '\r') --s->l;      /* (on the whole string, :138) */
    s->s[s->l] = 0;
    return (long)s->l;
}

/* >= 0: sequence length; -1: end of input; -2: truncated or mismatching quality */
static long rd_next(rd_t *r)
{
    int c;
    if (r->last_char == 0) {
        while ((c = rd_getc(r)) != -1 && c != '>' && c != '@') { }
        if (c == -1) return -1;
        r->last_char = c;
    }
    r->comment.l = r->seq.l = r->qual.l = 0;
    if (rd_until(r, SEP_SPACE, &r->name, &c, 0) < 0) return -1;
    if (c != '\n') rd_until(r, SEP_LINE, &r->comment, NULL, 0);
    str_reserve(&r->seq, 256);
    while ((c = rd_getc(r)) != -1 && c != '>' && c != '+' && c != '@') {
        if (c == '\n') continue;
        str_reserve(&r->seq, r->seq.l + 2);
        r->seq.s[r->seq.l++] = (char)c;
        rd_until(r, SEP_LINE, &r->seq, NULL, 1);
    }
    if (c == '>' || c == '@') r->last_char = c;
    str_reserve(&r->seq, r->seq.l + 1);
    r->seq.s[r->seq.l] = 0;
    if (c != '+') return (long)r->seq.l;
    str_reserve(&r->qual, r->seq.m);
    while ((c = rd_getc(r)) != -1 && c != '\n') { }
    if (c == -1) return -2;
    while (rd_until(r, SEP_LINE, &r->qual, NULL, 1) >= 0 && r->qual.l < r->seq.l) { }
    r->last_char = 0;
    return r->seq.l == r->qual.l ? (long)r->seq.l : -2;
}

static const struct option seq_options[] = {
    { "verbose", required_argument, 0, 'v' },
    { "min-len", required_argument, 0, 'm' },
    { "help", no_argument, 0, 'h' },
    { 0, 0, 0, 0 } };

static void seq_usage(FILE *fp)
{
    fprintf(fp, "Usage: cornetto seq <reads.fastq> \n");
    fprintf(fp, "   -m INT                     min length [%d]\n", 30000);
    fprintf(fp, "   -h                         help\n");
}

int seq_main(int argc, char *argv[])
{
    FILE *fp_help = stderr;
    int min_len = 30000, c, longindex = 0;
    while ((c = getopt_long(argc, argv, "hm:", seq_options, &longindex)) >= 0) {
        if (c == 'h') fp_help = stdout;
        else if (c == 'm') {
            min_len = atoi(optarg);
            if (min_len < 0) {
                fprintf(stderr, "Error: min-len must be a positive integer\n");
                seq_usage(fp_help);
                exit(EXIT_FAILURE);
            }
        } else {
            fprintf(stderr, "Unknown option: %s\n", argv[optind - 1]);
            seq_usage(fp_help);
            exit(EXIT_FAILURE);
        }
    }
    if (argc - optind != 1 || fp_help == stdout) {
        seq_usage(fp_help);
        exit(fp_help == stdout ? EXIT_SUCCESS : EXIT_FAILURE);
    }
    const char *path = argv[optind];
    rd_t r;
    memset(&r, 0, sizeof r);
    r.fp = gzopen(path, "r");
    CORN_F_CHK(r.fp, path);
    r.buf = (unsigned char *)malloc(RD_BUF);
    CORN_MALLOC_CHK(r.buf);

    uint64_t before = 0, after = 0, before_n = 0, after_n = 0;
    while (rd_next(&r) >= 0) {
        before += r.seq.l;
        ++before_n;
        if (r.seq.l >= (size_t)min_len) {
            after += r.seq.l;
            ++after_n;
            printf("@%s", r.name.s);
            if (r.comment.l) printf("\t%s", r.comment.s);
            printf("\n");
            printf("%s\n+\n%s\n", r.seq.s, r.qual.s);
        }
    }
    fprintf(stderr, "total reads: %lu\t%lu bases\t%.2f Gbases\n", (unsigned long)before_n, (unsigned long)before, before / 1e9);
    fprintf(stderr, "reads >= %d: %lu\t%lu bases\t%.2f Gbases\n", min_len, (unsigned long)after_n, (unsigned long)after, after / 1e9);
    gzclose(r.fp);
    free(r.buf); free(r.name.s); free(r.comment.s); free(r.seq.s); free(r.qual.s);
    return 0;
}
