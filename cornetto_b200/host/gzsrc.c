/* cornetto_b200/host/gzsrc.c -- decompressed text of a .gz input in large blocks, for the device parser.
 *
 * The reference reads every input through gzread() in 16 KiB pieces on its one thread (src/find_telomere.c:96,
 * src/kseq.h:234); zlib then delivers ~0.3 GB/s however fast the scan is.  Assemblies usually ship as .fa.gz.
 * Two cases:
 *   BGZF (bgzip, what samtools/htslib write and index): a chain of independent gzip members of at most 64 KiB, each
 *        announcing its compressed size in a 'BC' extra field (SAM spec 4.1) and its uncompressed size in its last four
 *        bytes.  A request for n bytes of text walks the member headers, lays the members out at their prefix-summed
 *        offsets in the destination and inflates them on GZ_THREADS threads, straight into place.
 *   any other gzip file (one member or several): one inflate stream (zlib, concatenated members as gzread handles
 *        them), on the thread that feeds the pipeline -- so it at least runs beside the GPU work and the text
 *        formatting of the previous block instead of in front of them.
 * Either way the text goes to corn_gpu_ingest() like the blocks of a plain file.  Same bytes as gzread(): the tests
 * compare with the oracle, which reads through gzread. */
#include <fcntl.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "cornetto.h"

#define GZ_THREADS 8

struct gzsrc {
    int       fd;
    const uint8_t *map;        /* the compressed file, mapped */
    uint64_t  size, pos;       /* compressed size / next compressed byte */
    int       bgzf, eof, failed;
    z_stream  zs;              /* plain gzip: the running stream */
    int       zs_open;
};

/* gzip member header at p (n bytes available): returns the BGZF block size (BSIZE + 1) or 0 if this is not a BGZF member */
static uint32_t bgzf_block_size(const uint8_t *p, uint64_t n)
{
    if (n < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
    const uint32_t xlen = p[10] | (uint32_t)p[11] << 8;
    if (12 + (uint64_t)xlen > n) return 0;
    for (uint32_t o = 0; o + 4 <= xlen;) {
        const uint8_t *f = p + 12 + o;
        const uint32_t slen = f[2] | (uint32_t)f[3] << 8;
        if (f[0] == 'B' && f[1] == 'C' && slen == 2 && o + 6 <= xlen) return (f[4] | (uint32_t)f[5] << 8) + 1u;
        o += 4 + slen;
    }
    return 0;
}

gzsrc_t *gzsrc_open(const char *path)
{
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return NULL;
    struct stat sb;
    if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode) || sb.st_size < 18) { close(fd); return NULL; }
    void *m = mmap(NULL, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { close(fd); return NULL; }
    madvise(m, (size_t)sb.st_size, MADV_SEQUENTIAL);
    gzsrc_t *g = (gzsrc_t *)calloc(1, sizeof *g);
    CORN_MALLOC_CHK(g);
    g->fd = fd; g->map = (const uint8_t *)m; g->size = (uint64_t)sb.st_size;
    g->bgzf = bgzf_block_size(g->map, g->size) != 0;
    return g;
}

void gzsrc_close(gzsrc_t *g)
{
    if (!g) return;
    if (g->zs_open) inflateEnd(&g->zs);
    munmap((void *)g->map, (size_t)g->size);
    close(g->fd);
    free(g);
}

int gzsrc_is_bgzf(const gzsrc_t *g) { return g->bgzf; }

/* ---- BGZF: members inflated in parallel, straight into their place ---- */
typedef struct { const uint8_t *src; uint32_t clen; uint8_t *dst; uint32_t ulen; } bg_job_t;
typedef struct { bg_job_t *job; size_t n, stride, first; int failed; pthread_t th; } bg_part_t;

static void *bg_worker(void *p)
{
    bg_part_t *w = (bg_part_t *)p;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) { w->failed = 1; return NULL; }          /* raw deflate: the header is skipped by hand */
    for (size_t i = w->first; i < w->n; i += w->stride) {
        const bg_job_t *j = &w->job[i];
        if (j->ulen == 0) continue;                                               /* (the empty end-of-file member) */
        const uint32_t xlen = j->src[10] | (uint32_t)j->src[11] << 8;
        const uint32_t hdr = 12 + xlen;
        if (j->clen < hdr + 8) { w->failed = 1; break; }
        inflateReset(&zs);
        zs.next_in = (Bytef *)(j->src + hdr); zs.avail_in = j->clen - hdr - 8;
        zs.next_out = j->dst; zs.avail_out = j->ulen;
        const int r = inflate(&zs, Z_FINISH);
        if (r != Z_STREAM_END || zs.avail_out != 0) { w->failed = 1; break; }
    }
    inflateEnd(&zs);
    return NULL;
}

static int64_t bgzf_read(gzsrc_t *g, uint8_t *dst, uint64_t n)
{
    size_t cap = 1 << 16, nj = 0;
    bg_job_t *job = (bg_job_t *)malloc(cap * sizeof *job);
    CORN_MALLOC_CHK(job);
    uint64_t out = 0;
    while (g->pos < g->size) {
        const uint32_t bs = bgzf_block_size(g->map + g->pos, g->size - g->pos);
        if (bs == 0 || g->pos + bs > g->size) { free(job); return -1; }          /* not BGZF after all / truncated */
        const uint8_t *p = g->map + g->pos;
        const uint32_t ulen = p[bs - 4] | (uint32_t)p[bs - 3] << 8 | (uint32_t)p[bs - 2] << 16 | (uint32_t)p[bs - 1] << 24;
        if (ulen > 65536u) { free(job); return -1; }
        if (out + ulen > n) break;                                               /* does not fit any more: next call */
        if (nj == cap) { cap *= 2; job = (bg_job_t *)realloc(job, cap * sizeof *job); CORN_MALLOC_CHK(job); }
        job[nj].src = p; job[nj].clen = bs; job[nj].dst = dst + out; job[nj].ulen = ulen;
        ++nj;
        out += ulen;
        g->pos += bs;
    }
    if (g->pos >= g->size) g->eof = 1;
    bg_part_t part[GZ_THREADS];
    const int nt = nj < 64 ? 1 : GZ_THREADS;
    int failed = 0;
    for (int t = 0; t < nt; ++t) {
        part[t].job = job; part[t].n = nj; part[t].stride = (size_t)nt; part[t].first = (size_t)t; part[t].failed = 0; part[t].th = 0;
        if (t == nt - 1 || pthread_create(&part[t].th, NULL, bg_worker, &part[t]) != 0) { part[t].th = 0; bg_worker(&part[t]); }
    }
    for (int t = 0; t < nt; ++t) { if (part[t].th) pthread_join(part[t].th, NULL); failed |= part[t].failed; }
    free(job);
    return failed ? -1 : (int64_t)out;
}

/* ---- any other gzip file: one stream, concatenated members as gzread() handles them ---- */
static int64_t plain_read(gzsrc_t *g, uint8_t *dst, uint64_t n)
{
    if (!g->zs_open) {
        memset(&g->zs, 0, sizeof g->zs);
        if (inflateInit2(&g->zs, 15 + 16) != Z_OK) return -1;
        g->zs_open = 1;
    }
    uint64_t out = 0;
    while (out < n && !g->eof) {
        if (g->pos >= g->size) { g->eof = 1; break; }
        const uint64_t in_left = g->size - g->pos, out_left = n - out;
        g->zs.next_in = (Bytef *)(g->map + g->pos);
        g->zs.avail_in = in_left > (1u << 30) ? (1u << 30) : (uInt)in_left;
        g->zs.next_out = dst + out;
        g->zs.avail_out = out_left > (1u << 30) ? (1u << 30) : (uInt)out_left;
        const uInt in0 = g->zs.avail_in, out0 = g->zs.avail_out;
        const int r = inflate(&g->zs, Z_NO_FLUSH);
        g->pos += in0 - g->zs.avail_in;
        out += out0 - g->zs.avail_out;
        if (r == Z_STREAM_END) {
            /* end of a member: another one follows if the next bytes are a gzip magic (gzread's rule); trailing
             * garbage ends the data */
            if (g->pos + 2 <= g->size && g->map[g->pos] == 0x1f && g->map[g->pos + 1] == 0x8b) inflateReset(&g->zs);
            else g->eof = 1;
        } else if (r != Z_OK && r != Z_BUF_ERROR) return -1;
        else if (r == Z_BUF_ERROR && in0 - g->zs.avail_in == 0 && out0 - g->zs.avail_out == 0) { g->eof = 1; break; }   /* truncated stream */
    }
    return (int64_t)out;
}

/* up to n bytes of text at dst; fewer only at the end of the input (*eof) -- or, BGZF, when the next member would not fit.
 * -1: the file is damaged (the caller falls back to the serial reader, which reports it the way gzread does) */
int64_t gzsrc_read(gzsrc_t *g, uint8_t *dst, uint64_t n, int *eof)
{
    int64_t r = g->failed ? -1 : (g->bgzf ? bgzf_read(g, dst, n) : plain_read(g, dst, n));
    if (r < 0) g->failed = 1;
    *eof = g->eof;
    return r;
}
