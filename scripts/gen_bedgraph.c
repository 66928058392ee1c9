/* scripts/gen_bedgraph.c -- synthetic per-base depth tables for scripts/bits_cli_bench.py (bedtools genomecov -d -style:
 * one line per base, name \t pos \t pos+1 \t depth).  usage: gen_bedgraph <cov-total.bg> <cov-mq.bg> <contigs> <bases per contig>
 * Depth: ~30 with noise; every 3 Mb a 40 kb low-coverage stretch, a 30 kb high-coverage one and a 50 kb stretch whose
 * MAPQ>=20 share drops to a fifth -- so both commands have something to print. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static char *put_u32(char *p, uint32_t v)
{
    char tmp[12];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    FILE *f1 = fopen(argv[1], "wb"), *f2 = fopen(argv[2], "wb");
    if (!f1 || !f2) return 1;
    const int n_ctg = atoi(argv[3]);
    const uint32_t len = (uint32_t)strtoul(argv[4], NULL, 10);
    static char b1[1 << 22], b2[1 << 22];
    char *p1 = b1, *p2 = b2;
    uint64_t s = 88172645463325252ull;
    for (int c = 0; c < n_ctg; ++c) {
        char name[32];
        const int nl = snprintf(name, sizeof name, "contig_%d", c + 1);
        for (uint32_t i = 0; i < len; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            uint32_t d = 22 + (uint32_t)(s & 15), q;
            const uint32_t ph = i % 3000000u;
            if (ph >= 1000000u && ph < 1040000u) d = 3 + (uint32_t)(s & 3);
            else if (ph >= 2000000u && ph < 2030000u) d = 100 + (uint32_t)(s & 31);
            q = d - (uint32_t)((s >> 8) & 3);
            if (ph >= 2500000u && ph < 2550000u) q = d / 5;
            for (int k = 0; k < nl; ++k) { *p1++ = name[k]; *p2++ = name[k]; }
            *p1++ = '\t'; p1 = put_u32(p1, i); *p1++ = '\t'; p1 = put_u32(p1, i + 1); *p1++ = '\t'; p1 = put_u32(p1, d); *p1++ = '\n';
            *p2++ = '\t'; p2 = put_u32(p2, i); *p2++ = '\t'; p2 = put_u32(p2, i + 1); *p2++ = '\t'; p2 = put_u32(p2, q); *p2++ = '\n';
            if (p1 - b1 > (1 << 22) - 128) { fwrite(b1, 1, (size_t)(p1 - b1), f1); p1 = b1; }
            if (p2 - b2 > (1 << 22) - 128) { fwrite(b2, 1, (size_t)(p2 - b2), f2); p2 = b2; }
        }
    }
    fwrite(b1, 1, (size_t)(p1 - b1), f1); fwrite(b2, 1, (size_t)(p2 - b2), f2);
    return fclose(f1) | fclose(f2);
}
