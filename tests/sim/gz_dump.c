/* tests/sim/gz_dump.c -- TEST INFRASTRUCTURE: the text gzsrc.c (cornetto_b200/host) delivers for a .gz file, read in
 * requests of the given size, to stdout.  usage: gz_dump <file.gz> <request bytes>      exit 3 = damaged input */
#include "../../cornetto_b200/host/cornetto.h"

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    gzsrc_t *g = gzsrc_open(argv[1]);
    if (!g) return 2;
    const uint64_t n = strtoull(argv[2], NULL, 10);
    uint8_t *buf = (uint8_t *)malloc(n + 1);
    int eof = 0;
    fprintf(stderr, "bgzf=%d\n", gzsrc_is_bgzf(g));
    while (!eof) {
        const int64_t k = gzsrc_read(g, buf, n, &eof);
        if (k < 0) return 3;
        fwrite(buf, 1, (size_t)k, stdout);
        if (k == 0 && !eof && n < 65536) return 4;      /* a request smaller than a BGZF member can never be served */
    }
    gzsrc_close(g);
    return 0;
}
