/* tests/sim/format_check.c -- outbuf_format_parallel() must produce the bytes of the serial loop
 * (TEST INFRASTRUCTURE for cornetto_b200/host/misc.c).  usage: format_check <n_items>; prints OK <bytes>. */
#include "../../cornetto_b200/host/cornetto.h"

static void fmt(outbuf_t *ob, uint64_t begin, uint64_t end, void *arg)
{
    const uint64_t mul = *(const uint64_t *)arg;
    for (uint64_t i = begin; i < end; ++i) {
        outbuf_str(ob, "ctg", 3); outbuf_u64(ob, i % 24);
        outbuf_chr(ob, '\t'); outbuf_u64(ob, i * mul);
        outbuf_chr(ob, '\t'); outbuf_i32(ob, (int32_t)(i * 2654435761u));
        outbuf_chr(ob, '\n');
    }
}

int main(int argc, char **argv)
{
    const uint64_t n = argc > 1 ? (uint64_t)atoll(argv[1]) : 1000000, mul = 977;
    outbuf_t a, b;
    outbuf_init(&a, NULL); outbuf_init(&b, NULL);
    fmt(&a, 0, n, (void *)&mul);
    outbuf_format_parallel(&b, n, fmt, (void *)&mul);
    if (a.n != b.n || memcmp(a.buf, b.buf, a.n) != 0) { printf("DIFFER %zu %zu\n", a.n, b.n); return 1; }
    printf("OK %zu\n", a.n);
    return 0;
}
