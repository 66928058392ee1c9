/* tests/sim/depthtxt_dump.c -- TEST INFRASTRUCTURE: what cornetto_b200/host/depthtxt.c makes of two depth tables.
 * usage: depthtxt_dump <cov-total.bg> <cov-mq.bg> serial|parallel [threads] [block bytes]
 * stdout: a text header (contigs, sums) followed by the two uint16 arrays; exit 3 = the parallel reader declined. */
#include "../../cornetto_b200/host/cornetto.h"

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    depth_text_t t1, t2;
    depth_table_t T;
    depthtxt_open(&t1, argv[1]);
    depthtxt_open(&t2, argv[2]);
    if (strcmp(argv[3], "serial") == 0) depthtxt_load_serial(&t1, &t2, &T);
    else if (depthtxt_load_parallel(&t1, &t2, argc > 4 ? atoi(argv[4]) : 4, argc > 5 ? (size_t)strtoull(argv[5], NULL, 10) : 4096, &T) != 0) return 3;
    printf("n_ctg %zu n_tot %llu tot_depth %.0f tot_mq %.0f\n", T.n_ctg, (unsigned long long)T.n_tot, T.tot_depth, T.tot_mq);
    for (size_t i = 0; i < T.n_ctg; ++i) printf("%s %llu %u\n", T.ctg[i].name, (unsigned long long)T.ctg[i].off, T.ctg[i].len);
    fwrite(T.depth, sizeof(uint16_t), T.n_tot, stdout);
    fwrite(T.mq, sizeof(uint16_t), T.n_tot, stdout);
    depth_table_free(&T);
    depthtxt_close(&t1); depthtxt_close(&t2);
    return 0;
}
