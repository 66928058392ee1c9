/* tests/sim/reader_dump.c -- dumps what the host reader (cornetto_b200/host/fastx.c) delivers:
 * one line per record  name \t length \t hex(sequence bytes).  TEST INFRASTRUCTURE: lets the
 * CPU-only suite compare the reader with the oracle's kseq restatement, also across batch
 * boundaries (CORNETTO_BATCH_MB / tiny capacities). */
#include "../../cornetto_b200/host/cornetto.h"

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    /* $READER_OFFSET: start at a byte offset that is a record boundary (what the CLI does when the device parser
     * hands the rest of a file to the serial reader) */
    const char *off = getenv("READER_OFFSET");
    fastx_t *fx = off ? fastx_open_at(argv[1], (uint64_t)atoll(off)) : fastx_open(argv[1]);
    if (!fx) return 1;
    uint64_t cap = (uint64_t)atoll(argv[2]);
    rec_batch_t *b = rec_batch_create(cap, argc > 3 ? (uint32_t)atoi(argv[3]) : 1024);
    while (rec_batch_fill(b, fx) > 0) {
        corn_batch_t v;
        corn_hbatch_view(b->hb, &v);
        for (uint32_t i = 0; i < v.n_rec; ++i) {
            printf("%s\t%u\t", b->name[i], v.length[i]);
            const uint8_t *p = v.seq + v.offset[i];
            for (uint32_t k = 0; k < v.length[i]; ++k) printf("%02x", p[k]);
            /* layout contract: zero padding up to the next record / end */
            uint64_t end = i + 1 < v.n_rec ? v.offset[i + 1] : v.total_bytes;
            for (uint64_t k = v.offset[i] + v.length[i]; k < end; ++k) if (v.seq[k]) { printf("\tBADPAD"); break; }
            if (v.offset[i] % CORN_ALIGN) printf("\tBADALIGN");
            printf("\n");
        }
    }
    rec_batch_destroy(b);
    fastx_close(fx);
    return 0;
}
