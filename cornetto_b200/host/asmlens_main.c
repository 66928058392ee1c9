/* cornetto_b200/host/asmlens_main.c -- `cornetto nx` and `cornetto report`: contig-length statistics.
 *
 * By-products of the record table: both commands only need the sequence length of
 * every record, which the reader delivers without copying a base anywhere.  Contracts restated from
 * nx_main(), src/nx.c:60-158, and report_main(), src/report.c:60-164:
 *   nx      [-g SIZE] <assembly.fasta>    "#x\tcontig_len" header, then for the contigs in descending length
 *           two lines each: (percent before, length) and (percent after, length), percent = cumulative length
 *           over the genome size (-g, with K/M/G suffix, rounded as the reference's mm_parse_num) or over the
 *           total length, printed with %f;
 *   report  <assembly.fasta> ...          one header line, then per file: name, number of contigs, largest
 *           contig, N50 and N90 in Mbases (%.3f); N50/N90 = length of the first contig (descending) at which
 *           the cumulative length reaches half / 90 % of the total.
 * getopt_long() is the parser in the reference too, so option permutation and its own diagnostics are identical. */
#include <getopt.h>

#include "cornetto.h"

static const struct option nx_options[] = {
    { "genome-size", required_argument, 0, 'g' },
    { "verbose", required_argument, 0, 'v' },
    { "help", no_argument, 0, 'h' },
    { 0, 0, 0, 0 } };

static const struct option report_options[] = {
    { "verbose", required_argument, 0, 'v' },
    { "help", no_argument, 0, 'h' },
    { 0, 0, 0, 0 } };

static void nx_usage(FILE *fp)
{
    fprintf(fp, "Usage: cornetto nx <assembly.fasta> \n");
    fprintf(fp, "   -g STR                     genome size (e.g. 3.1G). if unspecified, will use total contig length\n");
    fprintf(fp, "   -h                         help\n");
}

static void report_usage(FILE *fp)
{
    fprintf(fp, "Usage: cornetto report <assembly.fasta> ... \n");
    fprintf(fp, "   -h                         help\n");
}

/* "3.1G" -> 3100000000 (src/misc.c:72-84) */
static int64_t parse_size(const char *str)
{
    char *end;
    double x = strtod(str, &end);
    if (*end == 'G' || *end == 'g') x *= 1e9;
    else if (*end == 'M' || *end == 'm') x *= 1e6;
    else if (*end == 'K' || *end == 'k') x *= 1e3;
    return (int64_t)(x + .499);
}

static int cmp_u64(const void *a, const void *b)
{
    const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* lengths of all records of a file, ascending; exits like F_CHK when the file cannot be opened */
static uint64_t *sorted_lengths(const char *path, uint64_t *n_out, uint64_t *sum_out)
{
    fastx_t *fx = fastx_open(path);
    CORN_F_CHK(fx, path);
    uint64_t n = 0, cap = 128, sum = 0;
    uint64_t *len = (uint64_t *)malloc(cap * sizeof(uint64_t));
    CORN_MALLOC_CHK(len);
    uint8_t *scratch = (uint8_t *)malloc(1 << 20);
    CORN_MALLOC_CHK(scratch);
    while (fastx_next(fx)) {
        uint64_t l = 0;
        int done = 0;
        while (!done) l += fastx_seq(fx, scratch, 1 << 20, &done);
        if (fastx_finish(fx) != 0) break;
        if (n == cap) { cap *= 2; len = (uint64_t *)realloc(len, cap * sizeof(uint64_t)); CORN_MALLOC_CHK(len); }
        len[n++] = l;
        sum += l;
    }
    free(scratch);
    fastx_close(fx);
    qsort(len, n, sizeof(uint64_t), cmp_u64);
    *n_out = n; *sum_out = sum;
    return len;
}

int nx_main(int argc, char *argv[])
{
    FILE *fp_help = stderr;
    int64_t genome_size = -1;
    int c, longindex = 0;
    while ((c = getopt_long(argc, argv, "g:h", nx_options, &longindex)) >= 0) {
        if (c == 'h') fp_help = stdout;
        else if (c == 'g') {
            genome_size = parse_size(optarg);
            if (genome_size <= 0) { CORN_ERROR("%s", "Genome size should be larger than 0."); exit(EXIT_FAILURE); }
        }
    }
    if (argc - optind != 1 || fp_help == stdout) {
        nx_usage(fp_help);
        exit(fp_help == stdout ? EXIT_SUCCESS : EXIT_FAILURE);
    }
    uint64_t n, sum;
    uint64_t *len = sorted_lengths(argv[optind], &n, &sum);
    fprintf(stdout, "#x\tcontig_len\n");
    uint64_t cumsum = 0;
    double percent = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t l = len[n - i - 1];
        fprintf(stdout, "%f\t%lu\n", percent, (unsigned long)l);
        cumsum += l;
        percent = genome_size > 0 ? (double)cumsum / genome_size * 100 : (double)cumsum / sum * 100;
        fprintf(stdout, "%f\t%lu\n", percent, (unsigned long)l);
    }
    free(len);
    return 0;
}

int report_main(int argc, char *argv[])
{
    FILE *fp_help = stderr;
    int c, longindex = 0;
    while ((c = getopt_long(argc, argv, "h", report_options, &longindex)) >= 0)
        if (c == 'h') fp_help = stdout;
    if (argc - optind < 1 || fp_help == stdout) {
        report_usage(fp_help);
        exit(fp_help == stdout ? EXIT_SUCCESS : EXIT_FAILURE);
    }
    fprintf(stdout, "#asm\tNcontigs\tLargestcontig(Mbase)\tN50(Mbase)\tN90(Mbase)\n");
    while (optind < argc) {
        const char *fasta = argv[optind++];
        fprintf(stdout, "%s\t", fasta);
        uint64_t n, sum;
        uint64_t *len = sorted_lengths(fasta, &n, &sum);
        uint64_t cumsum = 0, n50 = 0, n90 = 0;
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t l = len[n - i - 1];
            cumsum += l;
            if (cumsum >= sum * 0.5 && n50 == 0) n50 = l;
            if (cumsum >= sum * 0.9 && n90 == 0) n90 = l;
        }
        /* (the reference reads length[-1] for a file without records; we print a largest contig of 0 there) */
        fprintf(stdout, "%ld\t%.3f\t%.3f\t%.3f\n", (long)n, (n ? len[n - 1] : 0) / 1e6, n50 / 1e6, n90 / 1e6);
        free(len);
    }
    return 0;
}
