/* cornetto_b200/host/telowin_main.c -- `cornetto telowin <telofind.tsv> <identity%> [threshold]`.
 *
 * Same contract as telomere_windows_main(), src/telomere_windows.c:45-86: usage + EXIT_FAILURE
 * with fewer than two arguments, threshold = argv[3] only when exactly three arguments are
 * given (else 0.4), adjusted by pow(identity/100, 6) in double, the "Given error rate ..." line
 * on stderr, consecutive lines with the same first column form one scaffold of length atoi(col2),
 * every line paints [atoi(col4), atoi(col5)), and windows print as
 *     Window \t name \t len \t i \t i+den \t %.3g(car/den).
 * The byte painting and the 1000/200 window loop are replaced by corn_gpu_telowin(). */
#include <ctype.h>
#include <math.h>

#include "cornetto.h"

#define LINE_CAP 2048

static int split_ws(char *line, char **tok, int maxtok)
{
    int n = 0;
    char *p = line;
    while (n < maxtok) {
        while (*p && isspace((unsigned char)*p)) ++p;
        if (!*p) break;
        tok[n++] = p;
        while (*p && !isspace((unsigned char)*p)) ++p;
        if (*p) *p++ = 0;
    }
    return n;
}

int telomere_windows_main(int argc, char *argv[])
{
    if (argc < 3) {
        fprintf(stderr, "Usage: cornetto telowin <input_file> <identity> <threshold>\n");
        fprintf(stderr, "This program analyzes telomere windows in a genome assembly.\n");
        fprintf(stderr, "Example usage: cornetto telowin input.telomere 99.9 0.4\n");
        return EXIT_FAILURE;
    }
    double threshold = 0.4;
    if (argc == 4) threshold = atof(argv[3]);
    const char *input_file = argv[1];
    const double identity = atof(argv[2]) / 100;
    threshold = threshold * pow(identity, 6);
    fprintf(stderr, "Given error rate of %.6f running with adjusted threshold of %.6f due to survival prob %.6f\n",
            identity, threshold, pow(identity, 6));

    FILE *fp = fopen(input_file, "r");
    CORN_F_CHK(fp, input_file);
    cornetto_gpu_prefetch();                 /* the driver starts up while the text is parsed */

    /* scaffolds in file order (a name that comes back later is a new scaffold, as in the reference) */
    char **names = NULL;
    uint32_t *lens = NULL;
    size_t n_sc = 0, m_sc = 0;
    corn_run_t *runs = NULL;
    size_t n_run = 0, m_run = 0;
    char line[LINE_CAP];
    while (fgets(line, sizeof line, fp)) {
        char *t[6];
        /* columns 1, 2, 4 and 5 are all the reference uses (sscanf takes whatever is there, src/telomere_windows.c:67-77):
         * a line cut after the fifth column still paints */
        if (split_ws(line, t, 6) < 5) continue;
        if (n_sc == 0 || strcmp(t[0], names[n_sc - 1]) != 0) {
            if (n_sc == m_sc) {
                m_sc = m_sc ? m_sc * 2 : 64;
                names = (char **)realloc(names, m_sc * sizeof(char *));
                lens = (uint32_t *)realloc(lens, m_sc * sizeof(uint32_t));
                CORN_MALLOC_CHK(names); CORN_MALLOC_CHK(lens);
            }
            names[n_sc] = strdup(t[0]);
            const int L = atoi(t[1]);
            lens[n_sc] = L > 0 ? (uint32_t)L : 0;
            ++n_sc;
        }
        const int s = atoi(t[3]), e = atoi(t[4]);
        if (s < e && e > 0) {
            if (n_run == m_run) {
                m_run = m_run ? m_run * 2 : 1024;
                runs = (corn_run_t *)realloc(runs, m_run * sizeof(corn_run_t));
                CORN_MALLOC_CHK(runs);
            }
            runs[n_run].rec = (uint32_t)(n_sc - 1);
            runs[n_run].strand = 0;
            runs[n_run].start = s > 0 ? (uint32_t)s : 0;
            runs[n_run].end = (uint32_t)e;
            ++n_run;
        }
    }
    fclose(fp);

    if (n_sc) {
        corn_ctx_t *ctx = cornetto_gpu();
        corn_hits_t hits;
        hits.run = runs; hits.n_run = n_run; hits._owner = NULL;
        corn_contigs_t contigs;
        contigs.length = lens; contigs.n = (uint32_t)n_sc;
        corn_windows_t w;
        int r = corn_gpu_telowin(ctx, &hits, &contigs, threshold, &w);
        if (r != CORN_OK) cornetto_gpu_die("telowin", r);
        for (uint64_t i = 0; i < w.n_win; ++i) {
            const corn_window_t *x = &w.win[i];
            const int den = (int)(x->end - x->start);
            printf("Window\t%s\t%d\t%d\t%d\t%.3g\n", names[x->rec], (int)lens[x->rec], (int)x->start, (int)x->end,
                   (double)x->car / den);
        }
        corn_gpu_windows_free(&w);
    }
    for (size_t i = 0; i < n_sc; ++i) free(names[i]);
    free(names); free(lens); free(runs);
    return EXIT_SUCCESS;
}
