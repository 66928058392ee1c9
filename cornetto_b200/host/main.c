/* cornetto_b200/host/main.c -- command dispatcher of the drop-in binary.
 *
 * Same behaviour as the reference's main() (src/main.c:95-152) for the hot-path commands:
 * dispatch on argv[1], --version / --help, and the Version / CMD / time / RSS footer on stderr.
 * Commands outside the sequence-scan path are not part of this build (scope). */
#include <unistd.h>

#include "cornetto.h"

static int print_usage(FILE *fp)
{
    fprintf(fp, "Usage: cornetto <command> [options]\n\n");
    fprintf(fp, "commands (B200 build: the sequence-scan path):\n");
    fprintf(fp, "   create panel:\n");
    fprintf(fp, "       noboringbits    print no boring bits in an assembly\n");
    fprintf(fp, "   telo:\n");
    fprintf(fp, "       telowin         analyse telomere windows in a fasta file\n");
    fprintf(fp, "       telobreaks      find telomere breaks in a fasta file\n");
    fprintf(fp, "       telofind        find telomere sequences in a fasta file\n");
    fprintf(fp, "       sdust           symmetric DUST (https://github.com/lh3/sdust)\n");
    fprintf(fp, "       telostats       scripts/telostats.sh in one pass: telofind, telowin, merged windows at contig ends, tally\n");
    fprintf(fp, "   misc:\n");
    fprintf(fp, "       fa2bed          create a bed file with assembly contig lengths\n");
    fprintf(fp, "       nx              nx or ngx plot tables\n");
    fprintf(fp, "       report          generate a report table for one or more assemblies\n");
    fprintf(fp, "       seq             print reads longer than a minimum length\n");
    fprintf(fp, "\n");
    fprintf(fp, "       --help, -h      print this help message\n");
    fprintf(fp, "       --version, -V   print version information\n");
    return fp == stdout ? EXIT_SUCCESS : EXIT_FAILURE;
}

/* CUDA start-up is the longest phase of every scan command (0.5-3 s, and it grows with the number of devices the
 * driver has to bring up).  A run uses $CORNETTO_GPUS devices (default 1) starting at $CORNETTO_GPU (default 0):
 * hide the others from the runtime before its first call, so that only those are initialised.  An existing
 * CUDA_VISIBLE_DEVICES is honoured (the selection is made among its entries); $CORNETTO_KEEP_VISIBLE=1 leaves it alone. */
static void restrict_visible_devices(void)
{
    const char *keep = getenv("CORNETTO_KEEP_VISIBLE");
    if (keep && atoi(keep) > 0) return;
    const char *eg = getenv("CORNETTO_GPUS"), *e0 = getenv("CORNETTO_GPU");
    int n = eg && atoi(eg) > 0 ? atoi(eg) : 1;
    const int first = (n == 1 && e0 && atoi(e0) > 0) ? atoi(e0) : 0;
    const char *vis = getenv("CUDA_VISIBLE_DEVICES");
    char out[1024];
    size_t o = 0;
    out[0] = 0;
    if (vis && *vis) {                                   /* entries first .. first + n - 1 of the existing list */
        int idx = 0;
        const char *p = vis;
        while (*p && n > 0) {
            const char *q = strchr(p, ',');
            const size_t len = q ? (size_t)(q - p) : strlen(p);
            if (idx >= first && len > 0 && o + len + 2 < sizeof out) {
                if (o) out[o++] = ',';
                memcpy(out + o, p, len); o += len; out[o] = 0;
                --n;
            }
            ++idx;
            if (!q) break;
            p = q + 1;
        }
        if (o == 0) return;                              /* selection outside the list: let the runtime report it */
    } else {
        for (int i = 0; i < n && o + 16 < sizeof out; ++i) o += (size_t)snprintf(out + o, sizeof out - o, "%s%d", i ? "," : "", first + i);
    }
    setenv("CUDA_VISIBLE_DEVICES", out, 1);
    if (first) setenv("CORNETTO_GPU", "0", 1);           /* the chosen device is now number 0 */
}

int main(int argc, char *argv[])
{
    double realtime0 = realtime();
    int ret = 1;
    restrict_visible_devices();

    if (argc < 2) return print_usage(stderr);
    else if (strcmp(argv[1], "telowin") == 0) ret = telomere_windows_main(argc - 1, argv + 1);
    else if (strcmp(argv[1], "telobreaks") == 0) ret = telomere_breaks_main(argc - 1, argv + 1);
    else if (strcmp(argv[1], "telofind") == 0) ret = find_telomere_main(argc - 1, argv + 1);
    else if (strcmp(argv[1], "sdust") == 0) ret = sdust_main(argc - 1, argv + 1);
    else if (strcmp(argv[1], "telostats") == 0) ret = telostats_main(argc - 1, argv + 1);
    else if (strcmp(argv[1], "noboringbits") == 0) ret = boringbits_main(argc - 1, argv + 1, 0);
    else if (strcmp(argv[1], "boringbits") == 0) ret = boringbits_main(argc - 1, argv + 1, 1);      /* deprecated in the reference, kept like it */
    else if (strcmp(argv[1], "fa2bed") == 0) ret = assbed_main(argc - 1, argv + 1);
    else if (strcmp(argv[1], "nx") == 0) ret = nx_main(argc - 1, argv + 1);
    else if (strcmp(argv[1], "report") == 0) ret = report_main(argc - 1, argv + 1);
    else if (strcmp(argv[1], "seq") == 0) ret = seq_main(argc - 1, argv + 1);
    else if (strcmp(argv[1], "--version") == 0 || strcmp(argv[1], "-V") == 0) {
        fprintf(stdout, "cornetto %s\n", CORNETTO_VERSION);
        exit(EXIT_SUCCESS);
    } else if (strcmp(argv[1], "--help") == 0 || strcmp(argv[1], "-h") == 0) return print_usage(stdout);
    else {
        fprintf(stderr, "[cornetto] Unrecognised command %s\n", argv[1]);
        return print_usage(stderr);
    }
    if (!cornetto_fast_exit()) cornetto_gpu_release();

    fprintf(stderr, "[%s] Version: %s\n", __func__, CORNETTO_VERSION);
    fprintf(stderr, "[%s] CMD:", __func__);
    for (int i = 0; i < argc; ++i) fprintf(stderr, " %s", argv[i]);
    fprintf(stderr, "\n[%s] Real time: %.3f sec; CPU time: %.3f sec; Peak RAM: %.3f GB\n\n", __func__,
            realtime() - realtime0, cputime(), (double)peakrss() / 1024.0 / 1024.0 / 1024.0);
    if (cornetto_fast_exit()) {
        /* everything is written; un-mapping GBs of device and host buffers one by one (and the CUDA
         * runtime's own atexit teardown) takes longer than the scan -- the kernel reclaims them wholesale */
        fflush(stdout);
        fflush(stderr);
        _exit(ret);
    }
    return ret;
}
