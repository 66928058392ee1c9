/* oracle/oracle_main.c -- CLI wrapper around the oracle (TEST INFRASTRUCTURE ONLY).
 * Same sub-commands, arguments and stdout text as the reference's hot-path commands
 * (src/main.c:111-122), so test scripts can diff the three implementations directly. */
#include "oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: oracle_cornetto <telofind|telowin|telobreaks|sdust|fa2bed> ...\n"); return 1; }
    const char *cmd = argv[1];
    if (!strcmp(cmd, "telofind")) {
        if (argc < 3) return 1;
        return orc_telofind_file(argv[2], argc >= 4 ? argv[3] : "TTAGGG", stdout) ? 1 : 0;
    } else if (!strcmp(cmd, "telowin")) {
        if (argc < 4) return 1;
        int have = (argc == 5);           /* src/telomere_windows.c:48: honoured only when argc==4 there */
        return orc_telowin_file(argv[2], atof(argv[3]), have, have ? atof(argv[4]) : 0.0, stdout) ? 1 : 0;
    } else if (!strcmp(cmd, "telobreaks")) {
        if (argc < 5) return 1;
        return orc_telobreaks_files(argv[2], argv[3], argv[4], stdout) ? 1 : 0;
    } else if (!strcmp(cmd, "sdust")) {
        int W = 64, T = 20; const char *file = NULL;
        for (int i = 2; i < argc; ++i) {
            if (!strcmp(argv[i], "-w") && i + 1 < argc) W = atoi(argv[++i]);
            else if (!strcmp(argv[i], "-t") && i + 1 < argc) T = atoi(argv[++i]);
            else if (!strncmp(argv[i], "-w", 2) && argv[i][2]) W = atoi(argv[i] + 2);
            else if (!strncmp(argv[i], "-t", 2) && argv[i][2]) T = atoi(argv[i] + 2);
            else if (!file) file = argv[i];
        }
        if (!file) return 1;
        return orc_sdust_file(file, T, W, stdout) ? 1 : 0;
    } else if (!strcmp(cmd, "fa2bed")) {
        if (argc < 3) return 1;
        return orc_fa2bed_file(argv[2], stdout) ? 1 : 0;
    }
    if (!strcmp(cmd, "noboringbits") || !strcmp(cmd, "boringbits")) {
        orc_bits_opt_t o;
        orc_bits_defaults(&o);
        o.boring = !strcmp(cmd, "boringbits");
        const char *tot = NULL, *mq = NULL;
        for (int i = 2; i < argc; ++i) {
            if (argv[i][0] == '-' && argv[i][1] && i + 1 < argc) {
                const char *v = argv[++i];
                switch (argv[i - 1][1]) {
                case 'q': mq = v; break;
                case 'w': o.window_size = atoi(v); break;
                case 'i': o.window_inc = atoi(v); break;
                case 'L': o.low_cov_thresh = atof(v); break;
                case 'H': o.high_cov_thresh = atof(v); break;
                case 'Q': o.low_mq_cov_thresh = atof(v); break;
                case 'm': o.min_ctg_len = atoi(v); break;
                case 'e': o.edge_len = atoi(v); break;
                default: break;
                }
            } else if (!tot) tot = argv[i];
        }
        if (!tot || !mq) return 1;
        return orc_bits_files(tot, mq, &o, stdout, stderr);
    }
    fprintf(stderr, "unknown command %s\n", cmd);
    return 1;
}
