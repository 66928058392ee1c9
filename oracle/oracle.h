/* oracle/oracle.h -- CPU restatement of cornetto's sequence-scanning hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or executed
 * from the product (cornetto_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and there only as the checker / baseline.
 *
 * Parity status: the reference ships no golden vectors for this path;
 * the oracle is pinned instead against the UNMODIFIED reference binary compiled from
 * /root/reference by oracle/Makefile (`make ref` -> oracle/_ref/cornetto): see
 * tests/test_oracle_vs_ref.py (runs wherever oracle/_ref exists) and the committed fixtures
 * under tests/golden/ that were generated from that binary by tests/golden/make_golden.py.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef CORNETTO_ORACLE_H
#define CORNETTO_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- FASTA/FASTQ records with kseq semantics (src/kseq.h:184-224) ---------------------- */
typedef struct {
    char   *name;   /* header up to the first isspace() byte (src/kseq.h:195) */
    char   *seq;    /* NUL terminated */
    size_t  len;
} orc_rec_t;

/* Reads every record kseq_read() would return >= 0 for (stops at the first -1/-2).
 * path may be plain or gzip; "-" is stdin.  Returns 0, or -1 if the file cannot be opened. */
int  orc_read_fastx(const char *path, orc_rec_t **recs, size_t *n_recs);
int  orc_parse_fastx_mem(const uint8_t *buf, size_t n, orc_rec_t **recs, size_t *n_recs);
void orc_free_recs(orc_rec_t *recs, size_t n_recs);

/* ---- telofind (src/find_telomere.c:24-81) ------------------------------------------------ */
typedef struct {
    uint32_t strand;      /* 0 forward motif, 1 reverse complement */
    uint64_t start, end;  /* [start,end) on the record, end-start is a multiple of strlen(motif) */
} orc_run_t;

/* rc(): src/find_telomere.c:24-42.  out must hold strlen(motif)+1 bytes. */
void   orc_revcomp_motif(const char *motif, char *out);
/* disambiguate()+find(): src/find_telomere.c:44-81.  seq is NOT modified (a folded copy is
 * scanned).  Runs come out in the reference's print order: all strand 0, then all strand 1. */
size_t orc_telofind(const char *seq, size_t len, const char *motif, orc_run_t **runs);
/* find_telomere_main(): src/find_telomere.c:83-111 -- writes the 6-column TSV. */
int    orc_telofind_file(const char *path, const char *motif, FILE *out);

/* ---- telowin (src/telomere_windows.c:28-86) ---------------------------------------------- */
typedef struct { int32_t start, end, car; } orc_win_t;
/* threshold adjustment: src/telomere_windows.c:53-54 */
double orc_telowin_threshold(double thr, double identity_percent);
/* process_scaffold(): src/telomere_windows.c:28-43 on a 0/1 byte map. */
size_t orc_telowin_contig(const uint8_t *marked, int length, double thr_adj, orc_win_t **wins);
/* telomere_windows_main(): text in (telofind TSV) -> "Window\t..." lines out. */
int    orc_telowin_file(const char *path, double identity_percent, int have_thr, double thr, FILE *out);

/* ---- sdust (src/sdust/sdust.c:23-171) ---------------------------------------------------- */
/* Same contract as the reference's sdust(): returns malloc'd start<<32|finish list. */
uint64_t *orc_sdust(const uint8_t *seq, int l_seq, int T, int W, int *n);
int       orc_sdust_file(const char *path, int T, int W, FILE *out);

/* ---- telobreaks (src/telomere_breaks.c:47-172, khash order src/khash.h:230-348,395-400) --- */
/* order[i] = index into names[] of the i-th key met when iterating a khash string map into
 * which names[0..n) were kh_put in that sequence (duplicates collapse to the first key
 * pointer, as kh_put leaves present keys untouched).  Returns the number of distinct keys. */
size_t orc_khash_order(const char *const *names, size_t n, size_t *order);
int    orc_telobreaks_files(const char *lens, const char *sdust, const char *telomere, FILE *out);

/* ---- fa2bed (src/assbed.c:97-100) -------------------------------------------------------- */
int    orc_fa2bed_file(const char *path, FILE *out);

/* ---- noboringbits / boringbits: windowed depth scan, src/boringbits_main.c:179-493 ---------------------- */
typedef struct {
    int window_size, window_inc;                 /* -w 2500, -i 50 */
    float low_cov_thresh, high_cov_thresh, low_mq_cov_thresh;   /* -L 0.4, -H 2.5, -Q 0.4 */
    int min_ctg_len, edge_len;                   /* -m 1000000, -e 100000 */
    int boring;                                  /* 1: boringbits, 0: noboringbits */
} orc_bits_opt_t;
void orc_bits_defaults(orc_bits_opt_t *o);
/* 0 ok; 1 = the reference would print an ERROR and exit(EXIT_FAILURE) (message in err, if not NULL) */
int  orc_bits_files(const char *cov_total, const char *cov_mq, const orc_bits_opt_t *o, FILE *out, FILE *err);

#ifdef __cplusplus
}
#endif
#endif
