// cornetto_b200/csrc/ingest_core.cuh -- line rules of the device-side FASTA/FASTQ parser.
//
// What is replaced: kseq_read() (src/kseq.h:184-224) on top of ks_getuntil2() (src/kseq.h:102-141).
// Restated:
//   * a record starts at the next '>' or '@' byte -- ANY byte when the previous record ended with
//     its quality block or at the start of the input (:188-191), the first byte of a line when the
//     previous record's sequence loop stopped there (:199,:213);
//   * the name is the header text up to the first isspace() byte, the rest of that line is the
//     comment (:195-196);
//   * sequence: for every following line, look at its first byte c (:199): '>', '+' or '@' ends the
//     sequence; '\n' (an empty line) is skipped (:200); otherwise c and the rest of the line are
//     appended (:201-202) and ONE trailing '\r' is dropped if the sequence so far is longer than one
//     byte (:138; note the test is on the whole string, not on the line);
//   * c == '+': the rest of that line is skipped (:216), then whole lines are appended to the quality
//     string until it is at least as long as the sequence (:218); a different length is the error -2
//     (:221), after which the reference's read loops stop.
//
// The device parser handles only text on which these rules collapse to LINE-LOCAL ones, checks
// that this is the case, and otherwise reports "irregular" so that the host's serial reader (which
// implements all of the above) takes over.  With line k = text[s, e) (e = position of its '\n', or
// the end of the text for a last line without one):
//
//   len'(k) = (e - s) - [ (e - s) > 1 and text[e-1] == '\r' ]
//
// FASTA mode (text[0] == '>'):  a line is a header iff text[s] == '>'.  Every other line must not
//   start with '@' or '+' and must not be exactly "\r" (whether that byte is kept depends on the
//   sequence being empty so far: not line-local).  Then the header lines delimit the records and a
//   sequence line contributes len'(k) bytes (an empty line: 0), because every non-empty sequence
//   line of length >= 2 makes the string longer than one byte.
// FASTQ mode (text[0] == '@'):  lines are taken four at a time: line 0 must start with '@'; line 1
//   must be non-empty and must not start with '>', '+' or '@' (else the sequence loop would end at
//   once); line 2 must start with '+'; line 3 must satisfy len'(3) == len'(1) -- shorter would pull
//   in a further quality line, longer is the error -2.  After line 3 the reader looks for the next
//   '@'/'>' byte, which by the line-0 rule is the first byte of the next group.  At the end of the
//   input up to three left-over lines are allowed if each is empty or a lone "\r" (the reader would
//   skip them while looking for a header).
// A NUL byte inside a sequence line is also irregular (0x00 is the layout's padding value).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define ING_HD __host__ __device__ __forceinline__
#else
#define ING_HD static inline
#endif

enum { ING_MODE_FASTA = 0, ING_MODE_FASTQ = 1 };

// result of looking at one line
struct ing_line {
    uint32_t contrib;     // sequence bytes this line adds to its record
    uint32_t header;      // 1: this line starts a record
    uint32_t irregular;   // 1: the text is outside the regular subset
};

ING_HD uint32_t ing_len_prime(const uint8_t *text, uint32_t s, uint32_t e)
{
    const uint32_t len = e - s;
    return len - (uint32_t)(len > 1u && text[e - 1] == '\r');
}

// k = line index, n_groups4 = number of complete four-line groups (FASTQ mode), final = the text ends
// at the end of the input, prev2_s/prev2_e = the bounds of line k-2 (FASTQ mode, role 3 only)
ING_HD ing_line ing_classify(const uint8_t *text, int mode, int final, uint32_t k, uint32_t s, uint32_t e,
                             uint32_t n_groups4, uint32_t prev2_s, uint32_t prev2_e)
{
    ing_line r;
    r.contrib = 0; r.header = 0; r.irregular = 0;
    const uint32_t len = e - s;
    const uint8_t c0 = len ? text[s] : (uint8_t)'\n';
    const bool lone_cr = len == 1u && c0 == '\r';
    if (mode == ING_MODE_FASTA) {
        if (c0 == '>') { r.header = 1; return r; }
        if (c0 == '@' || c0 == '+' || lone_cr) { r.irregular = 1; return r; }
        r.contrib = ing_len_prime(text, s, e);
        return r;
    }
    if (k >= 4u * n_groups4) {                       // left-over lines: the next block's, unless the input ends here
        r.irregular = (uint32_t)(final && !(len == 0u || lone_cr));
        return r;
    }
    switch (k & 3u) {
    case 0: r.header = 1; r.irregular = (uint32_t)(c0 != '@'); break;
    case 1: r.irregular = (uint32_t)(len == 0u || c0 == '>' || c0 == '+' || c0 == '@');
            r.contrib = r.irregular ? 0u : ing_len_prime(text, s, e); break;
    case 2: r.irregular = (uint32_t)(c0 != '+'); break;
    default: r.irregular = (uint32_t)(ing_len_prime(text, s, e) != ing_len_prime(text, prev2_s, prev2_e)); break;
    }
    return r;
}

// Block-level conditions (no line structure needed).  1 = hand the block to the serial reader:
//   * the first byte is not a header character: the reader would skip junk up to the next '>' / '@'
//     ANYWHERE in the text (:188-191), which is not a line rule;
//   * the input ends with a header character on a line of its own and nothing after it: the name
//     read hits the end of the file at once and no record is produced (:194).
static inline int ing_precheck(const uint8_t *text, uint64_t n, int final)
{
    if (n == 0) return 0;
    if (text[0] != '>' && text[0] != '@') return 1;
    if (final && (text[n - 1] == '>' || text[n - 1] == '@') && (n == 1 || text[n - 2] == '\n')) return 1;
    return 0;
}
