// cornetto_b200/csrc/telofind_core.cuh -- bit-parallel motif matching primitives.
//
// Everything here is plain integer arithmetic on 32-bit words, written so that the same code
// compiles for the device (telofind.cu) and for the host (tests/sim/, which brute-force checks
// these primitives against a byte-wise scan without needing a GPU).
//
// Replaces the per-byte toupper() + strstr()/strncmp() scan of src/find_telomere.c:44-81.
//
// Encoding.  For ASCII A/C/G/T in either case, bits 2..1 of the byte are a 2-bit code:
//     A 0x41 -> 00   C 0x43 -> 01   T 0x54 -> 10   G 0x47 -> 11        (code = (byte >> 1) & 3)
// and bit 5 is the case bit, so no case folding is needed to read the code.  A lane gathers
// those two bits of its 32 bytes into two 32-bit "planes" (bit j = base j).  Bytes that are not
// ACGT alias to some code; a plane match is therefore only a CANDIDATE and every candidate is
// verified byte-exactly (corn_occ_at) before it is reported, which keeps the result identical to
// the reference for arbitrary input bytes.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CORN_HD __host__ __device__ __forceinline__
#else
#define CORN_HD static inline
#endif

#define CORN_MAX_FAST_MOTIF 32   // plane matcher handles motifs over {A,C,G,T} up to this length

CORN_HD uint32_t corn_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        uint32_t s = (sel >> (4 * i)) & 7;
        r |= (uint32_t)((v >> (8 * s)) & 0xFF) << (8 * i);
    }
    return r;
#endif
}

// 64-bit funnel shift right: low 32 bits of ((hi:lo) >> d), 0 <= d < 32
CORN_HD uint32_t corn_funnel_r(uint32_t lo, uint32_t hi, int d)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, d);
#else
    return d ? (lo >> d) | (hi << (32 - d)) : lo;
#endif
}

// Gather bit 1 (plane p1) and bit 2 (plane p2) of 8 consecutive bytes held in two words.
// t keeps bits {1,2} of every byte of wa and, shifted up by four, of wb.  Multiplying the bit-1
// (bit-2) subset by 0x00810204 (0x00408102) moves the eight source bits, which sit 4 apart, to
// the contiguous positions 24..31; all other partial products land on distinct lower bits or
// overflow, so there are no carries (checked exhaustively in tests/sim/test_core.cpp).
CORN_HD void corn_gather8(uint32_t wa, uint32_t wb, uint32_t &y1, uint32_t &y2)
{
    uint32_t t = (wa & 0x06060606u) | ((wb << 4) & 0x60606060u);
    y1 = (t & 0x22222222u) * 0x00810204u;   // byte 3 = plane-1 bits of the 8 bases
    y2 = (t & 0x44444444u) * 0x00408102u;   // byte 3 = plane-2 bits
}

// planes of one 32-byte chunk (w[0] holds bases 0..3, little endian)
CORN_HD void corn_planes32(const uint32_t w[8], uint32_t &p1, uint32_t &p2)
{
    uint32_t a1, a2, b1, b2, c1, c2, d1, d2;
    corn_gather8(w[0], w[1], a1, a2);
    corn_gather8(w[2], w[3], b1, b2);
    corn_gather8(w[4], w[5], c1, c2);
    corn_gather8(w[6], w[7], d1, d2);
    p1 = corn_prmt(corn_prmt(a1, b1, 0x0073), corn_prmt(c1, d1, 0x0073), 0x5410);
    p2 = corn_prmt(corn_prmt(a2, b2, 0x0073), corn_prmt(c2, d2, 0x0073), 0x5410);
}

// 2-bit code of a motif character, or -1 if it is not one of upper-case A/C/G/T.
// (The reference never folds the motif, src/find_telomere.c:90-91, so a lower-case motif
// character can never equal a folded sequence byte.)
CORN_HD int corn_code_of(char c)
{
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'T' ? 2 : c == 'G' ? 3 : -1;
}

// positions whose code equals c
CORN_HD uint32_t corn_class(uint32_t s1, uint32_t s2, unsigned c)
{
    return ((c & 2u) ? s2 : ~s2) & ((c & 1u) ? s1 : ~s1);
}

// Candidate occurrence masks of a motif of length m given as packed 2-bit codes (position d in
// bits 2d+1..2d) for the forward motif (fc) and its reverse complement (rc).  (p1,p2) are the
// planes of this chunk, (n1,n2) those of the following 32 bytes.  Bit j of the result: the
// codes of bytes j..j+m-1 equal the motif's.
template <int M_CT>
CORN_HD void corn_match32(uint32_t p1, uint32_t p2, uint32_t n1, uint32_t n2, uint64_t fc, uint64_t rc, int m_rt,
                          uint32_t &mf, uint32_t &mr)
{
    const int m = M_CT > 0 ? M_CT : m_rt;
    mf = corn_class(p1, p2, (unsigned)(fc & 3));
    mr = corn_class(p1, p2, (unsigned)(rc & 3));
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int d = 1; d < (M_CT > 0 ? M_CT : CORN_MAX_FAST_MOTIF); ++d) {
        if (d >= m) break;
        uint32_t s1 = corn_funnel_r(p1, n1, d), s2 = corn_funnel_r(p2, n2, d);
        mf &= corn_class(s1, s2, (unsigned)((fc >> (2 * d)) & 3));
        mr &= corn_class(s1, s2, (unsigned)((rc >> (2 * d)) & 3));
    }
}

// toupper() in the C locale: only a..z change (src/find_telomere.c:76-81)
CORN_HD uint8_t corn_fold(uint8_t b) { return (uint8_t)(b - ((b >= 'a' && b <= 'z') ? 32 : 0)); }

// exact test "folded bytes p[0..m) equal pat[0..m)"
CORN_HD bool corn_occ_at(const uint8_t *p, const uint8_t *pat, int m)
{
    for (int d = 0; d < m; ++d)
        if (corn_fold(p[d]) != pat[d]) return false;
    return true;
}

// ---- host-side motif analysis (used by telofind.cu on the host, and by the tests) ------------
struct corn_motif_info {
    int      m;                 // length
    int      acgt;              // every character is upper-case A/C/G/T
    int      bordered;          // motif has a proper border (can overlap itself): greedy semantics needed
    int      strands_overlap;   // an occurrence of rc(motif) can overlap one of motif (or they are equal)
    uint64_t fc, rc;            // packed codes (valid when acgt && m <= 32)
    uint8_t  fwd[256], rev[256];
};

static inline int corn_can_overlap(const uint8_t *a, const uint8_t *b, int m)
{
    // some proper suffix of a equals a prefix of b  (b starts inside a)
    for (int k = 1; k < m; ++k) {
        int ok = 1;
        for (int i = 0; i < m - k && ok; ++i) ok = a[k + i] == b[i];
        if (ok) return 1;
    }
    return 0;
}

static inline void corn_analyse_motif(const char *motif, corn_motif_info *mi)
{
    int m = 0;
    while (motif[m] && m < 255) ++m;
    mi->m = m;
    mi->acgt = m > 0;
    mi->fc = mi->rc = 0;
    for (int i = 0; i < m; ++i) {
        mi->fwd[i] = (uint8_t)motif[i];
        char c = motif[m - 1 - i];      // rc(): src/find_telomere.c:24-42 (upper case only)
        mi->rev[i] = (uint8_t)(c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c);
        if (corn_code_of(motif[i]) < 0) mi->acgt = 0;
    }
    mi->fwd[m] = mi->rev[m] = 0;
    if (mi->acgt && m <= CORN_MAX_FAST_MOTIF)
        for (int i = 0; i < m; ++i) {
            mi->fc |= (uint64_t)corn_code_of((char)mi->fwd[i]) << (2 * i);
            mi->rc |= (uint64_t)corn_code_of((char)mi->rev[i]) << (2 * i);
        }
    mi->bordered = corn_can_overlap(mi->fwd, mi->fwd, m);
    int same = 1;
    for (int i = 0; i < m; ++i) same &= mi->fwd[i] == mi->rev[i];
    mi->strands_overlap = same || corn_can_overlap(mi->fwd, mi->rev, m) || corn_can_overlap(mi->rev, mi->fwd, m);
}
