// tests/sim/ingest_sim.cpp -- CPU walk through the device parser's rules (TEST INFRASTRUCTURE).
//
// Uses the very functions the kernels call (cornetto_b200/csrc/ingest_core.cuh: ing_precheck,
// ing_classify) with plain loops in place of the scans, block by block with the `consumed` carry of
// cornetto_b200/host/ingest.c.  Prints one line per record  name \t length \t hex(sequence)  and, if
// a block is irregular,  IRREGULAR \t <file offset of that block>  and stops -- the test then checks
// that the oracle's parse of the file equals the printed records followed by the oracle's parse of
// the rest of the file from that offset.
//
// usage: ingest_sim <file> <block bytes>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../cornetto_b200/csrc/ingest_core.cuh"

static std::vector<uint8_t> slurp(const char *path)
{
    std::vector<uint8_t> v;
    FILE *f = fopen(path, "rb");
    if (!f) exit(2);
    uint8_t buf[65536];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) v.insert(v.end(), buf, buf + k);
    fclose(f);
    return v;
}

struct Result { bool irregular; uint64_t consumed; };

static Result ingest_block(const uint8_t *text, uint64_t n64, int final)
{
    Result res = { false, 0 };
    if (n64 == 0) return res;
    if (ing_precheck(text, n64, final)) { res.irregular = true; return res; }
    const uint32_t n = (uint32_t)n64;
    const int mode = text[0] == '>' ? ING_MODE_FASTA : ING_MODE_FASTQ;
    std::vector<uint32_t> nl;
    for (uint32_t i = 0; i < n; ++i) if (text[i] == '\n') nl.push_back(i);
    if (final && text[n - 1] != '\n') nl.push_back(n);
    const uint32_t n_lines = (uint32_t)nl.size();
    if (n_lines == 0) return res;
    if ((uint64_t)n_lines * 8u > (uint64_t)n + 4096u) { res.irregular = true; return res; }    // (same resource guard as ingest.cu)
    std::vector<uint32_t> contrib(n_lines), hdr(n_lines);
    for (uint32_t k = 0; k < n_lines; ++k) {
        const uint32_t s = k ? nl[k - 1] + 1 : 0, e = nl[k];
        uint32_t s2 = 0, e2 = 0;
        if (mode == ING_MODE_FASTQ && (k & 3u) == 3u) { s2 = nl[k - 3] + 1; e2 = nl[k - 2]; }
        const ing_line r = ing_classify(text, mode, final, k, s, e, n_lines / 4, s2, e2);
        if (r.irregular) { res.irregular = true; return res; }
        contrib[k] = r.contrib;
        hdr[k] = mode == ING_MODE_FASTA ? r.header : (uint32_t)((k & 3u) == 0 && k < 4 * (n_lines / 4));
    }
    std::vector<uint32_t> hdr_line;
    for (uint32_t k = 0; k < n_lines; ++k) if (hdr[k]) hdr_line.push_back(k);
    const uint32_t n_hdr = (uint32_t)hdr_line.size();
    const uint32_t n_rec = (mode == ING_MODE_FASTA && !final) ? (n_hdr ? n_hdr - 1 : 0) : n_hdr;
    if (n_hdr == 0) { res.consumed = final ? n : 0; return res; }
    // NUL bytes inside records are irregular (the copy kernel's check)
    for (uint32_t r = 0; r < n_rec; ++r) {
        const uint32_t stop = r + 1 < n_hdr ? hdr_line[r + 1] : (mode == ING_MODE_FASTQ ? hdr_line[r] + 4 : n_lines);
        for (uint32_t k = hdr_line[r] + 1; k < stop && k < n_lines; ++k) {
            const uint32_t s = nl[k - 1] + 1;
            for (uint32_t i = 0; i < contrib[k]; ++i) if (text[s + i] == 0) { res.irregular = true; return res; }
        }
    }
    for (uint32_t r = 0; r < n_rec; ++r) {
        const uint32_t h = hdr_line[r];
        const uint32_t hs = h ? nl[h - 1] + 1 : 0;
        uint32_t p = hs + 1;
        while (p < n && !isspace(text[p])) { putchar(text[p]); ++p; }
        const uint32_t stop = r + 1 < n_hdr ? hdr_line[r + 1] : n_lines;
        uint64_t len = 0;
        for (uint32_t k = h + 1; k < stop; ++k) len += contrib[k];
        printf("\t%llu\t", (unsigned long long)len);
        for (uint32_t k = h + 1; k < stop; ++k) {
            const uint32_t s = nl[k - 1] + 1;
            for (uint32_t i = 0; i < contrib[k]; ++i) printf("%02x", text[s + i]);
        }
        putchar('\n');
    }
    if (final) res.consumed = n;
    else if (mode == ING_MODE_FASTA) res.consumed = hdr_line[n_hdr - 1] ? nl[hdr_line[n_hdr - 1] - 1] + 1 : 0;
    else res.consumed = (uint64_t)nl[4 * n_rec - 1] + 1;
    return res;
}

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    const std::vector<uint8_t> file = slurp(argv[1]);
    const uint64_t block = (uint64_t)atoll(argv[2]);
    uint64_t start = 0;                         // file offset of the current block (= carry start)
    uint64_t fresh_end = 0;                     // bytes of the file read so far
    for (;;) {
        const uint64_t carry = fresh_end - start;
        if (carry >= block) { printf("IRREGULAR\t%llu\n", (unsigned long long)start); return 0; }
        uint64_t want = block - carry;
        if (want > file.size() - fresh_end) want = file.size() - fresh_end;
        fresh_end += want;
        const int final = fresh_end >= file.size();
        const uint64_t n = fresh_end - start;
        const Result r = ingest_block(file.data() + start, n, final);
        if (r.irregular || (r.consumed == 0 && !final)) { printf("IRREGULAR\t%llu\n", (unsigned long long)start); return 0; }
        start += r.consumed;
        if (final) break;
    }
    return 0;
}
