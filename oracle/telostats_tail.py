#!/usr/bin/env python3
"""oracle/telostats_tail.py -- CPU restatement of the tail of scripts/telostats.sh (TEST INFRASTRUCTURE ONLY).

Only tests/ may import or execute this file; the product (`cornetto telostats`, cornetto_b200/host/telostats_main.c)
never does.  It restates, line by line, what /root/reference/scripts/telostats.sh:40-56 does with awk, bedtools, sort
and uniq to the `.windows.0.4` file (written by `cornetto telowin`) and the `.lens` file (written by `cornetto fa2bed`).

PARITY NOTE.  bedtools is not part of the reference's sources and is not installed in this image, so the two bedtools
operations are restated from bedtools' documented behaviour (v2.30 manual) -- parity for this step is pinned to the
RESTATED bedtools semantics, by the hand-checked fixtures in tests/golden/telostats_tail_cases.json, not to a run of
bedtools itself:
  * `bedtools merge -d 100` (telostats.sh:40): input must be grouped by chromosome and sorted by start inside a
    chromosome (telowin's output is: contigs in file order, i ascending).  Consecutive features of the same chromosome
    are merged while next.start - current.end <= 100; the merged end is the maximum end.  Output: chrom, start, end.
  * `bedtools intersect -wa -a A -b B` (telostats.sh:47): every feature of A is written, as it was read, once for every
    feature of B on the same chromosome that overlaps it by >= 1 base (half-open: a.start < b.end and b.start < a.end),
    A's order kept.
Everything else is plain awk/sort/uniq arithmetic on small integers.
"""
import sys


def merge_d(features, d=100):
    """features: list of (chrom, start, end) in file order -> merged list (bedtools merge -d d)."""
    out = []
    for chrom, s, e in features:
        if out and out[-1][0] == chrom and s <= out[-1][2] + d:
            if e > out[-1][2]:
                out[-1][2] = e
        else:
            out.append([chrom, s, e])
    return [tuple(x) for x in out]


def ends_bed(lens, ends=50000):
    """telostats.sh:44 -- awk over the .lens lines (name, length)."""
    out = []
    for name, length in lens:
        if length > ends * 2:
            out.append((name, 0, ends))
            out.append((name, length - ends, length))
        else:
            out.append((name, 0, length))
    return out


def intersect_wa(a, b):
    """bedtools intersect -wa -a a -b b: a's entry once per overlapping b feature."""
    by_chrom = {}
    for chrom, s, e in b:
        by_chrom.setdefault(chrom, []).append((s, e))
    out = []
    for chrom, s, e in a:
        for bs, be in by_chrom.get(chrom, []):
            if s < be and bs < e:
                out.append((chrom, s, e))
    return out


def tally(bed):
    """telostats.sh:56 -- cut -f1 | sort | uniq -c | awk: contigs with 1, 2, more than 2 lines."""
    counts = {}
    for chrom, _, _ in bed:
        counts[chrom] = counts.get(chrom, 0) + 1
    t1 = sum(1 for c in counts.values() if c == 1)
    t2 = sum(1 for c in counts.values() if c == 2)
    t3 = sum(1 for c in counts.values() if c > 2)
    return t1, t2, t3


def bed_text(rows):
    return "".join(f"{c}\t{s}\t{e}\n" for c, s, e in rows)


def run_tail(windows_text: str, lens_text: str, file_arg: str, prefix: str, version: str = "0.2.0"):
    """-> dict of the files the script leaves behind and its stdout."""
    feats = []
    for line in windows_text.splitlines():
        f = line.split()                       # awk '{print $2"\t"$(NF-2)"\t"$(NF-1)}' (telostats.sh:40)
        if f:
            feats.append((f[1], int(f[-3]), int(f[-2])))
    lens = []
    for line in lens_text.splitlines():
        f = line.split()
        if f:
            lens.append((f[0], int(f[-1])))
    merged = merge_d(feats, 100)
    ends = ends_bed(lens, 50000)
    final = intersect_wa(merged, ends)
    t1, t2, t3 = tally(final)
    stdout = (f"cornetto {version}\n" f"genome: {prefix}\nTHRESHOLD: 0.4\nends: 50000\nasm: {file_arg}\n"
              "Merge telomere motifs in 100bp\n\n" "Find those at end of scaffolds, within < 50000\n"
              f"FILE\t{file_arg}\n" f"total telomere regions at the end of contigs:\t{len(final)}\n\n\n"
              f"contigs with 1 telo:\t{t1}\ncontigs with 2 telo:\t{t2}\ncontigs with more than 2 telo:\t{t3}\n\n")
    return {"merged_bed": bed_text(merged), "ends_bed": bed_text(ends), "final_bed": bed_text(final), "stdout": stdout}


if __name__ == "__main__":
    if len(sys.argv) != 5:
        sys.exit("usage: telostats_tail.py <x.windows.0.4> <x.lens> <file argument> <prefix>")
    r = run_tail(open(sys.argv[1]).read(), open(sys.argv[2]).read(), sys.argv[3], sys.argv[4])
    sys.stdout.write(r["stdout"])
