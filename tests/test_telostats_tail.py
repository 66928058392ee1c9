"""The checker's restatement of the tail of scripts/telostats.sh (oracle/telostats_tail.py: bedtools merge -d 100,
contig ends, bedtools intersect -wa, the 1/2/>2 tally) against hand-checked fixtures.  bedtools is not installed and is
not part of the reference's sources, so this step's parity is pinned to the restated bedtools semantics."""
import json
import os
import sys

from util import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))


def test_tail_matches_hand_checked_cases():
    import telostats_tail as tt
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "telostats_tail_cases.json")))
    assert len(cases) >= 4
    for c in cases:
        r = tt.run_tail(c["windows"], c["lens"], "x/asm.fa", "asm")
        assert r["merged_bed"] == tt.bed_text([tuple(x) for x in c["merged"]]), c["name"]
        assert r["ends_bed"] == tt.bed_text([tuple(x) for x in c["ends"]]), c["name"]
        assert r["final_bed"] == tt.bed_text([tuple(x) for x in c["final"]]), c["name"]
        t1, t2, t3 = c["tally"]
        assert r["stdout"].endswith(f"total telomere regions at the end of contigs:\t{len(c['final'])}\n\n\n"
                                    f"contigs with 1 telo:\t{t1}\ncontigs with 2 telo:\t{t2}\ncontigs with more than 2 telo:\t{t3}\n\n"), c["name"]
        assert r["stdout"].startswith("cornetto 0.2.0\ngenome: asm\nTHRESHOLD: 0.4\nends: 50000\nasm: x/asm.fa\nMerge telomere motifs in 100bp\n\n"
                                      "Find those at end of scaffolds, within < 50000\nFILE\tx/asm.fa\n")
