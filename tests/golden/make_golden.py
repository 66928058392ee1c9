#!/usr/bin/env python3
"""Generates tests/golden/golden.json.gz from the UNMODIFIED reference binary
(oracle/_ref/cornetto, compiled from /root/reference by `make -C oracle ref`).

Run in the build container only:   python tests/golden/make_golden.py
The fixture holds, for each case: the input file bytes and the reference's stdout for
telofind (several motifs), fa2bed, telowin (several thresholds), sdust (several -w/-t) and
telobreaks.  It travels with the repository; /root/reference is not needed to USE it.
"""
import base64
import gzip
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from util import retab_telomere, lens_from_fa2bed  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "cornetto")

MOTIFS = ["TTAGGG", "ttaggg", "TATATA", "AAAAAA", "CCCTAA", "TTAGGGTTAGGG", "ACGT", "TTNGGG"]
TELOWIN = [["99.9", "0.4"], ["99.9", "0.1"], ["100"], ["95", "0.05"]]
SDUST = [[], ["-w", "32", "-t", "15"], ["-t", "10"], ["-w", "3"], ["-t", "0"], ["-t", "-3"]]
# windows beyond the tuned kernels' 128 (the generic instance, csrc/sdust_wide.cu); the reference needs O(W^2) and more
# per base inside low-complexity sequence there, so these run on the two smallest sdust inputs only
SDUST_WIDE = [["-w", "129"], ["-w", "200"], ["-w", "500", "-t", "30"]]
SDUST_WIDE_CASES = ("q6_sdust.fa", "sdust_wide.fa")


def ref(args, stdin=None):
    p = subprocess.run([REF] + args, input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True, timeout=300)
    return p.stdout


def b64(b: bytes) -> str:
    return base64.b64encode(b).decode()


def main():
    cases = dict(synth.quirk_corpus())
    cases["asm_small.fa"] = synth.fasta_bytes(
        synth.assembly(42, [150_000, 60_000, 999, 1000, 1200, 7], n_gaps=2, iupac_per_mb=40.0))
    cases["reads_small.fq"] = synth.fastq_bytes(synth.reads(3, 12, n50=8_000, p_telo=0.4))
    cases["sdust_wide.fa"] = synth.fasta_bytes(synth.assembly(77, [5000, 2500, 300], n_gaps=3, telo=(40, 120)))
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for name, data in cases.items():
            fa = os.path.join(td, name)
            open(fa, "wb").write(data)
            c = {"input": b64(data), "telofind": {}, "telowin": {}, "sdust": {}}
            for m in MOTIFS:
                c["telofind"][m] = b64(ref(["telofind", fa, m]))
            telo = retab_telomere(base64.b64decode(c["telofind"]["TTAGGG"]))
            tf = os.path.join(td, name + ".telomere"); open(tf, "wb").write(telo)
            c["fa2bed"] = b64(ref(["fa2bed", fa]))
            lf = os.path.join(td, name + ".lens"); open(lf, "wb").write(lens_from_fa2bed(base64.b64decode(c["fa2bed"])))
            for a in TELOWIN:
                c["telowin"][" ".join(a)] = b64(ref(["telowin", tf] + a))
            for a in SDUST + (SDUST_WIDE if name in SDUST_WIDE_CASES else []):
                c["sdust"][" ".join(a)] = b64(ref(["sdust"] + a + [fa]))
            sf = os.path.join(td, name + ".sdust"); open(sf, "wb").write(base64.b64decode(c["sdust"][""]))
            c["telobreaks"] = b64(ref(["telobreaks", lf, sf, tf]))
            out[name] = c
    # noboringbits / boringbits (src/boringbits_main.c): one pair of bedgraph files, several option sets
    with tempfile.TemporaryDirectory() as td:
        named = synth.depth_arrays(1, synth.BITS_LENGTHS)
        t = os.path.join(td, "cov-total.bg"); open(t, "wb").write(synth.bedgraph_bytes(named, 1))
        q = os.path.join(td, "cov-mq20.bg"); open(q, "wb").write(synth.bedgraph_bytes(named, 2))
        bits = {}
        for cmd in ("noboringbits", "boringbits"):
            for a in synth.BITS_OPTS:
                bits[cmd + " " + " ".join(a)] = b64(ref([cmd, t, "-q", q] + a))
        out["__bits__"] = {"outputs": bits}
    blob = json.dumps(out, sort_keys=True).encode()
    with gzip.GzipFile(os.path.join(HERE, "golden.json.gz"), "wb", mtime=0) as f:
        f.write(blob)
    n_out = sum(len(base64.b64decode(v)) for name, c in out.items() if name != "__bits__" for k in ("telofind", "telowin", "sdust") for v in c[k].values())
    print(f"{len(out)} cases, {n_out} bytes of reference output, fixture {os.path.getsize(os.path.join(HERE, 'golden.json.gz'))} bytes")


if __name__ == "__main__":
    main()
