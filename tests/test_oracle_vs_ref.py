"""Pins the oracle (oracle/oracle.c) against the UNMODIFIED reference binary built by
oracle/Makefile from /root/reference (oracle/_ref/cornetto).  CPU only."""
import os

import numpy as np
import pytest

import synth
from util import run, write, retab_telomere, lens_from_fa2bed


def both(oracle_bin, ref_bin, args, stdin=None):
    a, _, _ = run([oracle_bin] + args, stdin=stdin)
    b, _, _ = run([ref_bin] + args, stdin=stdin)
    return a, b


def pipeline(oracle_bin, ref_bin, tmp_path, fa, tag, motifs=("TTAGGG",), sdust_opts=((),)):
    for motif in motifs:
        a, b = both(oracle_bin, ref_bin, ["telofind", fa] + ([motif] if motif != "TTAGGG" else []))
        assert a == b, f"telofind {tag} motif={motif}"
    telo = write(str(tmp_path / f"{tag}.telomere"), retab_telomere(b))
    bed, _, _ = run([ref_bin, "fa2bed", fa])
    a2, _, _ = run([oracle_bin, "fa2bed", fa])
    assert a2 == bed
    lens = write(str(tmp_path / f"{tag}.lens"), lens_from_fa2bed(bed))
    for extra in (["99.9", "0.4"], ["99.9", "0.1"], ["100"], ["95", "0.05"]):
        a, b = both(oracle_bin, ref_bin, ["telowin", telo] + extra)
        assert a == b, f"telowin {tag} {extra}"
    sd = None
    for opts in sdust_opts:
        a, b = both(oracle_bin, ref_bin, ["sdust"] + list(opts) + [fa])
        assert a == b, f"sdust {tag} {opts}"
        if not opts:
            sd = b
    if sd is not None:
        sdf = write(str(tmp_path / f"{tag}.sdust"), sd)
        a, b = both(oracle_bin, ref_bin, ["telobreaks", lens, sdf, telo])
        assert a == b, f"telobreaks {tag}"
    return True


def test_quirk_corpus(oracle_bin, ref_bin, tmp_path):
    for name, data in synth.quirk_corpus().items():
        fa = write(str(tmp_path / name), data)
        pipeline(oracle_bin, ref_bin, tmp_path, fa, name,
                 motifs=("TTAGGG", "ttaggg", "TATATA", "AAAAAA", "CCCTAA", "TTAGGGTTAGGG", "ACGT", "GGGTTA", "TTNGGG"),
                 sdust_opts=((), ("-w", "32", "-t", "15"), ("-t", "10"), ("-w", "20")))


@pytest.mark.parametrize("seed", [1, 2])
def test_assembly(oracle_bin, ref_bin, tmp_path, seed):
    recs = synth.assembly(seed, [300_000, 120_000, 999, 1000, 1200, 50_000, 7], n_gaps=3, iupac_per_mb=30.0)
    fa = write(str(tmp_path / f"asm{seed}.fa"), synth.fasta_bytes(recs, width=60 if seed == 1 else 80))
    pipeline(oracle_bin, ref_bin, tmp_path, fa, f"asm{seed}", motifs=("TTAGGG", "AAAAAA"))


def test_many_contigs_khash_order(oracle_bin, ref_bin, tmp_path):
    """>16 contigs: khash resizes; telobreaks prints in bucket order (src/telomere_breaks.c:133)."""
    rng = np.random.default_rng(5)
    recs = []
    for i in range(150):
        s = np.concatenate([synth.tandem(rng, b"CCCTAA", 60), synth.random_dna(rng, int(rng.integers(500, 3000))),
                            synth.tandem(rng, b"TTAGGG", 70)])
        recs.append((f"h{i % 7}tg{i * 37 % 1000:06d}l_{'MATERNAL' if i % 2 else 'PATERNAL'}", s))
    fa = write(str(tmp_path / "many.fa"), synth.fasta_bytes(recs))
    pipeline(oracle_bin, ref_bin, tmp_path, fa, "many")


def test_reads_fastq(oracle_bin, ref_bin, tmp_path):
    recs = synth.reads(3, 40, n50=20_000, p_telo=0.3)
    fq = write(str(tmp_path / "reads.fq"), synth.fastq_bytes(recs))
    a, b = both(oracle_bin, ref_bin, ["telofind", fq])
    assert a == b and len(a) > 0
    a, b = both(oracle_bin, ref_bin, ["sdust", fq])
    assert a == b


def test_sdust_stdin(oracle_bin, ref_bin):
    data = synth.fasta_bytes(synth.assembly(9, [20_000], n_gaps=2))
    a, b = both(oracle_bin, ref_bin, ["sdust", "-"], stdin=data)
    assert a == b and len(a) > 0


def test_sdust_n_fuzz(oracle_bin, ref_bin, tmp_path):
    """N-rich fuzz: the stale-window quirk (src/sdust/sdust.c:152-156)."""
    rng = np.random.default_rng(11)
    recs = []
    for k in range(30):
        L = int(rng.integers(200, 4000))
        s = synth.make_contig(rng, L, telo=None, n_its=0, microsat_per_mb=5000.0,
                              n_gaps=int(rng.integers(1, 12)), gap_len=(1, int(rng.choice([3, 60, 400]))), p_lower=0.1)
        recs.append((f"f{k}", s))
    fa = write(str(tmp_path / "nfuzz.fa"), synth.fasta_bytes(recs))
    a, b = both(oracle_bin, ref_bin, ["sdust", fa])
    assert a == b and len(a) > 0
