"""GPU tests of the device-side FASTA/FASTQ parser (corn_gpu_ingest).

Checker: the oracle's kseq restatement (oracle/oracle.c: orc_parse_fastx_mem).  The parser may decline
text ("irregular"), in which case the CLI falls back to the serial reader -- but whatever it accepts
must come out byte-identical, padding included, and well-formed files must be accepted.
"""
import os

import numpy as np
import pytest

import golden_util
import synth
from test_gpu_parity import BIN, _parse_records, cornetto   # noqa: F401  (checker helpers)
from test_host_cpu import EDGE_FILES, INGEST_EDGE
from util import ROOT, run, write

pytestmark = pytest.mark.gpu

SPACE = b" \t\n\v\f\r"


@pytest.fixture(scope="module")
def capi():
    from cornetto_b200 import capi as m
    from cornetto_b200.build import ensure_built
    ensure_built()
    m.load()
    return m


@pytest.fixture(scope="module")
def ctx(capi):
    c = capi.Context(0)
    yield c
    c.close()


def records_of(res, text: bytes):
    out = []
    for off, ln, seq in zip(res["hdr_off"], res["length"], res["seq"]):
        p = int(off) + 1
        q = p
        while q < len(text) and text[q] not in SPACE:
            q += 1
        assert len(seq) == int(ln)
        out.append((text[p:q], seq))
    return out


MUST_BE_REGULAR = {"asm_small.fa", "q1_ends.fa", "q2_case.fa", "q3_adjacent.fa", "q4_crlf.fa", "q4_single.fq", "q5_winsizes.fa",
                   "q6_sdust.fa", "reads_small.fq", "win.fq", "fq_at_in_qual.fq", "fa_crlf.fa", "fa_blank_lines.fa", "fa_no_final_nl.fa",
                   "fa_ws_name.fa", "fa_inner_cr.fa", "fq_no_final_nl.fq", "fq_trailing_blank.fq", "fa_long_line.fa"}


def all_cases():
    cases = {k: c["input"] for k, c in golden_util.load().items()}
    cases.update(EDGE_FILES)
    cases.update(INGEST_EDGE)
    return cases


def test_ingest_whole_files(ctx):
    n_reg = 0
    for name, data in all_cases().items():
        res = ctx.ingest(data, final=True)
        if res["irregular"]:
            assert name not in MUST_BE_REGULAR, name
            assert res["n_rec"] == 0 and res["consumed"] == 0
            continue
        n_reg += 1
        assert res["consumed"] == len(data), name
        assert records_of(res, data) == _parse_records(data), name
    assert n_reg >= len(MUST_BE_REGULAR)


def test_ingest_block_seams(ctx):
    """Non-final blocks: the records before `consumed` plus the parse of the rest equal the whole parse,
    wherever the block ends (inside a header, a sequence line, a quality line, between records)."""
    rng = np.random.default_rng(3)
    for name in ("q3_adjacent.fa", "q4_crlf.fa", "q4_single.fq", "q5_winsizes.fa", "win.fq", "fa_blank_lines.fa", "fa_long_line.fa"):
        data = all_cases()[name]
        want = _parse_records(data)
        cuts = sorted(set(int(x) for x in rng.integers(1, len(data), 12)) | {1, 2, len(data) - 1})
        for cut in cuts:
            a = ctx.ingest(data[:cut], final=False)
            assert not a["irregular"], (name, cut)
            assert a["consumed"] <= cut
            rest = data[a["consumed"]:]
            b = ctx.ingest(rest, final=True)
            assert not b["irregular"], (name, cut)
            assert records_of(a, data[:cut]) + records_of(b, rest) == want, (name, cut)


def test_ingest_synthetic_assembly_and_reads(ctx):
    """MB-scale inputs: 60- and 80-column FASTA with soft-masked and IUPAC bytes, single-line FASTA, FASTQ reads."""
    recs = synth.assembly(21, [3_000_000, 1_200_000, 70_000, 999, 31, 32, 33, 0, 1], n_gaps=3, iupac_per_mb=50.0, p_lower=0.05)
    for width in (60, 80, 0):
        data = synth.fasta_bytes(recs, width=width)
        res = ctx.ingest(np.frombuffer(data, dtype=np.uint8), final=True, pin=True)
        assert not res["irregular"]
        assert records_of(res, data) == _parse_records(data), width
    rd = synth.reads(22, 4000, n50=1200)
    data = synth.fastq_bytes(rd)
    res = ctx.ingest(data, final=True)
    assert not res["irregular"] and res["n_rec"] == len(rd)
    assert records_of(res, data) == _parse_records(data)
    t = ctx.timing()
    assert t["launches"] >= 5 and t["scan_ms"] > 0


def test_ingest_large_pageable_text_uses_staged_copy(ctx):
    """70 MB of text in ordinary (pageable) memory: the host->device copy goes through the library's page-locked
    staging ring (corn_h2d) instead of one cudaMemcpyAsync; the parsed records must be the bytes that went in."""
    rng = np.random.default_rng(41)
    lens = [30_000_001, 25_000_000, 14_999_999, 64, 0]
    seqs = [np.frombuffer(b"ACGTNacgt", dtype=np.uint8)[rng.integers(0, 9, size=n)] for n in lens]
    for width in (0, 61):
        data = synth.fasta_bytes([(f"big{i} w{width}", s) for i, s in enumerate(seqs)], width=width)
        assert len(data) > 4 * (8 << 20)                    # above the staging threshold
        res = ctx.ingest(np.frombuffer(data, dtype=np.uint8).copy(), final=True)     # a fresh pageable array
        assert not res["irregular"] and [int(x) for x in res["length"]] == lens
        for got, want in zip(res["seq"], seqs):
            assert got == want.tobytes()
        assert [n for n, _ in records_of(res, data)] == [b"big%d" % i for i in range(len(lens))]


def test_ingest_feeds_resident_scans(ctx, capi):
    """The resident batch the parser leaves behind gives the same telofind / sdust results as the host-batch path."""
    recs = synth.assembly(23, [400_000, 90_000, 5_000], n_gaps=2, iupac_per_mb=20.0, p_lower=0.05)
    data = synth.fasta_bytes(recs, width=60)
    res = ctx.ingest(data, final=True, keep_db=True)
    assert not res["irregular"]
    hb = capi.HostBatch([s for _, s in _parse_records(data)])
    want_runs = ctx.telofind(hb, "TTAGGG")
    want_iv, want_first = ctx.sdust(hb)
    got_runs = ctx.telofind_dev(res["db"], "TTAGGG")
    got_iv, got_first = ctx.sdust_dev(res["db"])
    assert np.array_equal(got_runs, want_runs)
    assert np.array_equal(got_iv, want_iv) and np.array_equal(got_first, want_first)
    ctx.free(res["db"])
    hb.close()


def test_cli_ingest_equals_serial_reader(tmp_path):
    """The binary with the device parser (default), with tiny blocks (many seams + fallbacks) and with the
    serial reader only ($CORNETTO_INGEST=0) prints the same bytes; irregular files fall back mid-file."""
    recs = synth.assembly(24, [300_000, 120_000, 40_000, 999, 1000, 7], n_gaps=3, iupac_per_mb=100.0, p_lower=0.05)
    fa = write(str(tmp_path / "a.fa"), synth.fasta_bytes(recs, width=70))
    fq = write(str(tmp_path / "r.fq"), synth.fastq_bytes(synth.reads(25, 3000, n50=900)))
    # regular records, then an irregular tail (multi-line FASTQ record, '+' line inside FASTA)
    mixed = write(str(tmp_path / "m.fa"), synth.fasta_bytes(recs[:3], width=60) + b">odd\nACGT\n+\nIIII\n>after\nTTAGGGTTAGGGTTAGGG\n")
    for path in (fa, fq, mixed):
        for cmd in (["telofind", path], ["sdust", path], ["sdust", "-w", "20", "-t", "12", path]):
            want, _, _ = cornetto(cmd, env={"CORNETTO_INGEST": "0"})
            got, _, _ = cornetto(cmd)
            assert got == want, cmd
            for bb in ("100000", "40000"):
                got, _, _ = cornetto(cmd, env={"CORNETTO_BATCH_BYTES": bb})
                assert got == want, (cmd, bb)
    assert len(want) > 0
