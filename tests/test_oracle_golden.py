"""Oracle vs the committed golden vectors (reference outputs).  CPU only; needs no /root/reference."""
import os

import golden_util
from util import run, write, retab_telomere, lens_from_fa2bed


def test_oracle_matches_golden(oracle_bin, tmp_path):
    g = golden_util.load()
    assert len(g) >= 10
    n_checked = 0
    for name, c in g.items():
        fa = write(str(tmp_path / name), c["input"])
        for m in golden_util.MOTIFS:
            out, _, _ = run([oracle_bin, "telofind", fa, m])
            assert out == c["telofind"][m], (name, m)
            n_checked += 1
        out, _, _ = run([oracle_bin, "fa2bed", fa])
        assert out == c["fa2bed"], name
        tf = write(str(tmp_path / (name + ".telomere")), retab_telomere(c["telofind"]["TTAGGG"]))
        lf = write(str(tmp_path / (name + ".lens")), lens_from_fa2bed(c["fa2bed"]))
        for a in golden_util.TELOWIN:
            out, _, _ = run([oracle_bin, "telowin", tf] + a)
            assert out == c["telowin"][" ".join(a)], (name, a)
        for a in golden_util.SDUST + (golden_util.SDUST_WIDE if name in golden_util.SDUST_WIDE_CASES else []):
            out, _, _ = run([oracle_bin, "sdust"] + a + [fa])
            assert out == c["sdust"][" ".join(a)], (name, a)
        sf = write(str(tmp_path / (name + ".sdust")), c["sdust"][""])
        out, _, _ = run([oracle_bin, "telobreaks", lf, sf, tf])
        assert out == c["telobreaks"], name
    assert n_checked > 50
    # the fixture is not trivially empty
    assert sum(len(c["telofind"]["TTAGGG"]) for c in g.values()) > 5000
    assert sum(len(c["sdust"][""]) for c in g.values()) > 1500
    assert sum(len(c["telobreaks"]) for c in g.values()) > 100
    assert sum(len(c["telowin"]["99.9 0.4"]) for c in g.values()) > 500


def test_oracle_bits_match_golden(oracle_bin, tmp_path):
    """noboringbits / boringbits (src/boringbits_main.c): the oracle's restatement against the reference's outputs for
    several window / threshold / contig-length settings, incl. a window that is not a multiple of the increment, a
    window longer than every contig and a depth above 65535."""
    import synth
    named = synth.depth_arrays(1, synth.BITS_LENGTHS)
    t = write(str(tmp_path / "cov-total.bg"), synth.bedgraph_bytes(named, 1))
    q = write(str(tmp_path / "cov-mq20.bg"), synth.bedgraph_bytes(named, 2))
    want = golden_util.load_bits()
    assert len(want) == 2 * len(synth.BITS_OPTS) and sum(len(v) for v in want.values()) > 50_000
    for key, exp in want.items():
        out, _, _ = run([oracle_bin] + key.split()[:1] + [t, "-q", q] + key.split()[1:])
        assert out == exp, key
