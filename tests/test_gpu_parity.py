"""GPU parity tests proper (run on the B200 box with `-m gpu`).

Everything here goes through the product: either the drop-in `cornetto` binary (host C +
C ABI + CUDA) or the C ABI itself via ctypes.  The checkers are the committed golden vectors
(reference outputs, tests/golden/) and the oracle (oracle/, CPU restatement).  Integer/byte work:
the bar is bit-exact.
"""
import os
import sys
import subprocess

import numpy as np
import pytest

import golden_util
import synth
from util import ROOT, run, write, retab_telomere, lens_from_fa2bed

pytestmark = pytest.mark.gpu

BIN = os.path.join(ROOT, "cornetto_b200", "bin", "cornetto")


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


def cornetto(args, stdin=None, env=None, check=True):
    e = dict(os.environ)
    if env:
        e.update(env)
    return run([BIN] + args, stdin=stdin, env=e, check=check)


# ---------------------------------------------------------------------------------------------
# 1. drop-in binary vs the golden vectors made by the unmodified reference
# ---------------------------------------------------------------------------------------------
def test_cli_matches_golden(tmp_path):
    """Every golden case through the binary: default motif + one rotating extra motif, two telowin
    settings, two sdust settings, telobreaks, fa2bed.  (The full motif x option matrix runs
    in-process through the C ABI in test_abi_matches_golden, which avoids ~200 CUDA start-ups.)"""
    g = golden_util.load()
    for k, (name, c) in enumerate(sorted(g.items())):
        fa = write(str(tmp_path / name), c["input"])
        out, _, _ = cornetto(["telofind", fa])
        assert out == c["telofind"]["TTAGGG"], ("telofind default motif", name)
        m = golden_util.MOTIFS[1 + k % (len(golden_util.MOTIFS) - 1)]
        out, _, _ = cornetto(["telofind", fa, m])
        assert out == c["telofind"][m], ("telofind", name, m)
        out, _, _ = cornetto(["fa2bed", fa])
        assert out == c["fa2bed"], ("fa2bed", name)
        tf = write(str(tmp_path / (name + ".telomere")), retab_telomere(c["telofind"]["TTAGGG"]))
        lf = write(str(tmp_path / (name + ".lens")), lens_from_fa2bed(c["fa2bed"]))
        for a in (golden_util.TELOWIN[0], golden_util.TELOWIN[1 + k % 3]):
            out, _, _ = cornetto(["telowin", tf] + a)
            assert out == c["telowin"][" ".join(a)], ("telowin", name, a)
        for a in [golden_util.SDUST[0], golden_util.SDUST[1 + k % (len(golden_util.SDUST) - 1)]] + (golden_util.SDUST_WIDE[1:2] if name in golden_util.SDUST_WIDE_CASES else []):
            out, _, _ = cornetto(["sdust"] + a + [fa])
            assert out == c["sdust"][" ".join(a)], ("sdust", name, a)
        sf = write(str(tmp_path / (name + ".sdust")), c["sdust"][""])
        out, _, _ = cornetto(["telobreaks", lf, sf, tf])
        assert out == c["telobreaks"], ("telobreaks", name)


def _parse_records(data: bytes):
    """(name, bytes) records exactly as kseq delivers them, via the oracle's reader (checker side)."""
    import ctypes as C
    path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    L = C.CDLL(path)

    class Rec(C.Structure):
        _fields_ = [("name", C.c_char_p), ("seq", C.POINTER(C.c_ubyte)), ("len", C.c_size_t)]
    L.orc_parse_fastx_mem.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(Rec)), C.POINTER(C.c_size_t)]
    recs, n = C.POINTER(Rec)(), C.c_size_t()
    import gzip
    if data[:2] == b"\x1f\x8b":
        data = gzip.decompress(data)
    L.orc_parse_fastx_mem(data, len(data), C.byref(recs), C.byref(n))
    return [(recs[i].name, bytes(recs[i].seq[:recs[i].len])) for i in range(n.value)]


def test_abi_matches_golden(ctx, capi):
    """Full motif / threshold / (T,W) matrix of the golden fixture through the C ABI, in-process."""
    g = golden_util.load()
    for name, c in g.items():
        recs = _parse_records(c["input"])
        hb = capi.HostBatch([s for _, s in recs])
        for m in golden_util.MOTIFS:
            runs = ctx.telofind(hb, m)
            got = b"".join(b"%s\t%d\t%d\t%d\t%d\t%d\n" % (recs[r["rec"]][0], len(recs[r["rec"]][1]), r["strand"], r["start"], r["end"], r["end"] - r["start"]) for r in runs)
            assert got == c["telofind"][m], ("telofind", name, m)
        for a in golden_util.SDUST + (golden_util.SDUST_WIDE if name in golden_util.SDUST_WIDE_CASES else []):
            T, W = 20, 64
            if "-t" in a:
                T = int(a[a.index("-t") + 1])
            if "-w" in a:
                W = int(a[a.index("-w") + 1])
            iv, first = ctx.sdust(hb, T, W)
            got = b"".join(b"%s\t%d\t%d\n" % (recs[r][0], int(v >> np.uint64(32)), np.int32(np.uint32(v & np.uint64(0xFFFFFFFF))))
                           for r in range(len(recs)) for v in iv[int(first[r]):int(first[r + 1])])
            assert got == c["sdust"][" ".join(a)], ("sdust", name, a)
        hb.close()


def test_cli_small_batches_and_chunks(tmp_path, oracle_bin):
    """Record batches of a few KiB and 64..193-base sdust chunks: every seam type is crossed."""
    recs = synth.assembly(11, [40_000, 9_000, 999, 1000, 1200, 7, 3_000], n_gaps=4, iupac_per_mb=200.0,
                          microsat_per_mb=2000.0, telo=(20, 200))
    fa = write(str(tmp_path / "a.fa"), synth.fasta_bytes(recs))
    want_t, _, _ = run([oracle_bin, "telofind", fa])
    want_s, _, _ = run([oracle_bin, "sdust", fa])
    for bb in ("4096", "20000", "70000"):
        out, _, _ = cornetto(["telofind", fa], env={"CORNETTO_BATCH_BYTES": bb})
        assert out == want_t, bb
        for ch in ("64", "101", "193", "1024"):
            out, _, _ = cornetto(["sdust", fa], env={"CORNETTO_BATCH_BYTES": bb, "CORNETTO_SDUST_CHUNK": ch})
            assert out == want_s, (bb, ch)


def test_cli_stdin_gz_and_options(tmp_path, oracle_bin):
    recs = synth.assembly(12, [30_000, 5_000], n_gaps=1)
    data = synth.fasta_bytes(recs)
    fa = write(str(tmp_path / "b.fa"), data)
    want, _, _ = run([oracle_bin, "sdust", fa])
    out, _, _ = cornetto(["sdust", "-"], stdin=data)
    assert out == want
    want2, _, _ = run([oracle_bin, "sdust", "-w", "32", "-t", "15", fa])
    out, _, _ = cornetto(["sdust", fa, "-w", "32", "-t", "15"])      # options after the operand (ketopt permutes)
    assert out == want2
    out, _, _ = cornetto(["sdust", "-w32", "-t15", fa])
    assert out == want2
    import gzip
    gz = write(str(tmp_path / "b.fa.gz"), gzip.compress(data))
    want_t, _, _ = run([oracle_bin, "telofind", fa])
    out, _, _ = cornetto(["telofind", gz])
    assert out == want_t


def test_cli_errors(tmp_path):
    _, err, rc = cornetto(["telofind"], check=False)
    assert rc == 1 and b"Usage: find <input fasta>" in err
    _, err, rc = cornetto(["telofind", str(tmp_path / "nope.fa")], check=False)
    assert rc == 1 and b"Could not to open file" in err
    _, err, rc = cornetto(["sdust"], check=False)
    assert rc == 1 and err.startswith(b"Usage: sdust [-w 64] [-t 20] <in.fa>")
    _, err, rc = cornetto(["telowin", "x"], check=False)
    assert rc == 1 and b"Usage: cornetto telowin" in err
    out, err, rc = cornetto(["--version"])
    assert out == b"cornetto 0.2.0\n"


# ---------------------------------------------------------------------------------------------
# 2. C ABI vs the oracle on seeded inputs
# ---------------------------------------------------------------------------------------------
def _oracle_lib():
    import ctypes as C
    path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    L = C.CDLL(path)

    class Run(C.Structure):
        _fields_ = [("strand", C.c_uint32), ("start", C.c_uint64), ("end", C.c_uint64)]
    L.orc_telofind.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.POINTER(C.POINTER(Run))]
    L.orc_telofind.restype = C.c_size_t
    L.orc_sdust.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.orc_sdust.restype = C.POINTER(C.c_uint64)
    return L, Run


def oracle_telofind(records, motif):
    import ctypes as C
    L, Run = _oracle_lib()
    out = []
    for i, r in enumerate(records):
        p = C.POINTER(Run)()
        b = bytes(r)
        n = L.orc_telofind(b, len(b), motif.encode(), C.byref(p))
        out += [(i, p[k].strand, p[k].start, p[k].end) for k in range(n)]
    return np.array(out, dtype=np.uint64).reshape(-1, 4)


def oracle_sdust(records, T, W):
    import ctypes as C
    L, _ = _oracle_lib()
    ivs, first = [], [0]
    for r in records:
        b = bytes(r)
        n = C.c_int()
        p = L.orc_sdust(b, len(b), T, W, C.byref(n))
        ivs += [p[k] for k in range(n.value)]
        first.append(len(ivs))
    return np.array(ivs, dtype=np.uint64), np.array(first, dtype=np.uint64)


def as_rows(runs):
    return np.stack([runs["rec"], runs["strand"], runs["start"], runs["end"]], axis=1).astype(np.uint64).reshape(-1, 4)


@pytest.mark.parametrize("motif", ["TTAGGG", "TTTAGGG", "AAAAAA", "TATATA", "TTNGGG", "ACGTACGTACGTACGTACGTACGTACGTACGTACGT",
                                   "A", "GT", "CCCTAA", "GATTACAGATTACAGATTACAGATTACAGATT", "GATTACAGATTACAGATTACAGATTACAGATTA", "ttaggg", "TTAGGg"])
def test_abi_telofind(ctx, capi, motif):
    rng = np.random.default_rng(21)
    recs = [synth.make_contig(rng, int(L), n_gaps=2, iupac_per_mb=300.0, telo=(30, 300)) for L in (70_000, 33, 0, 5, 6, 120_001, 1024, 31, 32, 64)]
    recs.append(np.frombuffer(b"TTAGGG" * 2000, dtype=np.uint8))           # one run across many tiles' worth of chunks
    recs.append(np.frombuffer((b"TTAGGGA" * 3000), dtype=np.uint8))        # dense start/end events
    recs.append(np.frombuffer(b"GATTACAGATTACAGATTACAGATTACAGATTA" * 40 + b"gattacagattacagattacagattacagatt" * 3, dtype=np.uint8))
    recs.append(np.frombuffer(b"GT" * 500 + b"AAAAAAAAAAAAA" + b"TG" * 77, dtype=np.uint8))
    hb = capi.HostBatch(recs)
    got = as_rows(ctx.telofind(hb, motif))
    want = oracle_telofind(recs, motif)
    assert got.shape == want.shape and (got == want).all()


def test_abi_telofind_telowin_fused(ctx, capi, oracle_bin, tmp_path):
    recs = synth.assembly(31, [400_000, 150_000, 999, 1000, 1200, 1400, 2000, 1, 0], telo=(200, 900))
    hb = capi.HostBatch([s for _, s in recs])
    runs = ctx.telofind(hb, "TTAGGG")
    for thr_in, ident in ((0.4, 99.9), (0.1, 99.9), (0.05, 95.0), (0.0, 100.0)):
        thr = thr_in * (ident / 100) ** 6
        fused = ctx.telowin(thr)
        general = ctx.telowin(thr, runs=runs, lengths=hb.lengths)
        assert (fused == general).all() and len(fused) == len(general)
        # text check against the oracle CLI
        tsv = b"".join(b"%s\t%d\t%d\t%d\t%d\t%d\n" % (recs[r["rec"]][0].encode(), hb.lengths[r["rec"]], r["strand"], r["start"], r["end"], r["end"] - r["start"]) for r in runs)
        tf = write(str(tmp_path / "f.telomere"), tsv)
        want, _, _ = run([oracle_bin, "telowin", tf, repr(ident), repr(thr_in)])
        got = b"".join(b"Window\t%s\t%d\t%d\t%d\t%s\n" % (recs[w["rec"]][0].encode(), hb.lengths[w["rec"]], w["start"], w["end"],
                                                       (b"%.3g" % (float(w["car"]) / float(int(w["end"]) - int(w["start"])))))
                       for w in fused if True)
        # contigs without any run never reach the reference's text path (scripts/telostats.sh pipes telofind output)
        names_with_runs = {recs[r["rec"]][0].encode() for r in runs}
        got = b"".join(l + b"\n" for l in got.splitlines() if l.split(b"\t")[1] in names_with_runs)
        assert got == want, (thr_in, ident)


@pytest.mark.parametrize("tw", [(20, 64), (15, 32), (10, 64), (25, 100), (8, 20)])
def test_abi_sdust(ctx, capi, tw):
    T, W = tw
    rng = np.random.default_rng(41)
    recs = [synth.make_contig(rng, int(L), telo=None, n_its=0, microsat_per_mb=3000.0, n_gaps=g, gap_len=(1, gl), p_lower=0.1, iupac_per_mb=500.0)
            for L, g, gl in ((90_000, 6, 400), (5000, 30, 3), (4097, 0, 1), (4096, 1, 60), (1, 0, 1), (0, 0, 1), (2, 0, 1), (130_000, 40, 60))]
    recs.append(np.frombuffer(b"A" * 100 + b"N" + b"AAAAA", dtype=np.uint8))   # interval past the end of the record
    recs.append(np.frombuffer(b"A" * 20_000, dtype=np.uint8))                  # one interval across many chunks
    hb = capi.HostBatch(recs)
    iv, first = ctx.sdust(hb, T, W)
    wiv, wfirst = oracle_sdust(recs, T, W)
    assert (first == wfirst).all()
    assert len(iv) == len(wiv) and (iv == wiv).all()


def test_cli_telostats_matches_script(tmp_path, oracle_bin):
    """`cornetto telostats asm.fa` = scripts/telostats.sh in one pass: every file the script leaves behind and its
    stdout, against the oracle's telofind / fa2bed / telowin and the restated tail (oracle/telostats_tail.py)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import telostats_tail as tt
    recs = synth.assembly(91, [260_000, 140_000, 100_001, 90_000, 60_000, 900, 130_000], n_gaps=1, telo=(150, 700), n_its=2)
    # an interstitial telomere block in the middle of the first contig (must NOT reach the ends bed) and one just inside
    # the last 50 kb of the second (must)
    blk = np.frombuffer(b"TTAGGG" * 400, dtype=np.uint8)
    recs[0][1][120_000:120_000 + len(blk)] = blk
    recs[1][1][95_000:95_000 + len(blk)] = blk
    for env in ({}, {"CORNETTO_BATCH_BYTES": "300000"}, {"CORNETTO_INGEST": "0"}):
        wd = tmp_path / ("w%d" % len(env) + "".join(env))
        wd.mkdir()
        fa = write(str(wd / "asm.fa"), synth.fasta_bytes(recs))
        e = dict(os.environ)
        e.update(env)
        p = subprocess.run([BIN, "telostats", "asm.fa"], cwd=str(wd), env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert p.returncode == 0, p.stderr[-2000:]
        want_t, _, _ = run([oracle_bin, "telofind", fa])
        want_l = lens_from_fa2bed(run([oracle_bin, "fa2bed", fa])[0])
        tf = write(str(wd / "want.telomere"), retab_telomere(want_t))
        want_w, _, _ = run([oracle_bin, "telowin", tf, "99.9", "0.4"])
        tail = tt.run_tail(want_w.decode(), want_l.decode(), "asm.fa", "asm")
        t = wd / "tmp_asm_telostats"
        assert (t / "asm.telomere").read_bytes() == retab_telomere(want_t)
        assert (t / "asm.lens").read_bytes() == want_l
        assert (t / "asm.windows.0.4").read_bytes() == want_w and len(want_w) > 200
        assert (t / "asm.windows.0.4.bed").read_text() == tail["merged_bed"]
        assert (t / "asm.ends.bed").read_text() == tail["ends_bed"]
        assert (wd / "asm.windows.0.4.50kb.ends.bed").read_text() == tail["final_bed"]
        assert p.stdout.decode() == tail["stdout"]
        assert tail["final_bed"].count("\n") >= 10 and "contig_1\t120" not in tail["final_bed"]
    _, _, rc = cornetto(["telostats"], check=False)
    assert rc == 1


def test_cli_gzip_inputs_through_the_device_parser(tmp_path, oracle_bin):
    """.gz inputs are inflated by host/gzsrc.c (BGZF members in parallel) and parsed on the device; results must equal the
    oracle's, which reads through gzread: one member, several members, BGZF, small text blocks (carry-over between
    blocks), FASTQ, and with the serial reader forced."""
    import gzip
    fa = synth.fasta_bytes(synth.assembly(19, [300_000, 120_000, 40_000, 7], n_gaps=1, telo=(60, 300)))
    fq = synth.fastq_bytes(synth.reads(4, 20, n50=9_000, p_telo=0.4))
    files = {"a.fa.gz": gzip.compress(fa), "b.fa.gz": synth.bgzf_bytes(fa), "c.fq.gz": synth.bgzf_bytes(fq),
             "m.fa.gz": gzip.compress(fa[:100_000]) + gzip.compress(fa[100_000:])}
    for name, data in files.items():
        p = write(str(tmp_path / name), data)
        want_t, _, _ = run([oracle_bin, "telofind", p])
        want_s, _, _ = run([oracle_bin, "sdust", p])
        assert len(want_t) > 500
        for env in ({}, {"CORNETTO_BATCH_BYTES": "150000"}, {"CORNETTO_GZ_INGEST": "0"}):
            out, _, _ = cornetto(["telofind", p], env=env)
            assert out == want_t, (name, env)
            out, _, _ = cornetto(["sdust", p], env=env)
            assert out == want_s, (name, env)


def test_cli_noboringbits_matches_golden(tmp_path):
    """`cornetto noboringbits` / `boringbits` through the binary against the reference's outputs (golden), every option set."""
    named = synth.depth_arrays(1, synth.BITS_LENGTHS)
    t = write(str(tmp_path / "cov-total.bg"), synth.bedgraph_bytes(named, 1))
    q = write(str(tmp_path / "cov-mq20.bg"), synth.bedgraph_bytes(named, 2))
    for key, exp in golden_util.load_bits().items():
        out, err, _ = cornetto(key.split()[:1] + [t, "-q", q] + key.split()[1:])
        assert out == exp, key
        assert b"Number of contigs: 6" in err and b"truncated to 65535" in err
    # argument and input errors as in the reference: usage on stderr, exit 1; files in a different order
    _, err, rc = cornetto(["noboringbits", t], check=False)
    assert rc == 1 and err.startswith(b"Usage: cornetto boringbits cov-total.bg -q cov-mq20.bg")
    out, _, rc = cornetto(["noboringbits", "-h"], check=False)
    assert rc == 0 and out.startswith(b"Usage: cornetto boringbits")
    bad = write(str(tmp_path / "bad.bg"), synth.bedgraph_bytes(named[1:], 2))
    _, err, rc = cornetto(["noboringbits", t, "-q", bad], check=False)
    assert rc == 1 and b"The two files are not in the same order" in err
    # the block-parallel text reader (host/depthtxt.c) declines these files (one depth above 65535) and the one-pass
    # reader takes over: same output, same warning
    par = {"CORNETTO_DEPTH_PAR_MIN": "0", "CORNETTO_DEPTH_BLOCK": "3000"}
    key, exp = next(iter(golden_util.load_bits().items()))
    out, err, _ = cornetto(key.split()[:1] + [t, "-q", q] + key.split()[1:], env=par)
    assert out == exp and b"truncated to 65535" in err
    _, err, rc = cornetto(["noboringbits", t, "-q", bad], check=False, env=par)
    assert rc == 1 and b"The two files are not in the same order" in err


def test_cli_noboringbits_parallel_reader(tmp_path, oracle_bin):
    """Plain depth tables go through the block-parallel reader (forced here for small files, blocks of 3000 bytes and of
    1 MB, 1 and 5 threads): stdout identical to the oracle's for every option set, both commands."""
    named = [(nm, np.minimum(d, 65535), np.minimum(m, 65535)) for nm, d, m in synth.depth_arrays(21, [40000, 12000, 900, 2500, 2549, 7, 25000])]
    t = write(str(tmp_path / "cov-total.bg"), synth.bedgraph_bytes(named, 1))
    q = write(str(tmp_path / "cov-mq20.bg"), synth.bedgraph_bytes(named, 2))
    for cmd in ("noboringbits", "boringbits"):
        for k, opts in enumerate(synth.BITS_OPTS):
            want, _, _ = run([oracle_bin, cmd, t, "-q", q] + opts)
            block, threads = ("3000", "5") if k % 2 == 0 else ("1048576", "1")
            out, err, _ = cornetto([cmd, t, "-q", q, "-t", threads] + opts, env={"CORNETTO_DEPTH_PAR_MIN": "0", "CORNETTO_DEPTH_BLOCK": block})
            assert out == want, (cmd, opts)
            assert b"Number of contigs: 7" in err and b"truncated" not in err


def test_abi_depthwin(ctx, capi):
    """corn_gpu_depthwin against a direct numpy restatement of get_regs() + the selection rules: random depths with zeros
    (division by zero in the mapq ratio), windows that are / are not multiples of the increment, windows spanning
    thousands of increments (the direct kernel), contigs shorter than a window, many tiles per contig."""
    rng = np.random.default_rng(77)
    lens = [200_000, 33_333, 2500, 2499, 51, 1, 1_300_000]
    d = [rng.poisson(20, size=L).astype(np.uint16) for L in lens]
    q = [(x * rng.uniform(0, 1, size=len(x))).astype(np.uint16) for x in d]
    d[0][5000:9000] = 0
    q[0][5000:7000] = 0
    d[6][400_000:400_100] = 65535
    for (w, inc, lo, hi, thr, edge, minlen, boring) in ((2500, 50, 8, 50, 0.4, 1000, 30000, 0), (2500, 50, 8, 50, 0.4, 1000, 30000, 1),
                                                        (777, 50, 12, 30, 0.6, 0, 0, 0), (100, 7, 15, 25, 0.5, 10, 50, 1), (20000, 7, 10, 40, 0.5, 0, 2000, 0),
                                                        (50, 50, 19, 21, 0.45, 100, 2500, 0), (64, 8, 18, 22, 0.5, 0, 0, 0), (2501, 13, 19, 21, 0.5, 50, 60, 1),
                                                        (333, 16, 19, 21, 0.5, 0, 0, 0), (1000, 17, 18, 22, 0.45, 20, 40, 1)):
        got = ctx.depthwin(d, q, w, inc, lo, hi, thr, edge, minlen, boring)
        want = []
        for c, (dd, qq) in enumerate(zip(d, q)):
            L = len(dd)
            n_reg = max(1, int((L - w + inc - 1) / inc) + 1)          # C division truncates toward zero
            cd = np.concatenate([[0], np.cumsum(dd, dtype=np.int64)]); cq = np.concatenate([[0], np.cumsum(qq, dtype=np.int64)])
            st = np.arange(n_reg, dtype=np.int64) * inc
            en = np.minimum(st + w, L)
            dep = (cd[en] - cd[st]) // (en - st); mq = (cq[en] - cq[st]) // (en - st)
            with np.errstate(divide="ignore", invalid="ignore"):
                fun = (dep < lo) | (dep > hi) | ((mq / dep.astype(np.float64)) < np.float64(np.float32(thr)))
            sel = (fun & (L >= minlen)) if not boring else ((L > minlen) & (st > edge) & (en < L - edge) & ~fun)
            for k in np.flatnonzero(sel):
                want.append((c, int(st[k]), int(en[k]), int(dep[k]), int(mq[k])))
        assert [tuple(int(x) for x in r) for r in got] == want, (w, inc, boring)
    with pytest.raises(capi.CornError):
        ctx.depthwin(d, q, 0, 50)


def test_sdust_library_api(capi):
    """sdust() / sdust_buf_init / sdust_core / sdust_buf_destroy with the reference's signatures and ownership rules
    (src/sdust/sdust.h:16-21): l_seq < 0 means strlen, sdust()'s result is free()d by the caller, sdust_core()'s
    belongs to the buffer and is replaced by the next call."""
    import ctypes as C
    L = capi.load()
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(9)
    seqs = [synth.make_contig(rng, n, telo=None, n_its=0, microsat_per_mb=4000.0, n_gaps=g, gap_len=(1, 30)).tobytes()
            for n, g in ((20_000, 3), (700, 1), (5, 0))] + [b"A" * 100 + b"N" + b"AAAAA", b""]
    buf = L.sdust_buf_init(None)
    assert buf
    for T, W in ((20, 64), (12, 30), (20, 200)):
        for sq in seqs:
            want, _ = oracle_sdust([np.frombuffer(sq, dtype=np.uint8)], T, W)
            n = C.c_int(-5)
            p = L.sdust(None, sq, -1 if (len(sq) and 0 not in sq) else len(sq), T, W, C.byref(n))
            assert bool(p) and n.value == len(want)
            assert [int(p[i]) for i in range(n.value)] == [int(x) for x in want]
            libc.free(C.cast(p, C.c_void_p))
            n2 = C.c_int(-5)
            q = L.sdust_core(sq, len(sq), T, W, C.byref(n2), buf)
            assert bool(q) and n2.value == len(want) and [int(q[i]) for i in range(n2.value)] == [int(x) for x in want]
    L.sdust_buf_destroy(buf)
    L.sdust_buf_destroy(None)


def test_abi_sdust_stale_window_after_gaps(ctx, capi):
    """Low-complexity sequence right after runs of N: perfect intervals inserted while the window start is still pinned
    to the first base after the gap live up to 2W steps (src/sdust/sdust.c:146); the two-phase path must keep those
    positions in its items.  (bench.py's in-run check found a lost 7-base interval 57 bases after a 1.6 kb gap.)"""
    for seed in (9100, 9103, 9107):
        recs = [s for _, s in synth.stale_window_records(seed, n_rec=8)]
        hb = capi.HostBatch(recs)
        for T, W in ((20, 64), (22, 50), (20, 32)):
            iv, first = ctx.sdust(hb, T, W)
            wiv, wfirst = oracle_sdust(recs, T, W)
            assert (first == wfirst).all() and len(iv) == len(wiv) and (iv == wiv).all(), (seed, T, W)


def test_abi_errors(ctx, capi):
    hb = capi.HostBatch([b"ACGT"])
    with pytest.raises(capi.CornError):
        ctx.telofind(hb, "")
    # window sizes: whatever the reference runs with, runs (golden: -w 3, 129, 200, 500); below 3 the reference
    # crashes without output and this returns no interval; above 1024 its 32-bit score products overflow: refused
    for w in (2, 0, -7):
        iv, first = ctx.sdust(capi.HostBatch([b"A" * 300, b"ACGT"]), 20, w)
        assert len(iv) == 0 and list(first) == [0, 0, 0]
    with pytest.raises(capi.CornError):
        ctx.sdust(hb, 20, 1025)
    # empty batch is fine
    e = capi.HostBatch([])
    assert len(ctx.telofind(e)) == 0
    iv, first = ctx.sdust(e)
    assert len(iv) == 0 and list(first) == [0]


def test_abi_async_fused_matches_sync(ctx, capi):
    """telofind_dev(out=NULL) returns without a host sync; the fused telowin(hits=NULL) settles it.
    Results must equal the synchronous host-buffer path, also when the speculative event/run
    buffers are too small (dense batch) and the sparse phase has to be repeated."""
    thr = 0.4 * 0.999 ** 6
    cases = [
        [s for _, s in synth.assembly(51, [300_000, 80_000, 1200, 999], telo=(100, 600))],
        [np.frombuffer(b"TTAGGGA" * 40_000, dtype=np.uint8), np.frombuffer(b"CCCTAAG" * 30_000, dtype=np.uint8)],   # one run per 7 bases
    ]
    for recs in cases:
        hb = capi.HostBatch(recs)
        want_runs = ctx.telofind(hb, "TTAGGG")
        want_wins = ctx.telowin(thr)
        c2 = capi.Context(0)                      # fresh context: buffers start at their speculative sizes
        db = c2.upload(hb)
        for it in range(3):
            assert c2.telofind_dev(db, "TTAGGG", fetch=False) is None
            wins = c2.telowin(thr)
            assert len(wins) == len(want_wins) and (wins == want_wins).all()
            t = c2.timing()
            if it > 0:                            # (the first pass may have had to grow its buffers and repeat)
                assert t["scan_ms"] > 0 and t["launches"] >= 6, t
        runs = c2.telofind_dev(db, "TTAGGG", fetch=True)
        assert len(runs) == len(want_runs) and (runs == want_runs).all()
        c2.free(db)
        c2.close()
        hb.close()


def test_cli_multi_gpu_workers(tmp_path, oracle_bin):
    """CORNETTO_GPUS=N: batches round-robin over N contexts (N devices when present), stdout in batch
    order.  With one visible device the request is clamped and the two workers share it."""
    recs = synth.assembly(61, [60_000, 9_000, 33_000, 1200, 999, 48_000, 7, 21_000], n_gaps=2, telo=(40, 300), microsat_per_mb=1500.0)
    fa = write(str(tmp_path / "m.fa"), synth.fasta_bytes(recs))
    want_t, _, _ = run([oracle_bin, "telofind", fa])
    want_s, _, _ = run([oracle_bin, "sdust", fa])
    for gpus in ("1", "2", "4"):
        env = {"CORNETTO_GPUS": gpus, "CORNETTO_BATCH_BYTES": "40000"}
        out, _, _ = cornetto(["telofind", fa], env=env)
        assert out == want_t, gpus
        out, _, _ = cornetto(["sdust", fa], env=env)
        assert out == want_s, gpus


def test_abi_many_short_reads(ctx, capi):
    """FASTQ-like batch (BASELINE.json configs[4] in miniature): tens of thousands of records, many
    shorter than a tile, some shorter than the motif or empty -- exercises the record tables."""
    rng = np.random.default_rng(71)
    lens = np.maximum(0, (rng.lognormal(0.0, 0.9, size=30_000) * 900).astype(np.int64) - 50)
    lens[::997] = 0
    lens[1::991] = 5
    recs = []
    for i, L in enumerate(lens):
        s = synth.random_dna(rng, int(L))
        if L >= 60 and i % 7 == 0:
            k = int(rng.integers(2, 9))
            s[-6 * k:] = np.tile(np.frombuffer(b"TTAGGG", dtype=np.uint8), k)
        if L >= 60 and i % 11 == 0:
            k = int(rng.integers(2, 9))
            s[:6 * k] = np.tile(np.frombuffer(b"CCCTAA", dtype=np.uint8), k)
        recs.append(s)
    hb = capi.HostBatch(recs)
    got = as_rows(ctx.telofind(hb, "TTAGGG"))
    want = oracle_telofind(recs, "TTAGGG")
    assert got.shape == want.shape and (got == want).all()
    iv, first = ctx.sdust(hb, 20, 64)
    wiv, wfirst = oracle_sdust(recs, 20, 64)
    assert (first == wfirst).all() and len(iv) == len(wiv) and (iv == wiv).all()
    hb.close()
