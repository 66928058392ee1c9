"""CPU-only checks of the host side and of the device-side algorithms compiled for the host.

 - the streaming FASTA/FASTQ reader (cornetto_b200/host/fastx.c) against the oracle's kseq
   restatement, also across tiny batch capacities (carry / grow paths);
 - `cornetto telobreaks` and `cornetto fa2bed` (no GPU involved) against the golden vectors;
 - tests/sim: the kernels' host/device arithmetic (bit-plane matcher, chunked sdust with exact
   warm start and seam fold) against the oracle;
 - without a GPU the scan commands fail loudly instead of falling back.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import golden_util
import synth
from util import ROOT, run, write, retab_telomere, lens_from_fa2bed

BIN = os.path.join(ROOT, "cornetto_b200", "bin", "cornetto")
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "cornetto_b200", "lib")


@pytest.fixture(scope="session")
def built():
    from cornetto_b200.build import ensure_built
    ensure_built()
    return True


@pytest.fixture(scope="session")
def sim_bin(tmp_path_factory, oracle_bin):
    out = str(tmp_path_factory.mktemp("sim") / "sim")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", out, "-x", "c++", os.path.join(ROOT, "tests", "sim", "sim_main.cpp"),
                           "-x", "c", os.path.join(ROOT, "oracle", "oracle.c"), "-lz", "-lm"], stderr=subprocess.DEVNULL)
    return out


@pytest.fixture(scope="session")
def reader_dump(tmp_path_factory, built):
    out = str(tmp_path_factory.mktemp("rd") / "reader_dump")
    host = os.path.join(ROOT, "cornetto_b200", "host")
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-D_GNU_SOURCE", "-I" + INC, "-o", out, os.path.join(ROOT, "tests", "sim", "reader_dump.c"),
                           os.path.join(host, "fastx.c"), os.path.join(host, "misc.c"), "-L" + LIBDIR, "-lcorn_gpu",
                           "-Wl,-rpath," + LIBDIR, "-lz", "-lm"])
    return out


def oracle_records(path):
    L = C.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle.so"))

    class Rec(C.Structure):
        _fields_ = [("name", C.c_char_p), ("seq", C.POINTER(C.c_ubyte)), ("len", C.c_size_t)]
    L.orc_read_fastx.argtypes = [C.c_char_p, C.POINTER(C.POINTER(Rec)), C.POINTER(C.c_size_t)]
    recs, n = C.POINTER(Rec)(), C.c_size_t()
    assert L.orc_read_fastx(path.encode(), C.byref(recs), C.byref(n)) == 0
    return b"".join(recs[i].name + b"\t%d\t" % recs[i].len + bytes(recs[i].seq[:recs[i].len]).hex().encode() + b"\n" for i in range(n.value))


EDGE_FILES = {
    "eof_cr.fa": b">a\nACGT\r", "eof_cr2.fa": b">a\nACGT\n\r", "hdr_only.fa": b">a", "hdr_only2.fa": b">a\n",
    "junk_before.fa": b"junk\n>a b c\nAC\nGT\n>b\n\n\n>c\nA", "plus.fa": b">a\nACGT\n+\nIIII\n>b\nAC\n",
    "fq_noqual.fq": b"@a\nACGT\n+\n", "fq_at_in_qual.fq": b"@a\nACGT\n+\n@III\n@b\nGG\n+\n>I\n",
    "win.fq": b"@a\r\nACGT\r\n+\r\nIIII\r\n@b\r\nAC\r\n+\r\nII\r\n", "empty.fa": b"", "cr_first.fa": b">x\r\n\r\nAC\r\n",
}


def test_reader_matches_oracle(reader_dump, oracle_bin, tmp_path):
    cases = {k: c["input"] for k, c in golden_util.load().items()}
    cases.update(EDGE_FILES)
    for name, data in cases.items():
        p = write(str(tmp_path / name), data)
        want = oracle_records(p)
        for cap in ("64", "4096", str(1 << 20)):
            got, _, _ = run([reader_dump, p, cap])
            assert got == want, (name, cap)
            assert b"BADPAD" not in got and b"BADALIGN" not in got


@pytest.fixture(scope="session")
def ingest_sim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ing") / "ingest_sim")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", out, os.path.join(ROOT, "tests", "sim", "ingest_sim.cpp")])
    return out


INGEST_EDGE = {
    "fq_crlf_lone.fq": b"@a\r\n\r\n+\r\n\r\n@b\nAC\n+\nII\n", "fq_trailing_blank.fq": b"@a\nAC\n+\nII\n\n\r\n",
    "fq_short_qual.fq": b"@a\nACGT\n+\nII\nII\n@b\nA\n+\nI\n", "fq_long_qual.fq": b"@a\nAC\n+\nIIII\n@b\nA\n+\nI\n",
    "fq_multiline.fq": b"@a\nAC\nGT\n+\nII\nII\n@b\nA\n+\nI\n", "fq_empty_seq.fq": b"@a\n\n+\n\n@b\nA\n+\nI\n",
    "fq_gt_seq.fq": b"@a\n>CGT\n+\nIIII\n", "fa_at_line.fa": b">a\nAC\n@b\nGG\n>c\nT\n", "fa_gt_eof.fa": b">a\nAC\n>",
    "fa_gt_eof_nl.fa": b">a\nAC\n>\n", "fa_nul.fa": b">a\nAC\x00GT\n>b\nAA\n", "fa_blank_lines.fa": b">a\n\nAC\n\n\nGT\n\n>b\n\n",
    "fa_crlf.fa": b">a x\r\nACGT\r\nAC\r\n>b\r\nT\r\n", "fa_no_final_nl.fa": b">a\nACGT\nAC", "fa_ws_name.fa": b">a\tb c\nAC\n>\nGG\n> x\nTT\n",
    "fa_inner_cr.fa": b">a\nA\rC\nG\r\r\n", "fq_no_final_nl.fq": b"@a\nAC\n+\nII", "fq_hdr_only.fq": b"@a", "fa_long_line.fa": b">a\n" + b"ACGT" * 700 + b"\n>b\n" + b"T" * 33 + b"\n",
}


def test_ingest_rules_match_oracle(ingest_sim, oracle_bin, tmp_path):
    """The device parser's line rules (ingest_core.cuh, walked on the CPU): whatever they accept must be
    parsed exactly as the oracle's kseq restatement does, also across block seams, and text they reject
    must be rejected at a record boundary from which the serial reader reproduces the rest."""
    cases = {k: c["input"] for k, c in golden_util.load().items()}
    cases.update(EDGE_FILES)
    cases.update(INGEST_EDGE)
    must_be_regular = {"asm_small.fa", "q1_ends.fa", "q2_case.fa", "q3_adjacent.fa", "q4_crlf.fa", "q4_single.fq", "q5_winsizes.fa",
                       "q6_sdust.fa", "reads_small.fq"} | {"win.fq", "fq_at_in_qual.fq", "fa_crlf.fa", "fa_blank_lines.fa",
                                                                        "fa_no_final_nl.fa", "fa_ws_name.fa", "fa_inner_cr.fa",
                                                                        "fq_no_final_nl.fq", "fq_trailing_blank.fq", "fa_long_line.fa"}
    n_regular = 0
    for name, data in cases.items():
        p = write(str(tmp_path / name), data)
        want = oracle_records(p)
        for block in (48, 200, 4096, 1 << 30):
            got, _, _ = run([ingest_sim, p, str(block)])
            lines = got.split(b"\n")
            if len(lines) >= 2 and lines[-2].startswith(b"IRREGULAR\t"):
                off = int(lines[-2].split(b"\t")[1])
                head = b"\n".join(lines[:-2]) + (b"\n" if len(lines) > 2 else b"")
                rest = oracle_records(write(str(tmp_path / (name + ".rest")), data[off:]))
                assert head + rest == want, (name, block, off)
                assert not (block == 1 << 30 and name in must_be_regular), name
            else:
                assert got == want, (name, block)
                n_regular += 1
    assert n_regular > 40


def test_ingest_rules_fuzz(ingest_sim, oracle_bin, tmp_path):
    """Random line soups over the alphabet that matters to kseq ('>', '@', '+', CR, blank lines, NUL):
    accepted text parses like the oracle, rejected text is rejected at a valid resume point."""
    rng = np.random.default_rng(11)
    atoms = [b"ACGT", b"AC", b"A", b"", b"\r", b">", b"@", b"+", b">n1 c", b"@n2", b"+n2", b"IIII", b"II", b"AC\r", b"ACGT\r", b"\x00", b"A C"]
    weights = np.array([8, 6, 3, 2, 1, 1, 1, 2, 4, 4, 2, 6, 4, 3, 3, 0.3, 1], dtype=float)
    weights /= weights.sum()
    n_irregular = n_regular = 0
    for it in range(300):
        n_lines = int(rng.integers(1, 14))
        first = [b">x", b"@x", b">x y", b"@x y"][int(rng.integers(0, 4))] if rng.random() < 0.9 else b"junk"
        lines = [first] + [atoms[int(rng.choice(len(atoms), p=weights))] for _ in range(n_lines)]
        if rng.random() < 0.5:      # bias towards well-formed four-line groups
            lines = []
            for r in range(int(rng.integers(1, 5))):
                s = atoms[int(rng.choice(len(atoms), p=weights))]
                q = b"I" * len(s.rstrip(b"\r")) if rng.random() < 0.8 else atoms[int(rng.choice(len(atoms), p=weights))]
                lines += [b"@r%d" % r, s, b"+", q]
        data = b"\n".join(lines) + (b"\n" if rng.random() < 0.7 else b"")
        p = write(str(tmp_path / "fz"), data)
        want = oracle_records(p)
        for block in (24, 1 << 20):
            got, _, _ = run([ingest_sim, p, str(block)])
            out = got.split(b"\n")
            if len(out) >= 2 and out[-2].startswith(b"IRREGULAR\t"):
                off = int(out[-2].split(b"\t")[1])
                head = b"\n".join(out[:-2]) + (b"\n" if len(out) > 2 else b"")
                rest = oracle_records(write(str(tmp_path / "fz.rest"), data[off:]))
                assert head + rest == want, (data, block, off)
                n_irregular += 1
            else:
                assert got == want, (data, block)
                n_regular += 1
    assert n_regular > 100 and n_irregular > 100


def test_parallel_formatting_equals_serial(built, tmp_path):
    """outbuf_format_parallel (host/misc.c): several threads, same bytes as the serial loop."""
    exe = str(tmp_path / "format_check")
    host = os.path.join(ROOT, "cornetto_b200", "host")
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-D_GNU_SOURCE", "-I" + INC, "-o", exe, os.path.join(ROOT, "tests", "sim", "format_check.c"),
                           os.path.join(host, "misc.c"), "-L" + LIBDIR, "-lcorn_gpu", "-Wl,-rpath," + LIBDIR, "-lpthread", "-lm"])
    for n in ("0", "7", "199999", "200000", "1000003"):
        out, _, _ = run([exe, n])
        assert out.startswith(b"OK "), (n, out)


def test_reader_resumes_at_record_boundaries(reader_dump, ingest_sim, oracle_bin, tmp_path):
    """fastx_open_at(): the serial reader started where the device parser gave up (the offsets ingest_sim reports,
    plus every header offset of a regular file) delivers what the oracle parses from that suffix of the file."""
    cases = {k: c["input"] for k, c in golden_util.load().items() if not k.endswith(".gz")}
    cases.update(EDGE_FILES)
    cases.update(INGEST_EDGE)
    n = 0
    for name, data in cases.items():
        p = write(str(tmp_path / name), data)
        offsets = set()
        for block in (48, 200, 1 << 30):
            got, _, _ = run([ingest_sim, p, str(block)])
            lines = got.split(b"\n")
            if len(lines) >= 2 and lines[-2].startswith(b"IRREGULAR\t"):
                offsets.add(int(lines[-2].split(b"\t")[1]))
        if name in ("q3_adjacent.fa", "q4_single.fq", "fa_crlf.fa"):          # regular files: every line-initial header byte
            offsets.update(i for i in range(len(data)) if data[i:i + 1] in (b">", b"@") and (i == 0 or data[i - 1:i] == b"\n")
                           and (name.endswith(".fa") or data[i:i + 2] in (b"@s", b"@r")))
        for off in sorted(offsets):
            if off == 0:
                continue
            want = oracle_records(write(str(tmp_path / (name + ".rest")), data[off:]))
            got, _, _ = run([reader_dump, p, "4096"], env=dict(os.environ, READER_OFFSET=str(off)))
            assert got == want, (name, off)
            n += 1
    assert n >= 5


def test_nx_and_report_match_reference(built, ref_bin, tmp_path):
    """`cornetto nx` / `cornetto report` (host-only by-products of the reader) against the compiled
    reference: stdout byte for byte, exit codes, usage text; FASTA, FASTQ, gzip, options before and after the file."""
    paths = []
    for name, c in golden_util.load().items():
        if name in ("q4_trunc.fq",):          # (the reference's live assert aborts on nothing here, but keep to well-formed files)
            continue
        paths.append(write(str(tmp_path / name), c["input"]))
    big = write(str(tmp_path / "many.fa"), synth.fasta_bytes(synth.assembly(8, [5000, 1, 777, 120_000, 120_000, 31, 64_000, 9], telo=None)))
    paths.append(big)
    for p in paths:
        for args in (["nx", p], ["nx", "-g", "1.5M", p], ["nx", p, "-g", "250k"], ["nx", "--genome-size", "3G", p], ["report", p]):
            want = run([ref_bin] + args, check=False)
            got = run([BIN] + args, check=False)
            assert got[0] == want[0] and got[2] == want[2], (args, got[0][:200], want[0][:200])
    for args in (["report"] + paths[:4], ["report", big, paths[0]]):
        want, got = run([ref_bin] + args, check=False), run([BIN] + args, check=False)
        assert got[0] == want[0] and got[2] == want[2], args
    for args in (["nx"], ["nx", "-h"], ["report"], ["report", "-h"], ["nx", paths[0], paths[1]], ["nx", "-x", paths[0]]):
        want, got = run([ref_bin] + args, check=False), run([BIN] + args, check=False)
        assert got[0] == want[0] and got[2] == want[2], args
        strip = lambda e: b"\n".join(l for l in e.split(b"\n") if not l.startswith(b"[main"))       # (footer: timings differ)
        assert strip(got[1]) == strip(want[1]), args


def test_seq_matches_reference(built, ref_bin, tmp_path):
    """`cornetto seq` (read-length filter, host only) against the compiled reference: stdout, the totals on stderr and
    exit codes; FASTQ with comments / CRLF / multi-line records, FASTA (no quality: the reference prints what its
    quality buffer held before), gzip, option errors."""
    strip = lambda e: b"\n".join(l for l in e.split(b"\n") if not l.startswith(b"[main"))           # (footer: timings differ)
    files = {k: c["input"] for k, c in golden_util.load().items()}
    files.update({k: v for k, v in EDGE_FILES.items() if k not in ("empty.fa",)})
    files.update({k: v for k, v in INGEST_EDGE.items() if k != "fa_nul.fa"})                            # (NUL: the reference's assert aborts)
    files["mixed.fq"] = b"@a c1 c2\nACGTACGT\n+\nIIIIIIII\n>b fasta in between\nACGTAC\nGT\n@c\tx\r\nACGTACGTAC\r\n+\r\nJJJJJJJJJJ\r\n"
    for name, data in files.items():
        p = write(str(tmp_path / name), data)
        for m in ("0", "7", "9", "30", "100000"):
            want = run([ref_bin, "seq", "-m", m, p], check=False)
            got = run([BIN, "seq", "-m", m, p], check=False)
            assert got[0] == want[0] and got[2] == want[2] and strip(got[1]) == strip(want[1]), (name, m)
    p = write(str(tmp_path / "r.fq"), files["reads_small.fq"])
    for args in (["seq"], ["seq", "-h"], ["seq", p], ["seq", p, "-m", "500"], ["seq", "-m", "-3", p], ["seq", "-x", p], ["seq", "--verbose", "2", p],
                 ["seq", "--min-len", "1000", p], ["seq", p, p]):
        want, got = run([ref_bin] + args, check=False), run([BIN] + args, check=False)
        assert got[0] == want[0] and got[2] == want[2] and strip(got[1]) == strip(want[1]), args


def test_telobreaks_and_fa2bed_match_golden(built, tmp_path):
    for name, c in golden_util.load().items():
        fa = write(str(tmp_path / name), c["input"])
        out, _, _ = run([BIN, "fa2bed", fa])
        assert out == c["fa2bed"], name
        tf = write(str(tmp_path / (name + ".telomere")), retab_telomere(c["telofind"]["TTAGGG"]))
        lf = write(str(tmp_path / (name + ".lens")), lens_from_fa2bed(c["fa2bed"]))
        sf = write(str(tmp_path / (name + ".sdust")), c["sdust"][""])
        out, _, _ = run([BIN, "telobreaks", lf, sf, tf])
        assert out == c["telobreaks"], name


def test_telobreaks_many_contigs(built, oracle_bin, tmp_path):
    """khash bucket order with resizes (>16 contigs), run clamping at both contig ends, the
    99 / 100 bp flank, matched length 18 vs 24, names missing from the lens file, duplicates."""
    rng = np.random.default_rng(5)
    names = [f"h{i % 7}tg{i * 37 % 1000:06d}l_{'MAT' if i % 2 else 'PAT'}" for i in range(300)]
    lens = [int(rng.integers(2000, 9000)) for _ in names]
    lens_txt = "".join(f"{n}\t{l}\n" for n, l in zip(names, lens)) + f"{names[3]}\t{lens[3]}\n"
    sd, tl = [], []
    for n, l in zip(names, lens):
        sd.append(f"{n}\t0\t{int(rng.integers(300, 900))}\n")
        sd.append(f"{n}\t{l - int(rng.integers(300, 900))}\t{l + 40}\n")
        mid = int(rng.integers(1000, l - 1000))
        sd.append(f"{n}\t{mid}\t{mid + 260}\n")
        sd.append(f"{n}\t{mid + 260}\t{mid + 300}\n")          # adjacent: fuses with the previous
        tl.append(f"{n}\t{l}\t1\t0\t{int(rng.choice([18, 24, 120]))}\t{int(rng.choice([18, 24, 120]))}\n")
        tl.append(f"{n}\t{l}\t0\t{l - 60}\t{l}\t60\n")
        tl.append(f"{n}\t{l}\t0\t{mid + int(rng.choice([99, 100, 101]))}\t{mid + 130}\t30\n")
    sd.append("absent\t0\t100\n")
    tl.append("absent\t500\t0\t0\t60\t60\n")
    lf = write(str(tmp_path / "x.lens"), lens_txt.encode())
    sf = write(str(tmp_path / "x.sdust"), "".join(sd).encode())
    tf = write(str(tmp_path / "x.telomere"), "".join(tl).encode())
    want, _, _ = run([oracle_bin, "telobreaks", lf, sf, tf])
    got, _, _ = run([BIN, "telobreaks", lf, sf, tf])
    assert got == want and len(want) > 1000


def test_sim_core_primitives(sim_bin):
    out, _, _ = run([sim_bin, "coretest"])
    assert b"coretest ok" in out


def test_sim_matches_oracle(sim_bin, oracle_bin, tmp_path):
    """The device algorithms, compiled for the host, on the quirk corpus + an N-rich assembly."""
    files = dict(synth.quirk_corpus())
    files["asm.fa"] = synth.fasta_bytes(synth.assembly(1, [120_000, 30_000, 999, 1000, 7], n_gaps=6, iupac_per_mb=100.0, microsat_per_mb=1500.0))
    for name, data in files.items():
        if name.endswith(".gz"):
            continue
        p = write(str(tmp_path / name), data)
        for motif in ("TTAGGG", "TATATA", "AAAAAA", "TTNGGG", "TTAGGGTTAGGG"):
            a, _, _ = run([sim_bin, "telofind", p, motif])
            b, _, _ = run([oracle_bin, "telofind", p, motif])
            assert a == b, (name, motif)
        for opts in ([], ["-w", "32", "-t", "15"], ["-t", "8"], ["-w", "100", "-t", "25"]):
            b, _, _ = run([oracle_bin, "sdust"] + opts + [p])
            for chunk in ("64", "101", "4096"):
                a, _, _ = run([sim_bin, "sdust"] + opts + ["-c", chunk, p])
                assert a == b, (name, opts, chunk)


def test_sim_sdust_n_fuzz(sim_bin, oracle_bin, tmp_path):
    """Differential fuzz of the chunked sdust (exact warm start + seam fold) on N-rich input."""
    for seed in range(6):
        rng = np.random.default_rng(1000 + seed)
        recs = []
        for k in range(8):
            L = int(rng.integers(100, 20000))
            recs.append((f"f{k}", synth.make_contig(rng, L, telo=None, n_its=0, microsat_per_mb=float(rng.choice([50, 3000, 20000])),
                                                     n_gaps=int(rng.integers(0, 40)), gap_len=(1, int(rng.choice([3, 60, 400]))),
                                                     p_lower=0.1, iupac_per_mb=float(rng.choice([0, 2000])))))
        p = write(str(tmp_path / f"fz{seed}.fa"), synth.fasta_bytes(recs))
        for opts in ([], ["-w", "32", "-t", "15"]):
            b, _, _ = run([oracle_bin, "sdust"] + opts + [p])
            for chunk in ("64", "101", "193"):
                a, _, _ = run([sim_bin, "sdust"] + opts + ["-c", chunk, p])
                assert a == b, (seed, opts, chunk)


@pytest.fixture(scope="session")
def sim_vec_bin(tmp_path_factory, oracle_bin):
    """The simulator with the data-parallel find_perfect (what the warp-cooperative device routine evaluates) and the
    slack bound that skips calls which cannot find a candidate."""
    out = str(tmp_path_factory.mktemp("simv") / "sim_vec")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-DSD_USE_VEC", "-DSD_SLACK_STATS", "-o", out, "-x", "c++",
                           os.path.join(ROOT, "tests", "sim", "sim_main.cpp"), "-x", "c", os.path.join(ROOT, "oracle", "oracle.c"), "-lz", "-lm"],
                          stderr=subprocess.DEVNULL)
    return out


def test_sim_sdust_vector_form_and_slack_skip(sim_vec_bin, oracle_bin, tmp_path):
    """sd_find_perfect_vec + the slack bound (sdust_core.cuh) against the oracle: N-rich fuzz, several (T, W), chunk seams;
    and the bound must actually skip calls on ordinary sequence."""
    files = {}
    for seed in range(4):
        rng = np.random.default_rng(3000 + seed)
        recs = []
        for k in range(6):
            L = int(rng.integers(100, 25000))
            recs.append((f"v{k}", synth.make_contig(rng, L, telo=None, n_its=0, microsat_per_mb=float(rng.choice([50, 3000, 20000])),
                                                     n_gaps=int(rng.integers(0, 30)), gap_len=(1, int(rng.choice([3, 60, 400]))),
                                                     p_lower=0.1, iupac_per_mb=float(rng.choice([0, 2000])))))
        files[f"vz{seed}.fa"] = synth.fasta_bytes(recs)
    files["asm.fa"] = synth.fasta_bytes(synth.assembly(2, [150_000, 999, 7], n_gaps=4, microsat_per_mb=1500.0))
    skipped = 0
    for name, data in files.items():
        p = write(str(tmp_path / name), data)
        for opts in ([], ["-w", "32", "-t", "15"], ["-t", "5"], ["-w", "100", "-t", "25"]):
            b, _, _ = run([oracle_bin, "sdust"] + opts + [p])
            for chunk in ("64", "193", "4096"):
                a, err, _ = run([sim_vec_bin, "sdust"] + opts + ["-c", chunk, p])
                assert a == b, (name, opts, chunk)
                skipped += int(err.split(b"evaluated,")[1].split()[0])
    assert skipped > 1000


def test_sim_sdust_two_phase(sim_vec_bin, sim_bin, oracle_bin, tmp_path):
    """The two-phase execution of csrc/sdust.cu's default path, step by step on the CPU: scout (window half without
    loops) -> active 64-base blocks -> items -> full machine per item -> item seam fold; once with every item on the
    lane-per-item machine and once with every N-free item on the age-ordered machine the dense kernel runs (SIM_DENSE)."""
    files = {}
    for seed in range(3):
        rng = np.random.default_rng(7100 + seed)
        recs = []
        for k in range(5):
            L = int(rng.integers(100, 20000))
            recs.append((f"v{k}", synth.make_contig(rng, L, telo=None if k % 2 else (20, 200), n_its=1, microsat_per_mb=float(rng.choice([50, 3000, 20000])),
                                                     n_gaps=int(rng.integers(0, 25)), gap_len=(1, int(rng.choice([3, 60, 400]))),
                                                     p_lower=0.1, iupac_per_mb=float(rng.choice([0, 2000])))))
        files[f"tp{seed}.fa"] = synth.fasta_bytes(recs)
    ends = [("e1", np.frombuffer(b"ACGTTGCA" * 40 + b"A" * 300, dtype=np.uint8)), ("e2", np.frombuffer(b"AC" * 500, dtype=np.uint8)),
            ("e3", np.frombuffer(b"A" * 70, dtype=np.uint8)), ("e4", np.frombuffer(b"TTAGGG" * 1500, dtype=np.uint8)),
            ("e5", np.frombuffer(b"A" * 7, dtype=np.uint8)), ("e6", np.frombuffer(b"AAAAAAAC" * 100, dtype=np.uint8))]
    files["ends.fa"] = synth.fasta_bytes(ends)
    for seed in (9100, 9101):                      # low complexity right after runs of N: the long drain of the stale phase
        files[f"stale{seed}.fa"] = synth.fasta_bytes(synth.stale_window_records(seed))
    n = 0
    for name, data in files.items():
        p = write(str(tmp_path / name), data)
        for opts in ([], ["-w", "32", "-t", "20"], ["-t", "24"], ["-w", "7"]):
            b, _, _ = run([oracle_bin, "sdust"] + opts + [p])
            for chunk in ("64", "193", "2048"):
                for env in (None, dict(os.environ, SIM_DENSE="1")):
                    a, _, _ = run([sim_bin, "sdust2"] + opts + ["-c", chunk, p], env=env)
                    assert a == b, (name, opts, chunk, env is not None)
                    n += 1
    assert n >= 90
    # outside W <= 64, floor(2T/10) == 4 the scout's four-occurrence history does not apply: the simulator says so
    _, _, rc = run([sim_bin, "sdust2", "-t", "15", str(tmp_path / "ends.fa")], check=False)
    assert rc == 2


@pytest.fixture(scope="module")
def sim_wide_bin(tmp_path_factory, oracle_bin):
    """The simulator built on the WIDE instance of sdust_core.cuh (16-bit counters, 64-bit slots: what csrc/sdust_wide.cu runs)."""
    out = str(tmp_path_factory.mktemp("simw") / "sim_wide")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-DSD_WIDE", "-o", out, "-x", "c++",
                           os.path.join(ROOT, "tests", "sim", "sim_main.cpp"), "-x", "c", os.path.join(ROOT, "oracle", "oracle.c"), "-lz", "-lm"],
                          stderr=subprocess.DEVNULL)
    return out


def test_sim_sdust_wide_windows(sim_wide_bin, oracle_bin, tmp_path):
    """Windows beyond 128 (`sdust -w 200`): the generic instance, chunked with seams, against the oracle; N-rich inputs."""
    rng = np.random.default_rng(515)
    recs = [(f"w{k}", synth.make_contig(rng, L, telo=None, n_its=0, microsat_per_mb=3000.0, n_gaps=g, gap_len=(1, gl), p_lower=0.1))
            for k, (L, g, gl) in enumerate(((4000, 4, 300), (1500, 6, 3), (130, 0, 1), (3, 0, 1)))]
    recs.append(("homo", np.frombuffer(b"A" * 400 + b"N" + b"AAAAA", dtype=np.uint8)))
    p = write(str(tmp_path / "wide.fa"), synth.fasta_bytes(recs))
    for opts in (["-w", "129"], ["-w", "200"], ["-w", "333", "-t", "12"]):
        b, _, _ = run([oracle_bin, "sdust"] + opts + [p])
        assert len(b) > 20
        for chunk in ("700", "2048"):
            a, _, _ = run([sim_wide_bin, "sdust"] + opts + ["-c", chunk, p])
            assert a == b, (opts, chunk)


def test_gzip_text_source(tmp_path):
    """host/gzsrc.c (the feeder of the device parser for .gz inputs): same bytes as gzip itself for one member, several
    members, BGZF (members inflated in parallel) and trailing garbage, whatever the request size."""
    import gzip
    exe = str(tmp_path / "gz_dump")
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-D_GNU_SOURCE", "-I" + INC, "-o", exe, os.path.join(ROOT, "tests", "sim", "gz_dump.c"),
                           os.path.join(ROOT, "cornetto_b200", "host", "gzsrc.c"), "-lz", "-lpthread"])
    text = synth.fasta_bytes(synth.assembly(8, [400_000, 150_000, 70_000, 9], n_gaps=1))
    files = {"plain.gz": gzip.compress(text), "bgzf.gz": synth.bgzf_bytes(text), "garbage.gz": gzip.compress(text) + b"\x00\x00junk",
             "multi.gz": gzip.compress(text[:200_000]) + gzip.compress(text[200_000:500_001]) + gzip.compress(text[500_001:])}
    for name, data in files.items():
        p = write(str(tmp_path / name), data)
        for req in ("70000", "1000003", "67108864"):
            out, err, rc = run([exe, p, req])
            assert rc == 0 and out == text, (name, req)
            assert (b"bgzf=1" in err) == (name == "bgzf.gz")
    p = write(str(tmp_path / "bad.gz"), synth.bgzf_bytes(text)[:100_000] + b"\x00" * 64)      # damaged BGZF member chain
    _, _, rc = run([exe, p, "1000003"], check=False)
    assert rc == 3


def test_scan_commands_fail_loudly_without_gpu(built, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    fa = write(str(tmp_path / "a.fa"), b">a\nTTAGGGTTAGGG\n")
    for cmd in (["telofind", fa], ["sdust", fa]):
        out, err, rc = run([BIN] + cmd, check=False)
        assert rc == 1 and out == b"" and b"no usable CUDA device" in err
    # argument errors are reported before any GPU work, exactly like the reference
    _, err, rc = run([BIN, "telofind"], check=False)
    assert rc == 1 and b"Usage: find <input fasta>" in err
    _, err, rc = run([BIN, "sdust"], check=False)
    assert rc == 1 and err.startswith(b"Usage: sdust [-w 64] [-t 20] <in.fa>")
    _, err, rc = run([BIN, "telobreaks", "a", "b"], check=False)
    assert rc == 1 and b"Usage: telobreaks" in err


def test_depth_text_readers(tmp_path):
    """host/depthtxt.c: the parallel reader of noboringbits' two depth tables returns exactly what the one-pass reader
    (the restatement of get_depths(), src/boringbits_main.c:179-293) returns, whatever the block size and thread count --
    or declines (exit 3) for every input the reference would reject, warn about, or tokenise across lines."""
    exe = str(tmp_path / "depthtxt_dump")
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-D_GNU_SOURCE", "-I" + INC, "-o", exe, os.path.join(ROOT, "tests", "sim", "depthtxt_dump.c"),
                           os.path.join(ROOT, "cornetto_b200", "host", "depthtxt.c"), "-lpthread"])
    named = [(nm, np.minimum(d, 65535), np.minimum(q, 65535)) for nm, d, q in synth.depth_arrays(5, [30000, 1, 12000, 900, 7])]
    named.append((named[0][0], named[3][1], named[3][2]))                  # the first name again: a new contig (names are compared with the previous record only)
    f1, f2 = synth.bedgraph_bytes(named, 1), synth.bedgraph_bytes(named, 2)
    p1, p2 = write(str(tmp_path / "a1.bg"), f1), write(str(tmp_path / "a2.bg"), f2)
    want, _, _ = run([exe, p1, p2, "serial"])
    head, _, body = want.partition(b"\n")
    n_tot = sum(len(d) for _, d, _ in named)
    assert head == f"n_ctg {len(named)} n_tot {n_tot} tot_depth {sum(int(d.sum()) for _, d, _ in named)} tot_mq {sum(int(q.sum()) for _, _, q in named)}".encode()
    arrays = want[len(want) - 4 * n_tot:]
    assert arrays == np.concatenate([d for _, d, _ in named]).astype("<u2").tobytes() + np.concatenate([q for _, _, q in named]).astype("<u2").tobytes()
    for threads, block in (("1", "64"), ("3", "100"), ("8", "1000"), ("4", "4096"), ("5", "65537"), ("2", "100000000")):
        got, _, rc = run([exe, p1, p2, "parallel", threads, block], check=False)
        assert rc == 0 and got == want, (threads, block)
    # accepted variants of the same table: CRLF line ends, blank lines, spaces for tabs, no newline at the end
    for k, (g1, g2) in enumerate(((f1.replace(b"\n", b"\r\n"), f2), (f1.replace(b"\n", b"\n\n", 50), f2.replace(b"\t", b"  ")), (f1[:-1], f2[:-1] + b"\n \n"))):
        q1, q2 = write(str(tmp_path / f"v{k}_1.bg"), g1), write(str(tmp_path / f"v{k}_2.bg"), g2)
        for block in ("77", "5000"):
            got, _, rc = run([exe, q1, q2, "parallel", "4", block], check=False)
            assert rc == 0 and got == want, (k, block)
    # declined: everything the one-pass reader reports (or reads differently from a line-by-line parse)
    lines1 = f1.split(b"\n")
    def edited(i, new):
        return b"\n".join(lines1[:i] + [new] + lines1[i + 1:])
    bad = {
        "big": (edited(200, b"ctg1\t200\t201\t70000"), f2),
        "five": (edited(200, b"ctg1\t200\t201\t30\t9"), f2),
        "three": (edited(200, b"ctg1\t200\t201"), f2),
        "split": (edited(200, b"ctg1\t200\n201\t30"), f2),
        "sign": (edited(200, b"ctg1\t200\t201\t+30"), f2),
        "gap": (edited(200, b"ctg1\t201\t202\t30"), f2),
        "end": (edited(200, b"ctg1\t200\t202\t30"), f2),
        "start": (edited(30000, b"ctg2\t5\t6\t30") + b"ctg2\t6\t7\t30\n", f2),
        "names": (f1, f2.replace(b"ctg3\t", b"ctgX\t")),
        "longer": (f1, f2 + b"zz\t0\t1\t3\n"),
        "shorter": (f1 + b"zz\t0\t1\t3\n", f2),
    }
    for k, (g1, g2) in bad.items():
        q1, q2 = write(str(tmp_path / f"b{k}_1.bg"), g1), write(str(tmp_path / f"b{k}_2.bg"), g2)
        for block in ("90", "100000"):
            _, _, rc = run([exe, q1, q2, "parallel", "4", block], check=False)
            assert rc == 3, (k, block)
    # empty inputs
    e = write(str(tmp_path / "empty.bg"), b"")
    got, _, rc = run([exe, e, e, "parallel"], check=False)
    assert rc == 0 and got == b"n_ctg 0 n_tot 0 tot_depth 0 tot_mq 0\n"


def test_depth_text_readers_fuzz(tmp_path):
    """Mutated depth tables (bytes deleted, replaced, duplicated, lines swapped / cut, whitespace injected): whenever the
    parallel reader accepts a pair of files, the one-pass reader accepts it too, silently, with the same result."""
    exe = str(tmp_path / "depthtxt_dump")
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-D_GNU_SOURCE", "-I" + INC, "-o", exe, os.path.join(ROOT, "tests", "sim", "depthtxt_dump.c"),
                           os.path.join(ROOT, "cornetto_b200", "host", "depthtxt.c"), "-lpthread"])
    rng = np.random.default_rng(2024)
    named = [(nm, np.minimum(d, 65535), np.minimum(q, 65535)) for nm, d, q in synth.depth_arrays(9, [300, 1, 120, 45])]
    f1, f2 = synth.bedgraph_bytes(named, 1), synth.bedgraph_bytes(named, 2)
    alphabet = np.frombuffer(b"0123456789\t\n \r-+cx", dtype=np.uint8)
    accepted = 0
    for case in range(160):
        g = [bytearray(f1), bytearray(f2)]
        for _ in range(int(rng.integers(1, 4))):
            b = g[int(rng.integers(0, 2))]
            pos = int(rng.integers(0, len(b)))
            kind = int(rng.integers(0, 6))
            if kind == 0:
                del b[pos]
            elif kind == 1:
                b[pos] = int(alphabet[rng.integers(0, len(alphabet))])
            elif kind == 2:
                b.insert(pos, int(alphabet[rng.integers(0, len(alphabet))]))
            elif kind == 3:                                   # cut the file
                del b[pos:]
            elif kind == 4:                                   # duplicate a line
                a = b.rfind(b"\n", 0, pos) + 1
                e = b.find(b"\n", pos) + 1 or len(b)
                b[a:a] = b[a:e]
            else:                                             # the same edit in both files (still a consistent pair)
                for bb in g:
                    p2 = min(pos, len(bb) - 1)
                    if bb[p2:p2 + 1].isdigit():
                        bb[p2] = ord("1")
        p1, p2 = write(str(tmp_path / "z1.bg"), bytes(g[0])), write(str(tmp_path / "z2.bg"), bytes(g[1]))
        got, _, rc = run([exe, p1, p2, "parallel", str(int(rng.integers(1, 6))), str(int(rng.choice([64, 97, 500, 4096, 1 << 20])))], check=False)
        assert rc in (0, 3), case
        if rc == 0:
            accepted += 1
            want, err, rc2 = run([exe, p1, p2, "serial"], check=False)
            assert rc2 == 0 and err == b"" and got == want, case
    assert accepted >= 10          # (consistent edits and harmless whitespace do get through)
