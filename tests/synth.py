"""Seeded synthetic inputs for the parity tests.

Everything here is deterministic in (seed, parameters) so that the same bytes can be
regenerated on the GPU box without shipping large files.
"""
from __future__ import annotations

import gzip
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_dna(rng: np.random.Generator, n: int) -> np.ndarray:
    return ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def tandem(rng, unit: bytes, n_rep: int, p_variant: float = 0.0) -> np.ndarray:
    """n_rep copies of unit; each copy gets one random substituted base with prob p_variant."""
    u = np.frombuffer(unit, dtype=np.uint8)
    a = np.tile(u, n_rep)
    if p_variant > 0 and n_rep > 0:
        hit = np.nonzero(rng.random(n_rep) < p_variant)[0]
        for k in hit:
            a[k * len(u) + rng.integers(0, len(u))] = ACGT[rng.integers(0, 4)]
    return a


def make_contig(rng, length: int, *, telo=(500, 2500), p_variant=0.02, p_lower=0.05,
                n_its=3, microsat_per_mb=50.0, n_gaps=0, gap_len=(1, 400),
                iupac_per_mb=0.0) -> np.ndarray:
    """One T2T-like contig: (CCCTAA)n ... (TTAGGG)n ends, interstitial telomere-like blocks,
    microsatellites / homopolymers, optional N gaps and IUPAC codes, soft-masked lower case."""
    s = random_dna(rng, length)
    if telo is not None and length > 2 * 6 * telo[1] + 10:
        n5 = int(rng.integers(telo[0], telo[1] + 1))
        n3 = int(rng.integers(telo[0], telo[1] + 1))
        t5 = tandem(rng, b"CCCTAA", n5, p_variant)
        t3 = tandem(rng, b"TTAGGG", n3, p_variant)
        s[: len(t5)] = t5
        s[length - len(t3):] = t3
    for _ in range(n_its):
        if length < 2000:
            break
        unit = b"TTAGGG" if rng.random() < 0.5 else b"CCCTAA"
        blk = tandem(rng, unit, int(rng.integers(5, 41)), p_variant)
        p = int(rng.integers(0, length - len(blk)))
        s[p:p + len(blk)] = blk
    n_ms = int(microsat_per_mb * length / 1e6)
    for _ in range(n_ms):
        period = int(rng.integers(1, 7))
        total = int(rng.integers(20, 301))
        if total >= length:
            continue
        unit = bytes(random_dna(rng, period))
        blk = np.tile(np.frombuffer(unit, dtype=np.uint8), total // period + 1)[:total].copy()
        noise = rng.random(total) < 0.05
        blk[noise] = random_dna(rng, int(noise.sum()))
        p = int(rng.integers(0, length - total))
        s[p:p + total] = blk
    for _ in range(n_gaps):
        g = int(rng.integers(gap_len[0], gap_len[1] + 1))
        if g >= length:
            continue
        p = int(rng.integers(0, length - g))
        s[p:p + g] = ord("N")
    n_iu = int(iupac_per_mb * length / 1e6)
    if n_iu and length > 0:
        codes = np.frombuffer(b"RYKMSWBDHVNU", dtype=np.uint8)
        pos = rng.integers(0, length, size=n_iu)
        s[pos] = codes[rng.integers(0, len(codes), size=n_iu)]
    if p_lower > 0 and length > 0:
        # soft-masked stretches rather than isolated bases
        n_blk = max(1, int(p_lower * length / 200))
        for _ in range(n_blk):
            p = int(rng.integers(0, length))
            q = min(length, p + int(rng.integers(20, 400)))
            seg = s[p:q]
            up = (seg >= 65) & (seg <= 90)
            seg[up] += 32
    return s


def fasta_bytes(records, width: int = 60, crlf: bool = False) -> bytes:
    """records: iterable of (name, np.uint8 array | bytes)."""
    out = bytearray()
    eol = b"\r\n" if crlf else b"\n"
    for name, seq in records:
        out += b">" + (name.encode() if isinstance(name, str) else name) + eol
        b = bytes(seq)
        if width <= 0:
            out += b + eol
        else:
            for i in range(0, len(b), width):
                out += b[i:i + width] + eol
    return bytes(out)


def fastq_bytes(records, width: int = 0) -> bytes:
    out = bytearray()
    for name, seq in records:
        b = bytes(seq)
        out += b"@" + (name.encode() if isinstance(name, str) else name) + b"\n"
        if width <= 0:
            out += b + b"\n+\n" + b"I" * len(b) + b"\n"
        else:
            for i in range(0, len(b), width):
                out += b[i:i + width] + b"\n"
            out += b"+\n"
            q = b"I" * len(b)
            for i in range(0, len(q), width):
                out += q[i:i + width] + b"\n"
    return bytes(out)


def bgzf_bytes(data: bytes, block: int = 65280) -> bytes:
    """BGZF (bgzip) container of `data`: gzip members of <= 64 KiB with the 'BC' extra field, plus the empty end member."""
    import struct
    import zlib
    out = []
    for i in list(range(0, len(data), block)) + [None]:
        chunk = data[i:i + block] if i is not None else b""
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp +
                   struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    return b"".join(out)


def depth_arrays(seed: int, lengths):
    """per-contig (depth, mq_depth) integer arrays for noboringbits: Poisson coverage with low / high / zero stretches and
    stretches of low mapq coverage; one value above 65535 (truncated by the reader, src/boringbits_main.c:252-259)."""
    rng = np.random.default_rng(seed)
    out = []
    for i, L in enumerate(lengths):
        d = rng.poisson(30, size=L).astype(np.int64)
        for _ in range(max(1, L // 4000)):
            s0, n, k = int(rng.integers(0, L)), int(rng.integers(50, 900)), int(rng.integers(0, 4))
            seg = d[s0:s0 + n]
            if k == 0:
                seg[:] = rng.poisson(5, size=len(seg))
            elif k == 1:
                seg[:] = rng.poisson(120, size=len(seg))
            elif k == 2:
                seg[:] = 0
        q = (d * rng.uniform(0.2, 1.0, size=L)).astype(np.int64)
        for _ in range(max(1, L // 6000)):
            s0, n = int(rng.integers(0, L)), int(rng.integers(50, 900))
            q[s0:s0 + n] //= 5
        if i == 1 and L > 200:
            d[100], q[100] = 70000, 66000
        out.append((f"ctg{i + 1}", d, q))
    return out


def bedgraph_bytes(named, which: int) -> bytes:
    """one line per base: name, pos, pos+1, depth (which = 1) or mq depth (which = 2)"""
    return "".join("".join(f"{nm}\t{p}\t{p + 1}\t{v}\n" for p, v in enumerate(arr)) for nm, *arrs in named for arr in [arrs[which - 1]]).encode()


BITS_OPTS = [[], ["-m", "10000", "-e", "1000"], ["-m", "2000", "-e", "100", "-w", "1000", "-i", "30"],
             ["-m", "5000", "-e", "500", "-w", "777", "-i", "50", "-L", "0.6", "-H", "1.6", "-Q", "0.6"], ["-m", "2500", "-e", "0", "-w", "50", "-i", "50"],
             ["-m", "2000", "-e", "10", "-w", "20000", "-i", "7"]]
BITS_LENGTHS = [30000, 12000, 900, 2500, 2549, 7]


def _low_complexity(rng):
    kind = rng.integers(0, 3)
    if kind == 0:
        return np.full(int(rng.integers(5, 14)), ACGT[rng.integers(0, 4)], dtype=np.uint8)
    per = int(rng.integers(1, 7))
    return np.tile(ACGT[rng.integers(0, 4, size=per)], int(rng.integers(2, 40)))[:int(rng.integers(6, 120))]


def stale_window_record(rng, n_events: int):
    """sdust's stale-window quirk at full strength: runs of N (1..5000) with low-complexity sequence shortly BEFORE them
    (what the stale window holds) and 0..140 bases AFTER them (where the window start is still pinned, src/sdust/sdust.c:146,
    and perfect intervals live up to 2W steps instead of W).  A two-phase sdust that drains only W positions after a
    trigger loses intervals on this input."""
    parts = []
    for _ in range(n_events):
        parts.append(ACGT[rng.integers(0, 4, size=int(rng.integers(0, 300)))])
        if rng.random() < 0.5:
            parts.append(_low_complexity(rng))
        parts.append(ACGT[rng.integers(0, 4, size=int(rng.integers(0, 70)))])
        parts.append(np.full(int(rng.choice([1, 2, 3, 10, 64, 200, 1600, 5000])), ord("N"), dtype=np.uint8))
        parts.append(ACGT[rng.integers(0, 4, size=int(rng.integers(0, 140)))])
        parts.append(_low_complexity(rng))
        if rng.random() < 0.3:
            parts.append(ACGT[rng.integers(0, 4, size=int(rng.integers(0, 70)))])
            parts.append(_low_complexity(rng))
    return np.concatenate(parts)


def stale_window_records(seed: int, n_rec: int = 5):
    rng = np.random.default_rng(seed)
    return [(f"stale_{k}", stale_window_record(rng, int(rng.integers(3, 60)))) for k in range(n_rec)]


def assembly(seed: int, lengths, **kw):
    rng = np.random.default_rng(seed)
    return [(f"contig_{i + 1}", make_contig(rng, int(L), **kw)) for i, L in enumerate(lengths)]


def reads(seed: int, n_reads: int, n50: int = 100_000, p_telo: float = 0.01):
    """Log-normal read lengths scaled so that N50 ~ n50."""
    rng = np.random.default_rng(seed)
    sigma = 0.6
    raw = rng.lognormal(mean=0.0, sigma=sigma, size=n_reads)
    # N50 of a log-normal with median m is about m*exp(sigma^2); rescale accordingly
    scale = n50 / np.exp(sigma * sigma)
    out = []
    for i, x in enumerate(raw):
        L = max(50, int(x * scale))
        has_telo = rng.random() < p_telo
        s = make_contig(rng, L, telo=(50, 400) if has_telo else None, n_its=0,
                        microsat_per_mb=50.0, p_lower=0.0)
        out.append((f"read_{i + 1}", s))
    return out


# ------------------------------------------------------------------------------------------
# the quirk corpus.  Each entry: (file name, file bytes)
# ------------------------------------------------------------------------------------------

def quirk_corpus(seed: int = 7):
    rng = np.random.default_rng(seed)
    R = lambda n: bytes(random_dna(rng, n))
    files = {}
    # 1. run ending exactly at record end; record == motif; shorter than motif; empty record
    files["q1_ends.fa"] = fasta_bytes([
        ("end_exact", R(100) + b"TTAGGG" * 5),
        ("just_motif", b"TTAGGG"),
        ("short", b"TTAG"),
        ("empty", b""),
        ("rc_at_start", b"CCCTAA" * 4 + R(50)),
        ("one_base", b"A"),
    ])
    # 2. lower / mixed case
    files["q2_case.fa"] = fasta_bytes([
        ("lower", (R(40) + b"ttagggttaggg" + R(33) + b"ccctaaCCCTAAcccTAA" + R(20))),
        ("mixed", b"TtAgGgTTAGGGttaggg" + R(10).lower() + b"CcCtAa"),
    ])
    # 3. adjacent fwd/rev, near misses, non-ACGT inside candidates
    files["q3_adjacent.fa"] = fasta_bytes([
        ("adj", R(13) + b"TTAGGGCCCTAA" + R(7) + b"CCCTAATTAGGG" + R(3)),
        ("near", b"TTAGG" + R(9) + b"TTAGGGG" + R(5) + b"TTAGGGTTAGG" + R(4) + b"TTAGGGTTAGGGT"),
        ("nonacgt", b"TTNGGG" + b"TTAGGN" + b"NTTAGGGN" + b"TTAGGGNTTAGGG" + b"CCCTAR" + b"CCCTAA" + b"RCCCTAA"),
        ("aliases", b"TTAGGG" + b"TTEGGG" + b"PTAGGG" + b"TTAGWG" + b"CCCPAA" + b"SCCTAA" + b"ccc\x14aa" + b"TT\x01GGG"),
        ("digits", b"12TTAGGG34" + b"TTAGGG" * 3 + b"!@#"),
    ], width=0)
    # 4. parser quirks: CRLF, blank lines, comments, spaces in sequence lines
    files["q4_crlf.fa"] = fasta_bytes([("crlf1 some comment", R(70) + b"TTAGGG" * 3 + R(65)),
                                       ("crlf2\tcomment", b"CCCTAA" * 11)], width=60, crlf=True)
    files["q4_blank.fa"] = (b">blank comment here\nACGTTAGGG\n\nTTAGGGTTAGGG\n\n\nACGT\n"
                            b">space\nTTAG GGTTAGGG\r\nTTAGGG\n"
                            b">tabs\nTTAGGG\tTTAGGG\n"
                            b">crfirst\r\n\r\nTTAGGGTTAGGG\r\n")
    files["q4_multi.fq"] = fastq_bytes([("r1 c", R(150) + b"TTAGGG" * 8), ("r2", b"CCCTAA" * 9 + R(91)),
                                        ("r3", R(33))], width=50)
    files["q4_single.fq"] = fastq_bytes([("s1", R(40) + b"TTAGGG" * 8), ("s2", b"CCCTAA" * 9 + R(9))])
    files["q4_trunc.fq"] = fastq_bytes([("t1", R(40) + b"TTAGGG" * 2)]) + b"@t2\nACGTTAGGGA\n+\nIII\n"
    files["q4_gz.fa.gz"] = gzip.compress(fasta_bytes([("gz1", R(500) + b"TTAGGG" * 20 + R(300)),
                                                      ("gz2", b"CCCTAA" * 50 + R(1000))]), mtime=0)
    # 5. telowin sizes
    telo = lambda n: b"TTAGGG" * n
    files["q5_winsizes.fa"] = fasta_bytes([
        ("lt1000", telo(80) + R(300)),
        ("eq1000", telo(100) + R(400)),
        ("eq1200", R(600) + telo(100)),
        ("lastpartial", R(2950) + telo(40)),
        ("nohits", b"ACAC" * 500),
        ("len1400", telo(120) + R(1400 - 720)),
        ("len2000", R(1000) + telo(120) + R(280)),
        ("len1", b"T"),
    ])
    # 6. sdust quirks
    A = lambda n: b"A" * n
    files["q6_sdust.fa"] = fasta_bytes([
        ("homo7", R(50) + A(7) + R(50)),
        ("homo6", R(50) + A(6) + R(50)),
        ("quirkN", A(100) + b"N" + A(5)),
        ("quirkN2", A(100) + b"N" + b"ACGTA"),
        ("Nrun", R(100) + A(30) + b"N" * 200 + b"AC" * 30 + R(100)),
        ("Nstart", b"NN" + A(20) + R(40) + b"N" + b"CA" * 12),
        ("Nend", R(80) + b"GT" * 20 + b"N" + A(10)),
        ("iupac", R(30) + b"ACACACACACRACACACACACACYACACACAC" + R(30)),
        ("lower", (R(30) + b"AC" * 20 + R(30)).lower()),
        ("long_homo", R(20) + A(300) + R(20)),
        ("micro", R(64) + b"CAG" * 40 + R(64) + b"AT" * 50 + R(64)),
        ("bytes0123", b"\x01\x02\x03" * 20 + R(10)),
    ])
    # 7. kseq learns of the end of the input only from a read shorter than its 16384-byte buffer (src/kseq.h:107-108):
    #    files whose size is a multiple of 16384 and that end in a header character or a lone '\r' behave differently
    #    from the same endings at any other size (one more record with an empty name; the '\r' is dropped)
    def sized(total, head, tail):
        line = b"ACGT" * 15 + b"\n"
        body_len = total - len(head) - len(tail)
        body = line * (body_len // len(line))
        rest = body_len - len(body)
        if rest > 0:
            body += (b"A" * (rest - 1) + b"\n") if rest > 1 else b"\n"
        out = head + body + tail
        assert len(out) == total
        return out
    files["q7_eof16k_gt.fa"] = sized(16384, b">r1 c\n" + b"TTAGGG" * 8 + b"\n", b"\n>")
    files["q7_eof32k_cr.fa"] = sized(32768, b">r1\n", b"\n\r")
    files["q7_eof16k_off.fa"] = sized(16383, b">r1\n", b"\n>")
    return files
