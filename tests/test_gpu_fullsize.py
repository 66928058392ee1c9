"""Full-size parity on the GPU box (BASELINE.json configs[1] and [3]): the 3.1 Gb synthetic
T2T-like assembly that bench.py scans, generated in HBM.

The assembly carries N gaps (three per contig, 1..50 000 bases: BASELINE.md's "N-gap variant for exactness"), so the
stale-window behaviour of sdust at non-ACGT bytes is exercised at full size.  The oracle cannot scan all 3.1 Gb inside a
test budget, so parity at this size is checked by
 - exact comparison against the oracle on WHOLE contigs downloaded from the device: the longest (248 Mb), the
   shortest and a mid-sized one (~0.4 Gb together), for telofind and for sdust,
 - size-independent properties over ALL results: every run is a tandem repeat of the motif and
   is maximal (checked on the bytes of a random sample of 2000 runs), runs are sorted and disjoint
   per (record, strand), strand-0 runs precede strand-1 runs per record, fused == general telowin,
   every window passes the reference's own double-precision test, window counts re-derived from
   the runs on the host, sdust interval lists sorted/disjoint/non-touching per record.
"""
import os
import sys

import numpy as np
import pytest

from util import ROOT

pytestmark = [pytest.mark.gpu, pytest.mark.slow]
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def world():
    import bench
    from cornetto_b200 import capi
    from cornetto_b200.build import ensure_built
    ensure_built()
    ctx = capi.Context(0)
    lengths = bench.workload_lengths(os.environ.get("CORN_TEST_WORKLOAD", "c2"))
    db = ctx.alloc(lengths)
    ctx.fill_random(db, 42, rec_id=np.arange(len(lengths), dtype=np.uint32))
    tand, lower, gaps = bench.make_features(capi, lengths, 7, n_gaps=3)
    assert len(gaps) >= 2 * len(lengths)
    ctx.apply_features(db, tand)
    ctx.apply_features(db, lower)
    ctx.apply_features(db, gaps)
    yield ctx, capi, db, lengths, bench
    ctx.free(db)
    ctx.close()


def test_fullsize_telofind_telowin(world):
    from test_gpu_parity import oracle_telofind, as_rows
    ctx, capi, db, lengths, bench = world
    runs = ctx.telofind_dev(db, "TTAGGG")
    assert len(runs) > 1000 * (sum(lengths) // 10**7)
    rec, strand, start, end = (runs[k].astype(np.int64) for k in ("rec", "strand", "start", "end"))
    # order: record-major, strand 0 before strand 1, ascending and disjoint inside (record, strand)
    key = rec * 2 + strand
    assert (np.diff(key) >= 0).all()
    same = np.diff(key) == 0
    assert (start[1:][same] > end[:-1][same]).all() or (start[1:][same] >= end[:-1][same]).all()
    assert ((end - start) % 6 == 0).all() and (end > start).all()
    assert (end <= np.asarray(lengths, dtype=np.int64)[rec]).all()
    # exact vs the oracle on three whole contigs: the longest, the shortest and a mid-sized one
    order = np.argsort(lengths)
    for r in (order[-1], order[0], order[len(order) // 2]):
        seq = ctx.download(db, int(r), int(lengths[r]))
        want = oracle_telofind([seq], "TTAGGG")
        got = as_rows(runs[runs["rec"] == r])
        got[:, 0] = 0
        assert got.shape == want.shape and (got == want).all(), f"contig {r}"
    # bytes of a random sample of runs: tandem repeat, and maximal on both sides
    rng = np.random.default_rng(0)
    longest = int(np.argmax(lengths))
    seq = ctx.download(db, longest, int(lengths[longest]))
    up = seq & 0xDF
    mine = runs[runs["rec"] == longest]
    pick = mine[rng.choice(len(mine), size=min(2000, len(mine)), replace=False)]
    for r in pick:
        pat = np.frombuffer(b"TTAGGG" if r["strand"] == 0 else b"CCCTAA", dtype=np.uint8)
        s, e = int(r["start"]), int(r["end"])
        assert (up[s:e].reshape(-1, 6) == pat).all()
        assert s < 6 or not (up[s - 6:s] == pat).all()
        assert e + 6 > len(up) or not (up[e:e + 6] == pat).all()
    # telowin: fused (device-resident runs) == general (uploaded runs), and == a host recount
    thr = 0.4 * 0.999 ** 6
    ctx.telofind_dev(db, "TTAGGG", fetch=False)
    fused = ctx.telowin(thr)
    general = ctx.telowin(thr, runs=runs, lengths=np.asarray(lengths, dtype=np.uint32))
    assert len(fused) == len(general) and (fused == general).all() and len(fused) > 0
    marks = np.zeros(int(lengths[longest]) + 1200, dtype=np.int32)
    for r in mine:
        marks[r["start"]:r["end"]] = 1
    cum = np.concatenate([[0], np.cumsum(marks)])
    L = int(lengths[longest])
    want_w = []
    i = 0
    while True:
        car = int(cum[min(i + 1000, L)] - cum[i])
        den = 1000 if i + 1000 < L else L - i
        if car / den >= thr:
            want_w.append((i, i + den, car))
        if i + 1000 >= L:
            break
        i += 200
    got_w = [(int(w["start"]), int(w["end"]), int(w["car"])) for w in fused[fused["rec"] == longest]]
    assert got_w == want_w


def test_fullsize_sdust(world):
    from test_gpu_parity import oracle_sdust
    ctx, capi, db, lengths, bench = world
    iv, first = ctx.sdust_dev(db)
    assert len(iv) > 100 * (sum(lengths) // 10**6) and int(first[-1]) == len(iv)
    s = (iv >> np.uint64(32)).astype(np.int64)
    f = (iv & np.uint64(0xFFFFFFFF)).astype(np.int64)
    assert (f > s).all()
    for r in range(len(lengths)):
        a, b = int(first[r]), int(first[r + 1])
        assert (s[a + 1:b] > f[a:b - 1]).all()          # sorted, disjoint, not touching
    order = np.argsort(lengths)
    for r in (int(order[-1]), int(order[0]), int(order[len(order) // 2])):      # whole contigs, N gaps included
        seq = ctx.download(db, r, int(lengths[r]))
        assert (seq == ord("N")).sum() > 0 or lengths[r] <= 200_000
        wiv, _ = oracle_sdust([seq], 20, 64)
        got = iv[int(first[r]):int(first[r + 1])]
        assert len(got) == len(wiv) and (got == wiv).all(), f"contig {r}"
