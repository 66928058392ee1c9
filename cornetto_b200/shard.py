"""Multi-GPU sharding of the scan path: independent units, no data-path collective.

Records (contigs / reads) are the unit: every record is scanned by exactly one rank, so no halo
is needed between ranks and the per-rank results are simply re-ordered into file order on the
host (the reference prints strictly in file order, src/find_telomere.c:101-105).  Assignment is
longest-first onto the least loaded rank (byte balanced); a 3.1 Gb human assembly splits to
within a few percent over 8 GPUs.  Ranks exchange results with one gather of the sparse lists.
"""
from __future__ import annotations

import numpy as np


def plan_shards(lengths, world: int):
    """-> list (per rank) of ascending record indices; deterministic, byte balanced (LPT)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = [0] * world
    out = [[] for _ in range(world)]
    for idx in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(int(idx))
        load[r] += int(lengths[idx]) + 32
    return [sorted(x) for x in out]


def merge_runs(per_rank_runs, shards, n_rec: int):
    """per_rank_runs[r]: structured array with a 'rec' field indexing shards[r] -> one array in
    file order (record-major; order inside a record is already the reference's)."""
    parts = [None] * n_rec
    for runs, recs in zip(per_rank_runs, shards):
        recs = np.asarray(recs, dtype=np.int64)
        if len(runs) == 0:
            continue
        local = runs["rec"].astype(np.int64)
        # runs of one record are contiguous and ordered inside a rank's result
        bounds = np.flatnonzero(np.diff(local)) + 1
        for seg in np.split(np.arange(len(runs)), bounds):
            g = int(recs[local[seg[0]]])
            piece = runs[seg].copy()
            piece["rec"] = g
            parts[g] = piece
    keep = [p for p in parts if p is not None]
    if not keep:
        return per_rank_runs[0][:0].copy()
    return np.concatenate(keep)


def merge_intervals(per_rank, shards, n_rec: int):
    """per_rank[r] = (iv, rec_first) -> (iv, rec_first) in file order."""
    chunks = [np.zeros(0, dtype=np.uint64)] * n_rec
    for (iv, first), recs in zip(per_rank, shards):
        for k, g in enumerate(recs):
            chunks[g] = iv[int(first[k]):int(first[k + 1])]
    first = np.zeros(n_rec + 1, dtype=np.uint64)
    first[1:] = np.cumsum([len(c) for c in chunks])
    return (np.concatenate(chunks) if n_rec else np.zeros(0, np.uint64)), first
