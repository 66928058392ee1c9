#!/usr/bin/env python3
"""Runs corn_gpu_sdust_dev on a synthetic resident batch (for ncu / timing).  usage: prof_sdust.py [Mb] [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cornetto_b200 import capi  # noqa: E402

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 200
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lengths = [mb * 1_000_000 // 4] * 4
ctx = capi.Context(0)
db = ctx.alloc(lengths)
ctx.fill_random(db, 42)
mode = sys.argv[3] if len(sys.argv) > 3 else "full"     # full | plain (random only) | notelo (no telomeric ends) | gaps (full + N gaps)
if mode != "plain":
    tand, lower, gaps = bench.make_features(capi, lengths, 7, n_gaps=3 if mode == "gaps" else 0)
    if mode == "notelo":
        tand = tand[tand["len"] < 3000]
    ctx.apply_features(db, tand)
    ctx.apply_features(db, lower)
    if len(gaps):
        ctx.apply_features(db, gaps)
for _ in range(reps):
    iv, first = ctx.sdust_dev(db)
    t = ctx.timing()
    print(f"sdust {sum(lengths) / 1e6:.0f} Mb: kernel {t['scan_ms']:.2f} ms, post {t['post_ms']:.3f} ms, {len(iv)} intervals, "
          f"{sum(lengths) / t['scan_ms'] / 1e6:.2f} Gbases/s")
