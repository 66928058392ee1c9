#!/usr/bin/env python3
"""corn_gpu_depthwin (noboringbits' windowed depth scan) on synthetic per-base depth arrays resident in host memory:
kernel time, % of the HBM roofline (algorithmic bytes: 4 per base -- two uint16 arrays), and the reference's
get_regs()-equivalent loop on one core (numpy cumsum restatement is NOT timed; the compiled reference is, on a bedgraph
sample, parse excluded by differencing two runs).  usage: prof_depthwin.py [Mbases=1000] [reps=3]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cornetto_b200 import capi  # noqa: E402

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(3)
lens = [mb * 1_000_000 // 8] * 8
d = [rng.poisson(30, size=L).astype(np.uint16) for L in lens]
q = [(x // 2).astype(np.uint16) for x in d]
for x in d:
    x[1000:60000] = 3                      # a low-coverage stretch per contig
peak, src = bench.load_peak()
ctx = capi.Context(0)
best = None
for _ in range(reps):
    w = ctx.depthwin(d, q, 2500, 50, 12, 75, 0.4, 100000, 1000000, 0)
    t = ctx.timing()
    if best is None or t["scan_ms"] < best["scan_ms"]:
        best = t
n = sum(lens)
print(json.dumps({"what": f"corn_gpu_depthwin, {n / 1e9:.2f} G per-base depth values x 2 arrays, window 2500 / 50", "selected_windows": int(len(w)),
                  "kernel_ms": best["scan_ms"], "compact_ms": best["post_ms"], "h2d_ms": best["h2d_ms"],
                  "gbases_per_s_kernel": n / (best["scan_ms"] * 1e-3) / 1e9,
                  "hbm": {"algorithmic_bytes": 4 * n, "achieved_GBps": 4 * n / (best["scan_ms"] * 1e-3) / 1e9, "peak_GBps": peak, "frac": 4 * n / (best["scan_ms"] * 1e-3) / 1e9 / peak,
                          "peak_source": src},
                  "bound": "HBM for the kernel; end to end the command is bound by parsing one text line per base (tens of GB for a genome)"}))
