#!/usr/bin/env python3
"""CLI wall clock and peak RSS against the text block size of the device parser.  usage: block_sweep.py [Mb=3000]"""
import os
import resource
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
ours = os.path.join(ROOT, "cornetto_b200", "bin", "cornetto")
fa = "/tmp/sweep.fa"
rng = np.random.default_rng(3)
bench.write_fasta(fa, [(f"chr{i + 1}", bench.host_random_contig(rng, mb * 1_000_000 // 8)) for i in range(8)])
subprocess.run(["cat", fa], stdout=subprocess.DEVNULL)
ref_out = {}
for mbs in (0, 2048, 1024, 512, 256):
    for cmd in ("telofind", "sdust"):
        ts, rss = [], 0
        for _ in range(3):
            env = dict(os.environ)
            if mbs:
                env["CORNETTO_BATCH_MB"] = str(mbs)
            t0 = time.perf_counter()
            p = subprocess.run([ours, cmd, fa], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, check=True)
            ts.append(time.perf_counter() - t0)
            foot = [l for l in p.stderr.decode().splitlines() if "Peak RAM" in l]
            rss = foot[-1].split("Peak RAM:")[1].strip() if foot else "?"
            ref_out.setdefault(cmd, p.stdout)
            assert p.stdout == ref_out[cmd], (cmd, mbs)
        print(f"block {mbs or 'file'} MB  {cmd:8s} wall {min(ts):.3f} s (of {[round(t, 2) for t in ts]})  peak RSS {rss}", flush=True)
