#!/usr/bin/env python3
"""Wall-clock of the drop-in binary vs the compiled reference on a FASTA file (parse included).
usage: cli_bench.py [Mb=400]      (writes /tmp/corn_cli.fa; prints one JSON line)"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 400
ours = os.path.join(ROOT, "cornetto_b200", "bin", "cornetto")
ref, kind = bench.ref_binary()
rng = np.random.default_rng(3)
n_ctg = 8
fa = "/tmp/corn_cli.fa"
bench.write_fasta(fa, [(f"chr{i + 1}", bench.host_random_contig(rng, mb * 1_000_000 // n_ctg)) for i in range(n_ctg)])
subprocess.run(["cat", fa], stdout=subprocess.DEVNULL)          # page cache warm


def wall(cmd, out, env=None):
    e = dict(os.environ)
    e.update(env or {})
    t0 = time.perf_counter()
    with open(out, "wb") as f:
        subprocess.run(cmd, stdout=f, stderr=subprocess.DEVNULL, check=True, env=e)
    return time.perf_counter() - t0


res = {"input_Mb": mb, "reference_kind": kind}
for name, args in (("telofind", ["telofind", fa]), ("sdust", ["sdust", fa])):
    t_ref = wall([ref] + args, f"/tmp/corn_cli.{name}.ref")
    wall([ours] + args, f"/tmp/corn_cli.{name}.ours")                     # first run pays CUDA module load
    t_ours = min(wall([ours] + args, f"/tmp/corn_cli.{name}.ours") for _ in range(2))
    t_serial = min(wall([ours] + args, f"/tmp/corn_cli.{name}.serial", {"CORNETTO_INGEST": "0"}) for _ in range(2))
    same = (open(f"/tmp/corn_cli.{name}.ref", "rb").read() == open(f"/tmp/corn_cli.{name}.ours", "rb").read()
            == open(f"/tmp/corn_cli.{name}.serial", "rb").read())
    res[name] = {"reference_s": round(t_ref, 3), "ours_s": round(t_ours, 3), "ours_serial_reader_s": round(t_serial, 3), "identical_output": same,
                 "ours_Gbases_per_s": round(mb / 1e3 / t_ours, 3), "reference_Gbases_per_s": round(mb / 1e3 / t_ref, 3)}
print(json.dumps(res))
