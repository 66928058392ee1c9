#!/usr/bin/env python3
"""Wall-clock of `cornetto telofind` on a FASTQ of long reads (BASELINE.json configs[4] at reduced size):
drop-in binary with the device parser, with the serial reader, and the compiled reference.
usage: cli_bench_fastq.py [Mb=2000] [N50=100000]      (writes /tmp/corn_cli.fq; prints one JSON line)"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n50 = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
ours = os.path.join(ROOT, "cornetto_b200", "bin", "cornetto")
ref, kind = bench.ref_binary()
rng = np.random.default_rng(9)
fq = "/tmp/corn_cli.fq"
genome = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=64_000_000, dtype=np.uint8)]
tel = np.tile(np.frombuffer(b"TTAGGG", dtype=np.uint8), 2000)
sigma = 0.6
total, n_reads = 0, 0
with open(fq, "wb") as f:
    while total < mb * 1_000_000:
        L = max(200, int(rng.lognormal(0.0, sigma) * n50 / np.exp(sigma * sigma)))
        a = int(rng.integers(0, len(genome) - L)) if L < len(genome) else 0
        s = genome[a:a + L]
        if rng.random() < 0.01:                      # 1 % of the reads end in a telomere
            s = s.copy()
            k = min(len(tel), L // 2)
            s[L - k:] = tel[:k]
        n_reads += 1
        f.write(b"@read_%d\n" % n_reads)
        f.write(s.tobytes())
        f.write(b"\n+\n")
        f.write(b"I" * len(s))
        f.write(b"\n")
        total += len(s)
subprocess.run(["cat", fq], stdout=subprocess.DEVNULL)


def wall(cmd, out, env=None):
    e = dict(os.environ)
    e.update(env or {})
    t0 = time.perf_counter()
    with open(out, "wb") as f:
        subprocess.run(cmd, stdout=f, stderr=subprocess.DEVNULL, check=True, env=e)
    return time.perf_counter() - t0


args = ["telofind", fq]
t_ref = wall([ref] + args, "/tmp/corn_fq.ref")
wall([ours] + args, "/tmp/corn_fq.ours")
t_ours = min(wall([ours] + args, "/tmp/corn_fq.ours") for _ in range(2))
t_serial = min(wall([ours] + args, "/tmp/corn_fq.serial", {"CORNETTO_INGEST": "0"}) for _ in range(2))
same = open("/tmp/corn_fq.ref", "rb").read() == open("/tmp/corn_fq.ours", "rb").read() == open("/tmp/corn_fq.serial", "rb").read()
print(json.dumps({"input_Mb": total / 1e6, "reads": n_reads, "file_GB": os.path.getsize(fq) / 1e9, "reference_kind": kind,
                  "telofind": {"reference_s": round(t_ref, 3), "ours_s": round(t_ours, 3), "ours_serial_reader_s": round(t_serial, 3),
                               "identical_output": same, "ours_Gbases_per_s": round(total / 1e9 / t_ours, 3),
                               "reference_Gbases_per_s": round(total / 1e9 / t_ref, 3)}}))
