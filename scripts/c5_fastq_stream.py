#!/usr/bin/env python3
"""BASELINE.json configs[4] ("C5"): `cornetto telofind` on an on-disk FASTQ of ultra-long reads (N50 100 kb), streamed
to N GPUs by the drop-in binary (CORNETTO_GPUS=N: file blocks cut at record boundaries, parsed on the device, stdout
in file order).  100 Gb of reads is ~200 GB of FASTQ -- impractical on the bench box -- so the file is `Gb` gigabases
(default 5 => 10 GB on disk) and the rate is what extrapolates.

    c5_fastq_stream.py [Gb=5] [gpus=8] [multiline_permille=0]

Checks: the output for the first ~300 Mb of reads (a prefix of the file, written as its own FASTQ) must be
byte-identical to the compiled reference's output on that prefix file.  With multiline_permille > 0 that share of the
records is written as multi-line FASTQ (sequence and quality wrapped at 80 columns, legal for kseq, src/kseq.h:201-223):
the device parser only takes four-line records, reports the block as irregular and the serial reader takes over from
there -- the JSON then shows what that costs.
Prints one JSON line."""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 8
multi = int(sys.argv[3]) if len(sys.argv) > 3 else 0
n50 = 100_000
ours = os.path.join(ROOT, "cornetto_b200", "bin", "cornetto")
ref, kind = bench.ref_binary()
rng = np.random.default_rng(11)
d = os.environ.get("CORN_C5_DIR", "/tmp")
fq, prefix_fq = os.path.join(d, "corn_c5.fq"), os.path.join(d, "corn_c5_prefix.fq")
genome = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=128_000_000, dtype=np.uint8)]
low = genome[:1_000_000].copy()
low[:] = low | 0x20                                     # a soft-masked stretch
genome[5_000_000:6_000_000] = low
tel = np.tile(np.frombuffer(b"TTAGGG", dtype=np.uint8), 3000)
sigma = 0.6
total, n_reads, prefix_reads, prefix_bases, n_multi = 0, 0, 0, 0, 0
t_gen = time.perf_counter()


def wrap(a, width=80):
    full = len(a) // width * width
    body = np.empty((full // width, width + 1), dtype=np.uint8)
    body[:, :width] = a[:full].reshape(-1, width)
    body[:, width] = 10
    return body.tobytes() + (a[full:].tobytes() + b"\n" if full < len(a) else b"")


with open(fq, "wb", buffering=1 << 24) as f, open(prefix_fq, "wb", buffering=1 << 24) as fp:
    while total < gb * 1e9:
        L = max(200, int(rng.lognormal(0.0, sigma) * n50 / np.exp(sigma * sigma)))
        a = int(rng.integers(0, len(genome) - L)) if L < len(genome) else 0
        s = genome[a:a + L]
        if rng.random() < 0.01:                          # 1 % of the reads end in a telomere
            s = s.copy()
            k = min(len(tel), L // 2)
            s[L - k:] = tel[:k]
        n_reads += 1
        q = np.full(len(s), ord("I"), dtype=np.uint8)
        if multi and rng.integers(0, 1000) < multi:
            rec = b"@read_%d multi\n" % n_reads + wrap(s) + b"+\n" + wrap(q)
            n_multi += 1
        else:
            rec = b"@read_%d\n" % n_reads + s.tobytes() + b"\n+\n" + q.tobytes() + b"\n"
        f.write(rec)
        if prefix_bases < 300_000_000:
            fp.write(rec)
            prefix_reads, prefix_bases = n_reads, prefix_bases + len(s)
        total += len(s)
t_gen = time.perf_counter() - t_gen
subprocess.run(["cat", fq], stdout=subprocess.DEVNULL)     # page cache warm (the box has no fast disk to measure)


def wall(cmd, out, env=None):
    e = dict(os.environ)
    e.update(env or {})
    t0 = time.perf_counter()
    with open(out, "wb") as f:
        p = subprocess.run(cmd, stdout=f, stderr=subprocess.PIPE, env=e)
    if p.returncode != 0:
        raise SystemExit(f"{cmd} failed: {p.stderr[-2000:].decode(errors='replace')}")
    return time.perf_counter() - t0


out_n, out_1, out_ref = os.path.join(d, "corn_c5.out_n"), os.path.join(d, "corn_c5.out_1"), os.path.join(d, "corn_c5.out_ref")
t_ref = wall([ref, "telofind", prefix_fq], out_ref)
wall([ours, "telofind", prefix_fq], out_1, {"CORNETTO_GPUS": "1"})              # CUDA module load of the box's first process
t_n = min(wall([ours, "telofind", fq], out_n, {"CORNETTO_GPUS": str(gpus)}) for _ in range(2))
t_1 = wall([ours, "telofind", fq], out_1, {"CORNETTO_GPUS": "1"})
want = open(out_ref, "rb").read()
got_n = open(out_n, "rb").read()
same_n1 = got_n == open(out_1, "rb").read()
# the lines of the prefix reads are a prefix of the output (file order): exactly that many bytes must agree
print(json.dumps({"what": "C5: cornetto telofind on an on-disk long-read FASTQ, CORNETTO_GPUS=%d" % gpus, "input_Gb": total / 1e9, "reads": n_reads,
                  "multi_line_records": n_multi, "file_GB": os.path.getsize(fq) / 1e9, "generate_s": round(t_gen, 1), "reference_kind": kind,
                  "telofind": {"gpus": gpus, "ours_s": round(t_n, 3), "ours_Gbases_per_s": round(total / 1e9 / t_n, 2),
                               "ours_file_GB_per_s": round(os.path.getsize(fq) / 1e9 / t_n, 2),
                               "ours_1gpu_s": round(t_1, 3), "ours_1gpu_Gbases_per_s": round(total / 1e9 / t_1, 2),
                               "output_lines": got_n.count(b"\n"), "same_output_1_vs_n_gpus": same_n1,
                               "prefix_reads_checked": prefix_reads, "prefix_bases": prefix_bases, "prefix_identical_to_reference": bool(got_n[:len(want)] == want),
                               "reference_prefix_s": round(t_ref, 3), "reference_Gbases_per_s": round(prefix_bases / 1e9 / t_ref, 3)},
                  "bound": "host side: read(2) of the file from the page cache and the staged (pageable -> page-locked ring) H2D copies of all workers share one "
                           "host memory system; the scan itself runs at >4000 Gbases/s per GPU"}))
for p in (fq, prefix_fq, out_n, out_1, out_ref):
    try:
        os.remove(p)
    except OSError:
        pass
