#!/usr/bin/env python3
"""Drop-in check at assembly scale: runs scripts/telostats.sh's cornetto steps (telofind -> re-tab ->
fa2bed -> telowin 99.9 0.4) plus sdust and telobreaks with BOTH the compiled reference and this
repository's binary on the same synthetic T2T-like FASTA and compares every output byte for byte.
usage: full_pipeline_check.py [scale_divisor=1]   (1 = 3.1 Gb; prints one JSON line)"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from util import retab_telomere, lens_from_fa2bed  # noqa: E402

div = int(sys.argv[1]) if len(sys.argv) > 1 else 1
diploid = len(sys.argv) > 2 and sys.argv[2] == "diploid"      # BASELINE.json configs[2]: 48 contigs, 6.2 Gb
ours = os.path.join(ROOT, "cornetto_b200", "bin", "cornetto")
ref, kind = bench.ref_binary()
work = "/tmp/corn_full"
os.makedirs(work, exist_ok=True)
fa = os.path.join(work, "asm.fa")
rng = np.random.default_rng(11)
lengths = [L // div for L in bench.CHM13] * (2 if diploid else 1)
names = [f"chr{i % 24 + 1}" + (("_MATERNAL" if i < 24 else "_PATERNAL") if diploid else "") for i in range(len(lengths))]
t0 = time.perf_counter()
with open(fa, "wb") as f:
    pass
for i, L in enumerate(lengths):                       # contig by contig: bounded memory
    s = bench.host_random_contig(rng, L)
    # a few microsatellites / soft-masked stretches so that sdust and the case folding have work
    for _ in range(max(1, L // 200_000)):
        p = int(rng.integers(0, max(1, L - 400)))
        unit = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(rng.integers(1, 7)))]
        n = int(rng.integers(20, 300))
        s[p:p + n] = np.resize(unit, n)
        q = int(rng.integers(0, max(1, L - 3000)))
        s[q:q + 2000] |= 0x20
    with open(fa, "ab") as f:
        tmp = os.path.join(work, "one.fa")
        bench.write_fasta(tmp, [(names[i], s)])
        f.write(open(tmp, "rb").read())
gen_s = time.perf_counter() - t0
subprocess.run(["cat", fa], stdout=subprocess.DEVNULL)


def run(binary, args, out):
    t = time.perf_counter()
    with open(out, "wb") as f:
        subprocess.run([binary] + args, stdout=f, stderr=subprocess.DEVNULL, check=True)
    return time.perf_counter() - t


res = {"gpus": os.environ.get("CORNETTO_GPUS", "1"), "contigs": len(lengths), "bases": int(sum(lengths)), "fasta_bytes": os.path.getsize(fa), "generate_s": round(gen_s, 1), "reference_kind": kind, "steps": {}}
ok = True
for tag, binary in (("ref", ref), ("ours", ours)):
    d = os.path.join(work, tag)
    os.makedirs(d, exist_ok=True)
    t = {}
    t["telofind"] = run(binary, ["telofind", fa], f"{d}/raw.telomere")
    open(f"{d}/asm.telomere", "wb").write(retab_telomere(open(f"{d}/raw.telomere", "rb").read()))
    t["fa2bed"] = run(binary, ["fa2bed", fa], f"{d}/asm.bed")
    open(f"{d}/asm.lens", "wb").write(lens_from_fa2bed(open(f"{d}/asm.bed", "rb").read()))
    t["telowin"] = run(binary, ["telowin", f"{d}/asm.telomere", "99.9", "0.4"], f"{d}/asm.windows")
    t["sdust"] = run(binary, ["sdust", fa], f"{d}/asm.sdust")
    t["telobreaks"] = run(binary, ["telobreaks", f"{d}/asm.lens", f"{d}/asm.sdust", f"{d}/asm.telomere"], f"{d}/asm.breaks")
    res["steps"][tag] = {k: round(v, 2) for k, v in t.items()}
for name in ("raw.telomere", "asm.bed", "asm.windows", "asm.sdust", "asm.breaks"):
    a = open(f"{work}/ref/{name}", "rb").read()
    b = open(f"{work}/ours/{name}", "rb").read()
    res.setdefault("identical", {})[name] = (a == b)
    res.setdefault("lines", {})[name] = a.count(b"\n")
    ok &= a == b
res["all_identical"] = bool(ok)
print(json.dumps(res))
sys.exit(0 if ok else 1)
