#!/usr/bin/env python3
"""`cornetto noboringbits` / `boringbits` end to end on on-disk depth tables (one text line per base, two files), the
drop-in binary against the compiled reference (oracle/_ref/cornetto, when it travelled) or the oracle port.

    bits_cli_bench.py [Mbases_checked=40] [Mbases_big=400]

The small pair of files is run through both programs (outputs must be byte-identical); the big pair only through the
drop-in binary (the reference needs ~0.7 s per Mbase).  Prints one JSON line."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

small = int(sys.argv[1]) if len(sys.argv) > 1 else 40
big = int(sys.argv[2]) if len(sys.argv) > 2 else 400
d = os.environ.get("CORN_BITS_DIR", "/tmp")
ours = os.path.join(ROOT, "cornetto_b200", "bin", "cornetto")
ref, kind = bench.ref_binary()
gen = os.path.join(d, "corn_gen_bedgraph")
subprocess.check_call(["gcc", "-O2", "-o", gen, os.path.join(ROOT, "scripts", "gen_bedgraph.c")])


def wall(cmd, out, env=None):
    e = dict(os.environ)
    e.update(env or {})
    t0 = time.perf_counter()
    with open(out, "wb") as f:
        p = subprocess.run(cmd, stdout=f, stderr=subprocess.PIPE, env=e)
    if p.returncode != 0:
        raise SystemExit(f"{cmd} failed: {p.stderr[-2000:].decode(errors='replace')}")
    return time.perf_counter() - t0


def make(tag, mbases):
    n_ctg = max(2, mbases // 100)
    per = mbases * 1_000_000 // n_ctg
    f1, f2 = os.path.join(d, f"corn_bits_{tag}_total.bg"), os.path.join(d, f"corn_bits_{tag}_mq.bg")
    subprocess.check_call([gen, f1, f2, str(n_ctg), str(per)])
    subprocess.run(["cat", f1, f2], stdout=subprocess.DEVNULL)     # page cache warm
    return f1, f2, n_ctg * per


opts = ["-m", "1000000", "-e", "100000"]
res = {"what": "cornetto noboringbits / boringbits on on-disk per-base depth tables (two text files, one line per base)", "reference_kind": kind}
f1, f2, n = make("small", small)
o_ref, o_us, o_us1 = (os.path.join(d, "corn_bits.out_" + k) for k in ("ref", "us", "us1"))
wall([ours, "noboringbits", f1, "-q", f2] + opts, o_us)            # CUDA module load of the box's first process
chk = {"bases": n, "file_GB": (os.path.getsize(f1) + os.path.getsize(f2)) / 1e9}
for cmd in ("noboringbits", "boringbits"):
    t_ref = wall([ref, cmd, f1, "-q", f2] + opts, o_ref)
    t_us = min(wall([ours, cmd, f1, "-q", f2] + opts, o_us) for _ in range(2))
    t_us1 = wall([ours, cmd, f1, "-q", f2, "-t", "1"] + opts, o_us1)
    a, b, c = (open(p, "rb").read() for p in (o_ref, o_us, o_us1))
    chk[cmd] = {"reference_s": round(t_ref, 2), "ours_s": round(t_us, 2), "ours_one_reader_thread_s": round(t_us1, 2), "speedup": round(t_ref / t_us, 1),
                "output_lines": a.count(b"\n"), "identical": bool(a == b == c)}
res["checked"] = chk
for p in (f1, f2):
    os.remove(p)
if big:
    f1, f2, n = make("big", big)
    t = min(wall([ours, "noboringbits", f1, "-q", f2] + opts, o_us) for _ in range(2))
    t_tr = wall([ours, "noboringbits", f1, "-q", f2] + opts, o_us1, {"CORNETTO_TRACE": "1"})
    res["big"] = {"bases": n, "file_GB": round((os.path.getsize(f1) + os.path.getsize(f2)) / 1e9, 2), "ours_s": round(t, 2), "Mbases_per_s": round(n / 1e6 / t, 1),
                  "text_GB_per_s": round((os.path.getsize(f1) + os.path.getsize(f2)) / 1e9 / t, 2), "output_lines": open(o_us, "rb").read().count(b"\n"),
                  "reference_extrapolated_s": round(res["checked"]["noboringbits"]["reference_s"] / res["checked"]["bases"] * n, 1)}
    for p in (f1, f2):
        os.remove(p)
res["cores"] = os.cpu_count()
print(json.dumps(res))
for p in (o_ref, o_us, o_us1, gen):
    try:
        os.remove(p)
    except OSError:
        pass
