#!/usr/bin/env python3
"""bench.py -- the measurement contract of this repository.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

N = 1 (BASELINE.json configs[1], "c2"): telofind + telowin (threshold 0.4, identity 99.9) over a synthetic 3.1 Gb
T2T-like haploid assembly (24 contigs with CHM13-like lengths, (CCCTAA)n / (TTAGGG)n ends with 2 % variant
repeats, interstitial telomere blocks, microsatellites, 5 % soft-masked lower case), generated in place in HBM
from a seeded counter-based generator.  A "step" is one pass of the hot path over that assembly.

N > 1 (BASELINE.json configs[2], "c3"): ONE 6.2 Gb diploid assembly (48 contigs, N gaps included) split over the
N GPUs by the library's record partitioner (corn_shard_plan, csrc/shard.cu): one process per GPU (torchrun), each
holding its shard resident; no collective on the data path (NCCL carries the barriers and the timing reduction
only).  A step is one pass of every rank over its shard; `value` = 6.2 Gb x K / max-over-ranks time: STRONG
scaling (the same total work for every N > 1).  Inside the run every rank checks one whole contig of its shard
byte for byte against the compiled reference (telofind, telowin, sdust), and rank 0 merges the shards' results
into file order (corn_shard_merge_*) and runs telobreaks on them beside the reference's.  The N = 1 line carries
the same c3 job on one GPU (two resident batches) as `c3_1gpu`, the base of the strong-scaling curve.

  value          Gbases/s, all kernels of the step (scan + ordering + run assembly + bins + windows),
                 inputs resident in HBM, CUDA events on the launching stream, max over ranks
  roofline       dominant kernel (k_telofind_scan): algorithmic bytes = 1 byte per base + 16 bytes per
                 emitted run, divided by that kernel's event-timed duration, against the measured
                 HBM copy bandwidth in MEASURED_PEAKS.json; step_frac = the same bytes over the whole step
  e2e            same metric from FASTA TEXT in page-locked host memory through the C ABI: corn_gpu_ingest
                 (H2D of the text, parsing on the device) + corn_gpu_telofind_dev + corn_gpu_telowin, runs and
                 windows copied back -- every byte of input crosses PCIe inside the timed region
  e2e_from_file  what a user of the drop-in binary waits for: `cornetto telofind x.fa > x.telomere` +
                 `cornetto telowin x.telomere 99.9 0.4`, FASTA file (page cache) to text, process and CUDA
                 start-up included; the reference binary on the same file beside it
  cpu_baseline   the reference's own C implementation (oracle/_ref/cornetto, compiled from the
                 unmodified sources) or the oracle port, single thread as shipped, on a bounded
                 sample of the same assembly
  sdust          (extra) `sdust -w 64 -t 20` kernel time on the same resident assembly (BASELINE.json configs[3])
  ingest         (extra) device-side FASTA parsing of 0.8 GB of text: PCIe copy, line tables, gather kernel

`--impl reference` times the reference's CPU implementation alone (rank 0 only) on the host cores:
P independent processes over contig-split FASTAs, P = usable cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# T2T-CHM13v2-like chromosome lengths (bp), 1..22, X, Y
CHM13 = [248_387_328, 242_696_752, 201_105_948, 193_574_945, 182_045_439, 172_126_628, 160_567_428,
         146_259_331, 150_617_247, 134_758_134, 135_127_769, 133_324_548, 113_566_686, 101_161_492,
         99_753_195, 96_330_374, 84_276_897, 80_542_538, 61_707_364, 66_210_255, 45_090_682,
         51_324_926, 154_259_566, 62_460_029]
METRIC = "Gbases/s scanned (telofind+telowin)"
UNIT = "Gbases/s"
THR = 0.4 * (99.9 / 100.0) ** 6


def workload_lengths(name: str):
    if name == "c2":
        return list(CHM13)
    if name == "c3":                         # diploid: maternal set + a paternal set whose contigs differ by up to ~1 %
        return list(CHM13) + [L - (L // 997) * ((i * 7) % 11) for i, L in enumerate(CHM13)]
    if name == "small":                      # 1/64 scale, for quick runs and CI-sized boxes
        return [max(1000, L // 64) for L in CHM13]
    if name == "small3":
        return [max(1000, L // 64) for L in workload_lengths("c3")]
    if name == "c2q":                        # quarter scale (0.78 Gb): the size of one GPU's shard in the 8-GPU c3 run
        return [L // 4 for L in CHM13]
    raise SystemExit(f"unknown workload {name}")


def workload_names(name: str):
    n = len(CHM13)
    if name in ("c3", "small3"):
        return [f"chr{i + 1}_MATERNAL" for i in range(n)] + [f"chr{i + 1}_PATERNAL" for i in range(n)]
    return [f"chr{i + 1}" for i in range(n)]


# ------------------------------------------------------------------------------------------------
# synthetic features (host side: a few 10^5 small descriptors; the bytes are generated on the GPU)
# ------------------------------------------------------------------------------------------------
def make_features(capi, lengths, seed, n_gaps=0):
    """-> (tandem, lower, gaps) feature arrays over GLOBAL record numbers; every record draws from its own
    generator keyed by (seed, record), so a record gets the same features whichever shard holds it.
    n_gaps: N runs per contig (lengths log-uniform in 1..50 000), none inside the telomeric ends."""
    tand, lower, gaps = [], [], []

    def unit8(b):
        u = np.zeros(8, dtype=np.uint8)
        u[:len(b)] = np.frombuffer(b, dtype=np.uint8)
        return u

    for rec, L in enumerate(lengths):
        rng = np.random.default_rng([seed, rec])
        occupied = []
        if L > 40_000:
            n5, n3 = int(rng.integers(500, 2501)), int(rng.integers(500, 2501))
            tand.append((rec, 0, 6 * n5, 0, 6, int(rng.integers(1, 2**31)), 0.02, unit8(b"CCCTAA")))
            tand.append((rec, L - 6 * n3, 6 * n3, 0, 6, int(rng.integers(1, 2**31)), 0.02, unit8(b"TTAGGG")))
            occupied += [(0, 6 * n5), (L - 6 * n3, L)]
        n_ms = int(50 * L / 1e6) + 3
        starts = np.sort(rng.integers(0, max(1, L - 400), size=n_ms))
        for k, st in enumerate(starts):
            st = int(st)
            if k < 3:                       # interstitial telomere-like blocks
                unit = b"TTAGGG" if rng.random() < 0.5 else b"CCCTAA"
                ln, pv = 6 * int(rng.integers(5, 41)), 0.02
            else:                           # microsatellites / homopolymers
                period = int(rng.integers(1, 7))
                unit = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=period)])
                ln, pv = int(rng.integers(20, 301)), 0.05
            if any(st < b and st + ln > a for a, b in occupied) or st + ln > L:
                continue
            if k + 1 < n_ms and st + ln > int(starts[k + 1]):
                continue
            tand.append((rec, st, ln, 0, len(unit), int(rng.integers(1, 2**31)), pv, unit8(unit)))
        n_low = int(0.05 * L / 2000) + 1    # soft-masked stretches, ~2 kb each
        ls = np.sort(rng.integers(0, max(1, L - 4000), size=n_low))
        for k, st in enumerate(ls):
            ln = int(rng.integers(500, 3500))
            if k + 1 < n_low and st + ln > ls[k + 1]:
                ln = int(ls[k + 1] - st)
            if ln > 0:
                lower.append((rec, int(st), ln, 2, 1, 0, 0.0, unit8(b"")))
        if n_gaps and L > 200_000:
            gs = np.sort(rng.integers(60_000, L - 120_000, size=n_gaps))
            for k, st in enumerate(gs):
                ln = int(np.exp(rng.uniform(0.0, np.log(50_000.0))))
                if k + 1 < n_gaps and st + ln > gs[k + 1]:
                    ln = int(gs[k + 1] - st)
                if ln > 0:
                    gaps.append((rec, int(st), ln, 1, 1, 0, 0.0, unit8(b"")))

    def pack(rows):
        a = np.zeros(len(rows), dtype=capi.FEAT_DTYPE)
        for i, (rec, st, ln, kind, period, sd, pv, unit) in enumerate(rows):
            a[i] = (rec, st, ln, kind, period, sd, pv, unit, 0)
        return a
    return pack(tand), pack(lower), pack(gaps)


def build_resident(ctx, capi, lengths, records, seed, feats):
    """Resident batch holding the global records `records` (file order): bytes keyed by (seed, global record), then
    the features of those records.  Returns the batch handle."""
    records = np.asarray(records, dtype=np.uint32)
    db = ctx.alloc([lengths[int(r)] for r in records])
    ctx.fill_random(db, seed, rec_id=records)
    local = np.full(len(lengths), -1, dtype=np.int64)
    local[records] = np.arange(len(records))
    for f in feats:                          # tandem, then lower case, then N gaps (successive calls are ordered)
        mine = f[local[f["rec"]] >= 0].copy()
        mine["rec"] = local[mine["rec"]]
        if len(mine):
            ctx.apply_features(db, mine)
    return db


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, during the timed region)
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []          # (arrival time, fields)
        self.proc = None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 9:
                self.rows.append((time.time(), f))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        # samples that arrived inside the timed region (the sampler itself is started before the warm-up)
        inside = [f for (t, f) in self.rows if self.t0 is None or (self.t0 <= t <= (self.t1 or t) + 0.02)]
        if not inside:
            inside = [f for (_, f) in self.rows[-3:]]
        sm = [float(r[1]) for r in inside if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in inside if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# ------------------------------------------------------------------------------------------------
def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "cornetto")
    if os.path.exists(p):
        return p, "reference"
    p = os.path.join(ROOT, "oracle", "_build", "oracle_cornetto")
    if not os.path.exists(p):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return p, "port"


def write_fasta(path, named_seqs, width=60):
    with open(path, "wb") as f:
        for name, s in named_seqs:
            f.write(b">" + name.encode() + b"\n")
            a = np.asarray(s, dtype=np.uint8)
            full = len(a) // width * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = a[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if full < len(a):
                f.write(a[full:].tobytes() + b"\n")


def cpu_pipeline_seconds(binary, fasta, workdir):
    """telofind -> awk-style re-tab (scripts/telostats.sh:35) -> telowin 99.9 0.4; wall seconds of the two commands."""
    tel = os.path.join(workdir, os.path.basename(fasta) + ".telomere")
    err = None if os.environ.get("CORNETTO_TRACE") else subprocess.DEVNULL      # (phase times of the drop-in binary, for debugging)
    t0 = time.perf_counter()
    with open(tel, "wb") as out:
        subprocess.run([binary, "telofind", fasta], stdout=out, stderr=err, check=True)
    t1 = time.perf_counter()
    subprocess.run([binary, "telowin", tel, "99.9", "0.4"], stdout=subprocess.DEVNULL, stderr=err, check=True)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def bench_ingest(ctx, peak, mbases=800, width=60, n_rec=8):
    """corn_gpu_ingest on `mbases` Mb of 60-column FASTA text held in page-locked host memory: PCIe copy, line
    tables, compaction.  Algorithmic bytes of the device phase: 1 B read per text byte + 1 B written per base."""
    rng = np.random.default_rng(5)
    per = mbases * 1_000_000 // n_rec // width * width
    parts = []
    for i in range(n_rec):
        body = np.empty((per // width, width + 1), dtype=np.uint8)
        body[:, :width] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=per, dtype=np.uint8)].reshape(-1, width)
        body[:, width] = 10
        parts += [np.frombuffer(b">chr%d bench\n" % (i + 1), dtype=np.uint8), body.reshape(-1)]
    text = np.concatenate(parts)
    ctx.L.corn_gpu_host_register(text.ctypes.data, len(text))
    best = None
    for _ in range(3):
        res = ctx.ingest(text, final=True, keep_db=True)
        t = ctx.timing()
        assert not res["irregular"] and res["n_rec"] == n_rec and int(res["length"].sum()) == per * n_rec
        ctx.free(res["db"])
        if best is None or t["post_ms"] + t["scan_ms"] < best["post_ms"] + best["scan_ms"]:
            best = t
    ctx.L.corn_gpu_host_unregister(text.ctypes.data)
    dev_ms = best["post_ms"] + best["scan_ms"]
    return {"workload": f"corn_gpu_ingest on {len(text) / 1e9:.2f} GB of {width}-column FASTA text ({n_rec} records, page-locked host buffer)",
            "text_bytes": int(len(text)), "bases": int(per * n_rec), "h2d_ms": best["h2d_ms"], "tables_ms": best["post_ms"],
            "copy_kernel_ms": best["scan_ms"], "device_gbytes_per_s": len(text) / (dev_ms * 1e-3) / 1e9,
            "copy_kernel_hbm_frac": (len(text) + per * n_rec) / (best["scan_ms"] * 1e-3) / 1e9 / peak,
            "end_to_end_gbytes_per_s": len(text) / ((dev_ms + best["h2d_ms"]) * 1e-3) / 1e9,
            "bound": "PCIe (the H2D copy of the text); the device phase runs at HBM-class rates"}


def bind_to_gpu_numa_node(torch, local):
    """Runs this rank on the CPUs of the NUMA node its GPU hangs off, so that the page-locked staging buffers of
    the e2e leg are allocated (first touch) in memory local to the GPU's PCIe root.  Returns the node or None."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# the GPU generator (csrc/synth.cu), restated in numpy: the reference arm scans the SAME bytes
# ------------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return x ^ (x >> np.uint64(31))


def host_contig(length, rec_id, seed, feats):
    """Bytes of global record `rec_id` exactly as corn_bench_fill_random_rec + corn_bench_apply_features leave them
    (k_fill_random / k_apply_features): counter-based A/C/G/T keyed by (seed, record, 16-byte block), then the tandem,
    lower-case and N features of that record."""
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = np.empty(length, dtype=np.uint8)
    with np.errstate(over="ignore"):
        key = _mix64(np.uint64(seed) ^ (np.uint64(rec_id) << np.uint64(40)))
        shifts = (np.arange(16, dtype=np.uint64) * np.uint64(2))[None, :]
        step = 1 << 20                                           # blocks per slice (16 MB of bases)
        n_blk = (length + 15) // 16
        for b0 in range(0, n_blk, step):
            blk = np.arange(b0, min(n_blk, b0 + step), dtype=np.uint64)
            h = _mix64(key ^ (blk * np.uint64(0xD1342543DE82EF95)))
            codes = ((h[:, None] >> shifts) & np.uint64(3)).astype(np.uint8).reshape(-1)
            lo = b0 * 16
            hi = min(length, lo + len(codes))
            out[lo:hi] = acgt[codes[:hi - lo]]
        for f in feats:                                          # tandem, then lower case, then N gaps
            for row in f[f["rec"] == rec_id]:
                st, n = int(row["start"]), int(row["len"])
                if st >= length:
                    continue
                n = min(n, length - st)
                kind = int(row["kind"])
                if kind == 1:
                    out[st:st + n] = ord("N")
                elif kind == 2:
                    seg = out[st:st + n]
                    up = (seg >= ord("A")) & (seg <= ord("Z"))
                    seg[up] += 32
                else:
                    period = int(row["period"])
                    i = np.arange(n, dtype=np.uint64)
                    copy, k = i // np.uint64(period), (i % np.uint64(period)).astype(np.int64)
                    b = row["unit"][k].copy()
                    h = _mix64((np.uint64(int(row["seed"])) << np.uint64(32)) ^ copy)
                    frac = (h & np.uint64(0xFFFFFF)).astype(np.float32) * np.float32(1.0 / 16777216.0)
                    var = (frac < np.float32(row["p_variant"])) & (((h >> np.uint64(24)) % np.uint64(period)).astype(np.int64) == k)
                    b[var] = acgt[((h[var] >> np.uint64(40)) & np.uint64(3)).astype(np.int64)]
                    out[st:st + n] = b
    return out


def host_random_contig(rng, L):
    """numpy stand-in of the GPU generator for the reference arm (same composition, other bytes)."""
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=L, dtype=np.uint8)]
    if L > 40_000:
        n5, n3 = int(rng.integers(500, 2501)), int(rng.integers(500, 2501))
        s[:6 * n5] = np.tile(np.frombuffer(b"CCCTAA", dtype=np.uint8), n5)
        s[L - 6 * n3:] = np.tile(np.frombuffer(b"TTAGGG", dtype=np.uint8), n3)
    return s


def workload_text(wl, world):
    return {"c2": "c2: telofind(TTAGGG)+telowin(0.4, 99.9) on a synthetic 3.12 Gb T2T-like haploid assembly (24 contigs)",
            "c3": "c3: telofind(TTAGGG)+telowin(0.4, 99.9) on ONE synthetic 6.2 Gb diploid assembly (48 contigs, N gaps)"}.get(wl, wl)


def _write_part(job):
    """worker of the reference arm: one contig-split FASTA with the same bytes the GPU arm generates in HBM"""
    path, idxs, lengths, names, seed_bytes, feats = job
    write_fasta(path, [(names[j], host_contig(lengths[j], j, seed_bytes, feats)) for j in idxs])
    return path


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from cornetto_b200 import capi          # (dtype of the feature table only: the CUDA library is not loaded on this arm)
    binary, kind = ref_binary()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    wl = args.workload or ("c2" if args.gpus == 1 else "c3")
    lengths_all = workload_lengths(wl)
    names = workload_names(wl)
    feats = make_features(capi, lengths_all, 7, n_gaps=3 if wl in ("c3", "small3") else 0)
    # Same bytes as the GPU arm (host_contig restates its generator).  The whole assembly per step when the run stays
    # within a few minutes (~0.55 s per Gb on 16+ cores), else the longest contigs up to a budget -- a bounded sample.
    per_gb = 0.6
    budget_s = 150.0
    n_steps = args.steps + args.warmup
    total_all = float(sum(lengths_all))
    keep = list(range(len(lengths_all)))
    if total_all / 1e9 * per_gb * n_steps > budget_s:
        want = budget_s / n_steps / per_gb * 1e9
        order = sorted(keep, key=lambda j: -lengths_all[j])
        keep, acc = [], 0.0
        for j in order:
            if acc + lengths_all[j] <= want or not keep:
                keep.append(j); acc += lengths_all[j]
        keep.sort()
    lengths = {j: lengths_all[j] for j in keep}
    P = max(1, min(cores, len(keep)))
    with tempfile.TemporaryDirectory(prefix="corn_ref_") as td:
        # contig-split FASTAs, longest-first round robin over P processes
        order = sorted(keep, key=lambda j: -lengths_all[j])
        groups = [[] for _ in range(P)]
        for i, idx in enumerate(order):
            groups[i % P].append(int(idx))
        jobs = [(os.path.join(td, f"part{g}.fa"), idxs, lengths_all, names, 42, feats) for g, idxs in enumerate(groups)]
        import multiprocessing as mp
        with mp.Pool(P) as pool:
            files = pool.map(_write_part, jobs)
        total = float(sum(lengths.values()))

        def one_step():
            t0 = time.perf_counter()
            procs = []
            for path in files:
                cmd = (f"{binary} telofind {path} > {path}.telomere 2>/dev/null && "
                       f"{binary} telowin {path}.telomere 99.9 0.4 > /dev/null 2>&1")
                procs.append(subprocess.Popen(["bash", "-c", cmd]))
            for p in procs:
                if p.wait() != 0:
                    raise RuntimeError("reference pipeline failed")
            return time.perf_counter() - t0
        for _ in range(args.warmup):
            one_step()
        times = [one_step() for _ in range(args.steps)]
    t = sum(times)
    value = total * args.steps / t / 1e9
    sample = ("the whole assembly" if len(keep) == len(lengths_all) else f"its {len(keep)} longest contigs ({total / 1e6:.0f} Mb)") + \
             f", same bytes as the GPU arm, {P} processes over contig-split FASTAs (files in, text out)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_text(wl, args.gpus), "contigs": len(lengths_all), "bases": int(total_all)},
            "sample": sample, "processes": P,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": P, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram bytes per launch of k_telofind_scan from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "telofind_scan_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def fasta_text(named_seqs, width=60):
    """FASTA text of the records as one uint8 array (what a file of them holds)."""
    parts = []
    for name, a in named_seqs:
        a = np.asarray(a, dtype=np.uint8)
        parts.append(np.frombuffer(b">" + name.encode() + b"\n", dtype=np.uint8))
        full = len(a) // width * width
        if full:
            body = np.empty((full // width, width + 1), dtype=np.uint8)
            body[:, :width] = a[:full].reshape(-1, width)
            body[:, width] = 10
            parts.append(body.reshape(-1))
        if full < len(a):
            parts += [a[full:], np.frombuffer(b"\n", dtype=np.uint8)]
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


def fmt_telofind(runs, names, lengths):
    return b"".join(b"%s\t%d\t%d\t%d\t%d\t%d\n" % (names[r], lengths[r], st, a, b, b - a)
                    for r, st, a, b in zip(runs["rec"].tolist(), runs["strand"].tolist(), runs["start"].tolist(), runs["end"].tolist()))


def fmt_windows(wins, names, lengths):
    return b"".join(b"Window\t%s\t%d\t%d\t%d\t%s\n" % (names[r], lengths[r], a, b, (b"%.3g" % (c / (b - a))))
                    for r, a, b, c in zip(wins["rec"].tolist(), wins["start"].tolist(), wins["end"].tolist(), wins["car"].tolist()))


def fmt_sdust(iv, first, names):
    out = []
    s = (iv >> np.uint64(32)).astype(np.int64).tolist()
    f = (iv & np.uint64(0xFFFFFFFF)).astype(np.uint32).astype(np.int32).tolist()
    for r in range(len(names)):
        a, b = int(first[r]), int(first[r + 1])
        out.append(b"".join(b"%s\t%d\t%d\n" % (names[r], s[k], f[k]) for k in range(a, b)))
    return b"".join(out)


def check_contig_against_reference(ctx, capi, db, local_rec, name, length, runs, wins, iv, first, workdir):
    """One whole contig of a resident batch against the compiled reference (or the oracle port): telofind lines,
    telowin (99.9, 0.4) lines and sdust lines byte for byte.  Returns (verdicts, reference seconds for telofind and
    telowin, kind of checker)."""
    binary, kind = ref_binary()
    seq = ctx.download(db, int(local_rec), int(length))
    fa = os.path.join(workdir, f"check_{name}.fa")
    write_fasta(fa, [(name, seq)])
    tf, tw = cpu_pipeline_seconds(binary, fa, workdir)            # leaves <fa>.telomere in workdir
    tel = os.path.join(workdir, os.path.basename(fa) + ".telomere")
    want_t = open(tel, "rb").read()
    want_w = subprocess.run([binary, "telowin", tel, "99.9", "0.4"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    want_s = subprocess.run([binary, "sdust", fa], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    bname = [name.encode()]
    mine = runs[runs["rec"] == local_rec].copy()
    mine["rec"] = 0
    mw = wins[wins["rec"] == local_rec].copy()
    mw["rec"] = 0
    a, b = int(first[local_rec]), int(first[local_rec + 1])
    ok = {"telofind": fmt_telofind(mine, bname, [int(length)]) == want_t,
          "telowin": fmt_windows(mw, bname, [int(length)]) == want_w,
          "sdust": fmt_sdust(iv[a:b], np.array([0, b - a], dtype=np.uint64), bname) == want_s,
          "lines": [want_t.count(b"\n"), want_w.count(b"\n"), want_s.count(b"\n")]}
    return ok, tf, tw, kind


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("CORN_BENCH_WORKLOAD", ""),
                    help="c2 (default at N=1), c3 (default at N>1), small, small3, c2q")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sdust", action="store_true")
    ap.add_argument("--no-ingest", action="store_true")
    ap.add_argument("--no-c3", action="store_true", help="N=1: skip the 6.2 Gb c3 job on one GPU (strong-scaling base)")
    ap.add_argument("--no-cli-full", action="store_true", help="skip the drop-in binary's run on the full-size FASTA")
    ap.add_argument("--no-check", action="store_true", help="N>1: skip the per-rank contig check against the reference and telobreaks")
    ap.add_argument("--no-pipeline", action="store_true", help="one host context: step k+1 is issued only after step k's windows are on the host")
    ap.add_argument("--profile-only", action="store_true", help="warm-up + steps only (for ncu): no e2e, no CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import ctypes as C
    import torch
    from cornetto_b200 import capi
    from cornetto_b200.build import ensure_built
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ensure_built()
    dist = gloo = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        gloo = dist.new_group(backend="gloo")          # host-side gathers of result lists (outside the timed regions)
        dist.barrier()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the scan path has no CPU fallback")
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(torch, local)
    ctx = capi.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    Lc = ctx.L
    peak, peak_src = load_peak()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- workload: which records this rank holds, in how many resident batches -------------------------------
    wl = args.workload or ("c2" if world == 1 else "c3")
    sharded = world > 1
    lengths = workload_lengths(wl)
    names = workload_names(wl)
    n_bases_total = int(sum(lengths))
    seed_bytes, seed_feat = 42, 7
    feats = make_features(capi, lengths, seed_feat, n_gaps=3 if wl in ("c3", "small3") else 0)

    def plan_batches(lens, n_ranks, my_rank):
        """records of this rank, cut into batches of at most CORN_MAX_BATCH_BYTES (one batch unless a single GPU holds > 4 GiB)"""
        shard_of = capi.shard_plan(lens, n_ranks)
        mine = capi.shard_records(shard_of, my_rank)
        my_bytes = sum(int(lens[int(r)]) + 64 for r in mine)
        n_b = max(1, -(-my_bytes // 0xF0000000))
        if n_b == 1:
            return shard_of, [mine]
        sub = capi.shard_plan([lens[int(r)] for r in mine], n_b)
        return shard_of, [mine[capi.shard_records(sub, b)] for b in range(n_b)]

    shard_of, my_batches = plan_batches(lengths, world, rank)
    dbs = [build_resident(ctx, capi, lengths, recs, seed_bytes, feats) for recs in my_batches]
    my_bases = int(sum(lengths[int(r)] for recs in my_batches for r in recs))
    total_bytes = int(sum(Lc.corn_gpu_dbatch_bytes(db) for db in dbs))

    # Two host contexts on the one GPU double-buffer the steps (what the drop-in binary's pipeline does with its two
    # workers per device): while the host waits for step k's windows, step k+1's scan is already queued on the other
    # context's stream, so the GPU never idles for the host round trip between steps.  Every step is still a complete
    # pass (scan, ordering, run assembly, bins, windows, windows copied to the host) and all of them complete inside
    # the timed region; --no-pipeline issues them strictly one after the other on one context.
    ctxs = [ctx] if args.no_pipeline else [ctx, capi.Context(local)]
    wins_s, tim_s = capi.Windows(), capi.Timing()
    p_wins, p_tim = C.byref(wins_s), C.byref(tim_s)

    def issue(c, db):
        # resident batch, results stay on the device: telofind_dev(out=NULL) returns without a host sync
        rc = Lc.corn_gpu_telofind_dev(c.ctx, db, b"TTAGGG", None)
        if rc != 0:
            capi._check(c.ctx, rc, "telofind_dev")

    def finish(c):
        # the fused telowin(hits=NULL) call is the step's single synchronisation point (it brings the passing windows
        # back to pinned host memory); last_timing then covers both calls (scan_ms = k_telofind_scan, post_ms = every
        # other kernel of the step).  Plain ctypes calls on preallocated structs keep the harness's own per-step
        # overhead to a few microseconds.
        rc = Lc.corn_gpu_telowin(c.ctx, None, None, THR, p_wins)
        if rc != 0:
            capi._check(c.ctx, rc, "fused telowin")
        n = wins_s.n_win
        Lc.corn_gpu_windows_free(p_wins)
        Lc.corn_gpu_last_timing(c.ctx, p_tim)
        return tim_s.scan_ms, tim_s.post_ms, tim_s.out_bytes, n

    def run_steps(batches, steps, use):
        """`steps` passes over `batches`; returns per-step sums of (scan_ms, post_ms) and the last step's (out_bytes, n_win)"""
        items = [db for _ in range(steps) for db in batches]
        n, nb = len(items), len(batches)
        scan, post, out_b, n_w = [0.0] * steps, [0.0] * steps, [0] * steps, [0] * steps
        if len(use) > 1:
            issue(use[0], items[0])
        for i in range(n):
            if len(use) > 1:
                if i + 1 < n:
                    issue(use[(i + 1) % 2], items[i + 1])      # the next step is queued before this one's windows are waited for
            else:
                issue(use[0], items[i])
            sc, po, ob, nw = finish(use[i % len(use)])
            k = i // nb
            scan[k] += sc; post[k] += po; out_b[k] += ob; n_w[k] += nw
        return scan, post, out_b[-1], n_w[-1]

    def timed_run(batches, steps, warmup, clocks=None, use=None):
        use = use or ctxs
        run_steps(batches, max(3, warmup), use)
        barrier()
        launches0 = sum(c.total_launches() for c in use)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if clocks:
            clocks.mark_begin()
        ev0.record(stream)
        scan_ms, post_ms, out_bytes, n_win = run_steps(batches, steps, use)
        ev1.record(stream)                         # (every context's last step ended with a host sync on its own stream)
        barrier()
        if clocks:
            clocks.mark_end()
        return {"elapsed_ms": allmax(ev0.elapsed_time(ev1)), "launches": sum(c.total_launches() for c in use) - launches0,
                "scan_ms": sum(scan_ms) / len(scan_ms), "post_ms": sum(post_ms) / len(post_ms), "out_bytes": int(out_bytes), "n_win": int(n_win)}

    clocks = Clocks(local)
    clocks.start()
    res = timed_run(dbs, args.steps, args.warmup, clocks)
    clk = clocks.stop()
    elapsed_ms = res["elapsed_ms"]
    value = float(n_bases_total) * args.steps / (elapsed_ms * 1e-3) / 1e9
    # per-kernel event times are taken from steps issued on ONE context: with two contexts in flight a kernel's
    # start event is recorded while the other context's scan still holds the SMs, which would count waiting as running
    solo = res if len(ctxs) == 1 else timed_run(dbs, min(args.steps, 50), 3, use=[ctx])

    # roofline of the dominant kernel on this rank's shard (max over ranks of the kernel time: the slowest shard)
    scan_avg_ms = solo["scan_ms"]
    achieved = (my_bases + res["out_bytes"]) / (scan_avg_ms * 1e-3) / 1e9
    traffic = load_traffic() if wl == "c2" else None
    step_ms = elapsed_ms / args.steps
    roofline = {"bound": "hbm", "kernel": "k_telofind_scan", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": my_bases + res["out_bytes"], "kernel_ms": scan_avg_ms,
                "post_kernels_ms": solo["post_ms"],
                "kernel_timing": "CUDA events around the kernel on its launching stream, steps issued one at a time (no other context's work in flight)",
                "step_frac": (allsum(my_bases + res["out_bytes"]) / world) / (step_ms * 1e-3) / 1e9 / peak,
                "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                "traffic_source": traffic.get("source") if traffic else None}
    loads = None
    if sharded:
        per = [int(sum(lengths[int(r)] for r in capi.shard_records(shard_of, k))) for k in range(world)]
        loads = {"bases_per_gpu": per, "imbalance": max(per) / (sum(per) / world)}
        # the slowest rank's event-timed kernels bound the step: report them beside this rank's
        roofline["kernel_ms_max_over_ranks"] = allmax(scan_avg_ms)
        roofline["post_kernels_ms_max_over_ranks"] = allmax(solo["post_ms"])

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_text(wl, world), "contigs": len(lengths), "bases": n_bases_total,
                       "bases_this_rank": my_bases, "hbm_bytes_this_rank": total_bytes, "resident_batches_this_rank": len(dbs),
                       "l2_policy": "every shard (>= 0.78 GB) is far larger than the 126 MB L2: every step streams it from HBM",
                       "windows_found_this_rank": res["n_win"],
                       "parallelism": (f"{world} record shards (corn_shard_plan, longest-first), no collective on the data path" if sharded else "1 GPU"),
                       "shards": loads, "host_numa_node": numa_node,
                       "strong_scaling_note": "N>1 lines all scan the same 6.2 Gb (c3); the N=1 line scans c2 (3.1 Gb) per the bench contract "
                                              "and carries c3 on one GPU as c3_1gpu"},
            "roofline": roofline, "clocks": clk, "gpu_launches": int(res["launches"])}

    line["config"]["step_issue"] = ("one host context, steps strictly one after the other" if len(ctxs) == 1 else
                                    "two host contexts double-buffer the steps on one GPU: step k+1's scan is queued while the host fetches step k's windows")
    if args.profile_only:
        if rank == 0:
            print(json.dumps(line), flush=True)
        return 0
    if len(ctxs) > 1:
        n1 = min(args.steps, 50)
        line["unpipelined"] = {"ms_per_step": solo["elapsed_ms"] / n1, "value": float(n_bases_total) * n1 / (solo["elapsed_ms"] * 1e-3) / 1e9, "steps": n1,
                               "gpu_launches_per_step": solo["launches"] / n1,
                               "what": "the same steps on ONE host context (the host's wait for the windows of step k delays the launch of step k+1)"}

    td_obj = tempfile.TemporaryDirectory(prefix=f"corn_bench_{rank}_")
    td = td_obj.name

    # ---- sdust over the same resident shard (BASELINE.json configs[3]); aggregate = total bases / slowest rank ----
    sd_results = []
    if not args.no_sdust:
        sd_ms = sd_post = 0.0
        for db in dbs:
            ctx.sdust_dev(db)                      # warm-up: sizes the per-chunk interval slots
            iv, first = ctx.sdust_dev(db)
            ts = ctx.timing()
            sd_ms += ts["scan_ms"]
            sd_post += ts["post_ms"]
            sd_results.append((iv, first))
        n_iv = int(allsum(sum(len(iv) for iv, _ in sd_results)))
        sd_max = allmax(sd_ms)
        line["sdust"] = {"workload": f"sdust -w 64 -t 20 on the same {n_bases_total / 1e9:.2f} Gb assembly ({world} GPU(s))", "bases": n_bases_total,
                         "intervals": n_iv, "kernel_ms": sd_max, "post_ms": allmax(sd_post),
                         "gbases_per_s_kernel": n_bases_total / (sd_max * 1e-3) / 1e9,
                         "hbm_frac": (n_bases_total + 8 * n_iv) / world / (sd_max * 1e-3) / 1e9 / peak,
                         "bound": "instruction issue / shared memory (serial state machine per chunk), not HBM"}

    # ---- N > 1: every rank checks one whole contig against the reference; rank 0 merges and runs telobreaks ----
    if sharded and not args.no_check:
        per_batch = []
        for db in dbs:
            runs = ctx.telofind_dev(db, "TTAGGG")
            wins = ctx.telowin(THR)
            per_batch.append((runs, wins))
        recs0 = my_batches[0]
        pick = int(np.argmin([lengths[int(r)] for r in recs0]))            # shortest contig of this rank's (first) batch
        g = int(recs0[pick])
        iv0, first0 = sd_results[0] if sd_results else ctx.sdust_dev(dbs[0])
        ok, tf, tw, kind = check_contig_against_reference(ctx, capi, dbs[0], pick, names[g], lengths[g], per_batch[0][0], per_batch[0][1], iv0, first0, td)
        all_ok = all(v for k, v in ok.items() if k != "lines")
        checks = [None] * world
        dist.all_gather_object(checks, {"rank": rank, "contig": names[g], "bases": int(lengths[g]), **ok,
                                        "reference_telofind_s": tf, "reference_telowin_s": tw}, group=gloo)
        line["parity_in_run"] = {"what": "one whole contig per GPU, byte for byte against " + ("oracle/_ref/cornetto (unmodified reference)" if kind == "reference" else "the oracle port"),
                                 "all_identical": all(c["telofind"] and c["telowin"] and c["sdust"] for c in checks), "per_rank": checks}
        if not args.no_cpu_baseline:
            c0 = checks[0]
            line["cpu_baseline"] = {"value": c0["bases"] / (c0["reference_telofind_s"] + c0["reference_telowin_s"]) / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
                                    "sample": f"contig {c0['contig']} ({c0['bases'] / 1e6:.0f} Mb) of the same assembly, FASTA on tmpfs/disk, page-cache warm",
                                    "telofind_s": c0["reference_telofind_s"], "telowin_s": c0["reference_telowin_s"]}
        # gather the sparse results of every shard on rank 0 (gloo, host memory), merge into file order with the
        # library's own merge, and call breaks on the whole assembly beside the reference's telobreaks
        gathered = [None] * world
        payload = {"runs": per_batch[0][0], "iv": sd_results[0][0] if sd_results else None, "first": sd_results[0][1] if sd_results else None}
        dist.gather_object(payload, gathered if rank == 0 else None, dst=0, group=gloo)
        if rank == 0 and sd_results:
            t0 = time.perf_counter()
            runs_all = capi.shard_merge_runs([p["runs"] for p in gathered], shard_of)
            iv_all, first_all = capi.shard_merge_intervals([(p["iv"], p["first"]) for p in gathered], shard_of)
            t_merge = time.perf_counter() - t0
            bnames = [n.encode() for n in names]
            long_runs = runs_all[(runs_all["end"] - runs_all["start"]) >= 24]     # telobreaks ignores shorter runs (MIN_TEL, src/telomere_breaks.c:10,99)
            tel = os.path.join(td, "asm.telomere"); open(tel, "wb").write(fmt_telofind(long_runs, bnames, lengths))
            sdf = os.path.join(td, "asm.sdust"); open(sdf, "wb").write(fmt_sdust(iv_all, first_all, bnames))
            lens = os.path.join(td, "asm.lens"); open(lens, "wb").write(b"".join(b"%s\t%d\n" % (n, L) for n, L in zip(bnames, lengths)))
            ours = os.path.join(ROOT, "cornetto_b200", "bin", "cornetto")
            binary, kind = ref_binary()
            t0 = time.perf_counter()
            a = subprocess.run([ours, "telobreaks", lens, sdf, tel], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            t1 = time.perf_counter()
            b = subprocess.run([binary, "telobreaks", lens, sdf, tel], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            t2 = time.perf_counter()
            line["telobreaks"] = {"what": "shards merged into file order (corn_shard_merge_*), then `cornetto telobreaks` on the whole assembly (host join) beside the reference's",
                                  "runs_total": int(len(runs_all)), "runs_ge_24bp": int(len(long_runs)), "sdust_intervals": int(len(iv_all)),
                                  "merge_s": t_merge, "ours_s": t1 - t0, "reference_s": t2 - t1, "breaks": a.count(b"\n"), "identical": a == b}

    # ---- N = 1: the c3 job (6.2 Gb, two resident batches) on this one GPU: the base of the strong-scaling curve ----
    if not sharded and wl == "c2" and not args.no_c3:
        l3 = workload_lengths("c3")
        f3 = make_features(capi, l3, seed_feat, n_gaps=3)
        _, b3 = plan_batches(l3, 1, 0)
        db3 = [build_resident(ctx, capi, l3, recs, seed_bytes, f3) for recs in b3]
        r3 = timed_run(db3, min(args.steps, 50), 3)
        line["c3_1gpu"] = {"workload": "c3 (6.2 Gb diploid, 48 contigs, N gaps) on ONE GPU: the N>1 lines' job unsharded", "bases": int(sum(l3)),
                           "resident_batches": len(db3), "ms_per_step": r3["elapsed_ms"] / min(args.steps, 50),
                           "value": float(sum(l3)) * min(args.steps, 50) / (r3["elapsed_ms"] * 1e-3) / 1e9, "unit": UNIT,
                           "scan_ms": r3["scan_ms"], "post_kernels_ms": r3["post_ms"]}
        for d in db3:
            ctx.free(d)

    # ---- device-side FASTA parsing: 60-column text of a bounded sample -> resident batch ----
    if not args.no_ingest and rank == 0 and not sharded:
        line["ingest"] = bench_ingest(ctx, peak)

    # ---- e2e: FASTA text in page-locked host memory through the public C ABI: ingest (H2D + device parse) +
    #      telofind + telowin, results copied back.  Every rank does its own shard. ----
    seqs, seq_ids = [], []
    for db, recs in zip(dbs, my_batches):
        flat = ctx.download_all(db)
        off = 0
        for r in recs:
            Lr = int(lengths[int(r)])
            seqs.append((names[int(r)], flat[off:off + Lr]))
            seq_ids.append(int(r))
            off += (Lr + 1 + 31) // 32 * 32
    # the reference arm (bench.py --impl reference) builds its FASTA files with host_contig(): same bytes?  (shortest contig of this rank)
    j = int(np.argmin([len(a) for _, a in seqs]))
    same_bytes = bool(np.array_equal(seqs[j][1], host_contig(len(seqs[j][1]), seq_ids[j], seed_bytes, feats)))
    line["config"]["same_bytes_as_reference_arm"] = {"contig": seqs[j][0], "identical": same_bytes}
    for db in dbs:
        ctx.free(db)                              # make room: the e2e path builds its own resident copy
    # text blocks of at most ~3.7 GB (what one corn_gpu_ingest call takes), cut at record boundaries
    blocks, cur, cur_b = [], [], 0
    for nm, a in seqs:
        if cur and cur_b + len(a) > 3_600_000_000:
            blocks.append(cur); cur, cur_b = [], 0
        cur.append((nm, a)); cur_b += len(a)
    if cur:
        blocks.append(cur)
    texts = [fasta_text(b) for b in blocks]
    for t in texts:
        Lc.corn_gpu_host_register(t.ctypes.data, len(t))
    def e2e_step(c):
        hits = capi.Hits()
        nrun = nwin = 0
        for t in texts:
            ing = c.ingest(t, final=True, keep_db=True)
            capi._check(c.ctx, Lc.corn_gpu_telofind_dev(c.ctx, ing["db"], b"TTAGGG", C.byref(hits)), "corn_gpu_telofind_dev")
            nrun += hits.n_run
            Lc.corn_gpu_hits_free(C.byref(hits))
            nwin += len(c.telowin(THR))
            c.free(ing["db"])
        return nrun, nwin

    def e2e_run(steps, use):
        """`steps` complete passes.  With two host contexts, two host threads take the steps alternately (every ABI call
        is synchronous, ctypes releases the GIL): while one step's text is parsed and scanned, the next step's text is
        already crossing PCIe -- the same double buffering as the resident loop above, and what the drop-in binary's
        two workers per device do with file blocks."""
        if len(use) == 1:
            res = [e2e_step(use[0]) for _ in range(steps)]
            return res[-1]
        import threading
        nxt, res, errs, mu = [0], [None] * steps, [], threading.Lock()

        def worker(c):
            try:
                while True:
                    with mu:
                        i = nxt[0]; nxt[0] += 1
                    if i >= steps:
                        return
                    res[i] = e2e_step(c)
            except BaseException as e:          # noqa: BLE001 (re-raised on the main thread)
                errs.append(e)
        th = [threading.Thread(target=worker, args=(c,)) for c in use]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        return res[-1]

    def e2e_timed(steps, use):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(stream)
        nrun, nwin = e2e_run(steps, use)
        ev1.record(stream)                      # (every step ended with a host sync: the device is idle here)
        barrier()
        t_wall = time.perf_counter() - t0
        return allmax(ev0.elapsed_time(ev1)), allmax(t_wall * 1e3), nrun, nwin

    for c in ctxs:
        e2e_step(c)                             # warm-up: buffers of both contexts sized
    one_ms, one_wall, nrun, nwin = e2e_timed(args.e2e_steps, ctxs[:1])
    e2e_ms, e2e_wall, e2e_steps = one_ms, one_wall, args.e2e_steps
    if len(ctxs) > 1:
        e2e_steps = 2 * args.e2e_steps
        e2e_ms, e2e_wall, nrun, nwin = e2e_timed(e2e_steps, ctxs)
    text_bytes = int(sum(len(t) for t in texts))
    line["e2e"] = {"value": float(n_bases_total) * e2e_steps / (e2e_ms * 1e-3) / 1e9, "unit": UNIT,
                   "h2d_bytes_per_step": int(allsum(text_bytes)), "d2h_bytes_per_step": int(allsum(int(nrun) * 16 + int(nwin) * 16)),
                   "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "host_wall_ms_per_step": e2e_wall / e2e_steps,
                   "host_contexts": len(ctxs),
                   "one_context": {"value": float(n_bases_total) * args.e2e_steps / (one_ms * 1e-3) / 1e9, "ms_per_step": one_ms / args.e2e_steps, "steps": args.e2e_steps},
                   "what": "FASTA text (60 columns) in page-locked host memory -> corn_gpu_ingest (PCIe + device parse) -> corn_gpu_telofind_dev -> corn_gpu_telowin -> runs + windows on the host"
                           + ("; two host contexts take the steps alternately (the next step's H2D copy overlaps this step's parse + scan)" if len(ctxs) > 1 else ""),
                   "bound": "PCIe: the H2D copy of the text (1.02 bytes per base) dominates; parse + scan run ~30x faster"}
    for t in texts:
        Lc.corn_gpu_host_unregister(t.ctypes.data)

    # ---- CPU baseline + the drop-in binary from file to text (N = 1) ----
    if not args.no_cpu_baseline and rank == 0 and not sharded:
        binary, kind = ref_binary()
        sample, got = [], 0
        budget = 600_000_000 if wl == "c2" else 10**12
        for nm, a in seqs:
            if got < budget and (got + len(a) <= budget or not sample):
                sample.append((nm, a))
                got += len(a)
        ours = os.path.join(ROOT, "cornetto_b200", "bin", "cornetto")
        fa = os.path.join(td, "sample.fa")
        write_fasta(fa, sample)
        tf, tw = cpu_pipeline_seconds(binary, fa, td)
        # the drop-in binary on the same file, same two commands (process start, CUDA start-up, file read,
        # device-side parsing, scan and text output all inside the wall clock)
        runs3 = [cpu_pipeline_seconds(ours, fa, td) for _ in range(3)]      # CUDA start-up of a fresh process is noisy: best of 3
        of, ow = min(r[0] for r in runs3), min(r[1] for r in runs3)
        cli = {"sample": {"bases": int(got), "telofind_s": of, "telowin_s": ow, "gbases_per_s": got / (of + ow) / 1e9,
                          "reference_telofind_s": tf, "reference_telowin_s": tw}}
        line["cpu_baseline"] = {"value": got / (tf + tw) / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": f"{len(sample)} contigs ({got / 1e6:.0f} Mb) of the same assembly, FASTA on tmpfs/disk, page-cache warm",
                                "telofind_s": tf, "telowin_s": tw}
        if not args.no_cli_full:
            full = os.path.join(td, "full.fa")
            write_fasta(full, seqs)
            runs2 = [cpu_pipeline_seconds(ours, full, td) for _ in range(2)]
            ff, fw = min(r[0] for r in runs2), min(r[1] for r in runs2)
            cli["full"] = {"bases": int(n_bases_total), "fasta_bytes": os.path.getsize(full), "telofind_s": ff, "telowin_s": fw,
                           "gbases_per_s": n_bases_total / (ff + fw) / 1e9}
            # the same product from ONE process: `cornetto telostats full.fa` (scripts/telostats.sh fused: device parse,
            # telofind, telowin on the resident runs, merged windows at the contig ends, tally) -- and its files must be
            # the two commands' files
            def telostats_once():
                t0 = time.perf_counter()
                subprocess.run([ours, "telostats", "full.fa"], cwd=td, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
                return time.perf_counter() - t0
            ts = min(telostats_once() for _ in range(2))
            subprocess.run(f"{ours} telowin {td}/full.fa.telomere 99.9 0.4 > {td}/two.windows 2>/dev/null", shell=True, check=True)
            same = (open(os.path.join(td, "tmp_full_telostats", "full.telomere"), "rb").read() == open(os.path.join(td, "full.fa.telomere"), "rb").read()
                    and open(os.path.join(td, "tmp_full_telostats", "full.windows.0.4"), "rb").read() == open(os.path.join(td, "two.windows"), "rb").read())
            line["e2e_from_file"] = {"value": n_bases_total / ts / 1e9, "unit": UNIT,
                                     "what": "`cornetto telostats full.fa`: FASTA file (page cache) -> .telomere, .lens, .windows.0.4, merged BED, contig-end BED and tally "
                                             "(everything scripts/telostats.sh leaves behind) from one process, process and CUDA start-up included; best of 2",
                                     "telostats_s": ts, "files_identical_to_the_two_commands": bool(same),
                                     "two_commands": {"what": "`cornetto telofind full.fa > x.telomere` + `cornetto telowin x.telomere 99.9 0.4`, best of 2",
                                                      "telofind_s": ff, "telowin_s": fw, "value": n_bases_total / (ff + fw) / 1e9},
                                     "reference_same_sample_gbases_per_s": got / (tf + tw) / 1e9,
                                     "bound": "CUDA start-up of a fresh process, the read(2) of 3.2 GB, the staged H2D copy and text formatting; kernels are <1 % of it"}
        cli["what"] = ("wall clock of the drop-in `cornetto telofind` + `cornetto telowin` commands on FASTA files (parse included), "
                       "bound by CUDA start-up (0.4-2 s per process on these boxes), the file read and text formatting")
        line["cli"] = cli
    if rank == 0:
        print(json.dumps(line), flush=True)
    td_obj.cleanup()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
