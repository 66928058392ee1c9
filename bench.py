#!/usr/bin/env python3
"""bench.py -- the measurement contract of this repository.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1], "c2"): telofind + telowin (threshold 0.4, identity 99.9) over
a synthetic 3.1 Gb T2T-like haploid assembly (24 contigs with CHM13-like lengths, (CCCTAA)n /
(TTAGGG)n ends with 2 % variant repeats, interstitial telomere blocks, microsatellites, 5 %
soft-masked lower case), generated in place in HBM from a seeded counter-based generator.
A "step" is one pass of the hot path over that assembly.  With N > 1 every rank (one process per
GPU, torchrun) scans its own 3.1 Gb assembly -- the shards are independent, there is no
collective on the data path -- so scaling is weak and `value` is the whole-job aggregate.

  value          Gbases/s, all kernels of the step (scan + ordering + run assembly + bins + windows),
                 inputs resident in HBM, CUDA events on the launching stream, max over ranks
  roofline       dominant kernel (k_telofind_scan): algorithmic bytes = 1 byte per base + 16 bytes per
                 emitted run, divided by that kernel's event-timed duration, against the measured
                 HBM copy bandwidth in MEASURED_PEAKS.json
  e2e            same metric through the host-buffer C ABI call (corn_gpu_telofind + corn_gpu_telowin):
                 H2D of the pinned sequence bytes and D2H of the runs/windows inside the timed region
  cpu_baseline   the reference's own C implementation (oracle/_ref/cornetto, compiled from the
                 unmodified sources) or the oracle port, single thread as shipped, on a bounded
                 sample of the same assembly
  sdust          (extra) `sdust -w 64 -t 20` kernel time on the same resident assembly (BASELINE.json configs[3])
  ingest         (extra) device-side FASTA parsing of 0.8 GB of text: PCIe copy, line tables, gather kernel
  cli            (extra) wall clock of the drop-in `cornetto telofind` + `cornetto telowin` commands on the
                 cpu_baseline's FASTA sample and on the whole assembly written as FASTA (parse, CUDA start-up
                 and text output included); best of a few runs

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
    if name == "small":                      # 1/64 scale, for quick runs and CI-sized boxes
        return [max(1000, L // 64) for L in CHM13]
    raise SystemExit(f"unknown workload {name}")


# ------------------------------------------------------------------------------------------------
# synthetic features (host side: a few 10^5 small descriptors; the bytes are generated on the GPU)
# ------------------------------------------------------------------------------------------------
def make_features(capi, lengths, seed):
    rng = np.random.default_rng(seed)
    tand, lower = [], []

    def unit8(b):
        u = np.zeros(8, dtype=np.uint8)
        u[:len(b)] = np.frombuffer(b, dtype=np.uint8)
        return u

    for rec, L in enumerate(lengths):
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

    def pack(rows):
        a = np.zeros(len(rows), dtype=capi.FEAT_DTYPE)
        for i, (rec, st, ln, kind, period, sd, pv, unit) in enumerate(rows):
            a[i] = (rec, st, ln, kind, period, sd, pv, unit, 0)
        return a
    return pack(tand), pack(lower)


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


def host_random_contig(rng, L):
    """numpy stand-in of the GPU generator for the reference arm (same composition, other bytes)."""
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=L, dtype=np.uint8)]
    if L > 40_000:
        n5, n3 = int(rng.integers(500, 2501)), int(rng.integers(500, 2501))
        s[:6 * n5] = np.tile(np.frombuffer(b"CCCTAA", dtype=np.uint8), n5)
        s[L - 6 * n3:] = np.tile(np.frombuffer(b"TTAGGG", dtype=np.uint8), n3)
    return s


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    binary, kind = ref_binary()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    scale = 4                                    # bounded sample: the c2 contig set at 1/4 length (779 Mb), ~0.4 s per step
    est = 0.4 * (args.steps + args.warmup)       # keep the whole run within a few minutes whatever K and W are
    if est > 150:
        scale = int(min(64, -(-4 * est // 150)))
    lengths = [L // scale for L in workload_lengths("c2")]
    P = max(1, min(cores, len(lengths)))
    rng = np.random.default_rng(1234)
    with tempfile.TemporaryDirectory(prefix="corn_ref_") as td:
        # contig-split FASTAs, longest-first round robin over P processes
        order = np.argsort(lengths)[::-1]
        groups = [[] for _ in range(P)]
        for i, idx in enumerate(order):
            groups[i % P].append(int(idx))
        files = []
        for g, idxs in enumerate(groups):
            path = os.path.join(td, f"part{g}.fa")
            write_fasta(path, [(f"chr{j + 1}", host_random_contig(rng, lengths[j])) for j in idxs])
            files.append(path)
        total = float(sum(lengths))

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
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "c2_t2t_haploid_3.1Gb telofind+telowin(0.4, 99.9)", "sample": f"24 contigs at 1/{scale} length ({total / 1e6:.0f} Mb) per step",
                       "processes": P},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": P, "kind": kind,
                             "sample": f"c2 contig set at 1/{scale} length, {P} processes over contig-split FASTAs"},
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("CORN_BENCH_WORKLOAD", "c2"))
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sdust", action="store_true")
    ap.add_argument("--no-ingest", action="store_true")
    ap.add_argument("--no-cli-full", action="store_true", help="skip the drop-in binary's run on the full-size FASTA")
    ap.add_argument("--profile-only", action="store_true", help="warm-up + steps only (for ncu): no e2e, no CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    from cornetto_b200 import capi
    from cornetto_b200.build import ensure_built
    ensure_built()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the scan path has no CPU fallback")
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(torch, local)
    ctx = capi.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    lengths = workload_lengths(args.workload)
    n_bases = int(sum(lengths))
    db = ctx.alloc(lengths)
    ctx.fill_random(db, 42 + 1000 * rank)
    tand, lower = make_features(capi, lengths, 7 + rank)
    ctx.apply_features(db, tand)
    ctx.apply_features(db, lower)
    total_bytes = int(ctx.L.corn_gpu_dbatch_bytes(db))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    import ctypes as C
    Lc = ctx.L
    wins_s, tim_s = capi.Windows(), capi.Timing()
    p_wins, p_tim = C.byref(wins_s), C.byref(tim_s)

    def step():
        # resident batch, results stay on the device: telofind_dev(out=NULL) returns without a host sync
        # and the fused telowin(hits=NULL) call is the step's single synchronisation point (it brings
        # the passing windows back to pinned host memory); last_timing then covers both calls
        # (scan_ms = k_telofind_scan, post_ms = every other kernel of the step).  Plain ctypes calls
        # on preallocated structs keep the harness's own per-step overhead to a few microseconds.
        rc = Lc.corn_gpu_telofind_dev(ctx.ctx, db, b"TTAGGG", None)
        if rc == 0:
            rc = Lc.corn_gpu_telowin(ctx.ctx, None, None, THR, p_wins)
        if rc != 0:
            capi._check(ctx.ctx, rc, "fused step")
        n = wins_s.n_win
        Lc.corn_gpu_windows_free(p_wins)
        Lc.corn_gpu_last_timing(ctx.ctx, p_tim)
        return tim_s, n

    clocks = Clocks(local)
    clocks.start()
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    launches0 = ctx.total_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms, post_ms, out_bytes, n_win = [], [], 0, 0
    barrier()
    clocks.mark_begin()
    ev0.record(stream)
    for _ in range(args.steps):
        tm, n_win = step()
        scan_ms.append(tm.scan_ms)
        post_ms.append(tm.post_ms)
        out_bytes = tm.out_bytes
    ev1.record(stream)
    barrier()
    clocks.mark_end()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = ctx.total_launches() - launches0
    clk = clocks.stop()

    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    units = torch.tensor([float(n_bases) * args.steps], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
    elapsed_ms = float(t.item())
    value = float(units.item()) / (elapsed_ms * 1e-3) / 1e9

    peak, peak_src = load_peak()
    scan_avg_ms = sum(scan_ms) / len(scan_ms)
    runs_bytes = int(out_bytes)
    achieved = (n_bases + runs_bytes) / (scan_avg_ms * 1e-3) / 1e9
    traffic = load_traffic()
    roofline = {"bound": "hbm", "kernel": "k_telofind_scan", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": n_bases + runs_bytes, "kernel_ms": scan_avg_ms,
                "post_kernels_ms": sum(post_ms) / len(post_ms),
                "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                "traffic_source": traffic.get("source") if traffic else None}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{args.workload}: telofind(TTAGGG)+telowin(0.4, 99.9) on a synthetic {n_bases / 1e9:.2f} Gb T2T-like haploid assembly per GPU",
                       "contigs": len(lengths), "bases_per_gpu": n_bases, "hbm_bytes_per_gpu": total_bytes,
                       "l2_policy": "input (3.1 GB) is far larger than the 126 MB L2: every step streams it from HBM",
                       "windows_found": n_win, "parallelism": f"{world} independent shards, no collective",
                       "host_numa_node": numa_node},
            "roofline": roofline, "clocks": clk, "gpu_launches": int(launches)}

    if args.profile_only:
        if rank == 0:
            print(json.dumps(line), flush=True)
        return 0

    # ---- sdust over the same resident assembly (BASELINE.json configs[3]); reported beside the headline ----
    if not args.no_sdust and rank == 0:
        ctx.sdust_dev(db)                      # warm-up: sizes the per-chunk interval slots
        iv, _ = ctx.sdust_dev(db)
        ts = ctx.timing()
        line["sdust"] = {"workload": f"sdust -w 64 -t 20 on the same {n_bases / 1e9:.2f} Gb assembly", "bases": n_bases, "intervals": int(len(iv)),
                         "kernel_ms": ts["scan_ms"], "post_ms": ts["post_ms"],
                         "gbases_per_s_kernel": n_bases / (ts["scan_ms"] * 1e-3) / 1e9,
                         "hbm_frac": (n_bases + 8 * len(iv)) / (ts["scan_ms"] * 1e-3) / 1e9 / peak,
                         "bound": "instruction issue / shared memory (serial state machine per chunk), not HBM"}

    # ---- device-side FASTA parsing: 60-column text of a bounded sample -> resident batch ----
    if not args.no_ingest and rank == 0:
        line["ingest"] = bench_ingest(ctx, peak)

    # ---- e2e: host buffers through the public C ABI (H2D + kernels + D2H inside the timed region) ----
    L = ctx.L
    hb = C.c_void_p()
    capi._check(None, L.corn_hbatch_create(total_bytes + 64, len(lengths), C.byref(hb)), "corn_hbatch_create")
    capi._check(None, L.corn_hbatch_pin(hb), "corn_hbatch_pin")
    cur = L.corn_hbatch_cursor(hb)
    capi._check(ctx.ctx, L.corn_bench_download_all(ctx.ctx, db, cur), "download")
    for Lr in lengths:                           # same layout rule => same offsets; padding is already zero
        capi._check(None, L.corn_hbatch_commit(hb, Lr), "commit")
    view = capi.Batch()
    L.corn_hbatch_view(hb, C.byref(view))
    hits = capi.Hits()

    def e2e_step():
        capi._check(ctx.ctx, L.corn_gpu_telofind(ctx.ctx, C.byref(view), b"TTAGGG", C.byref(hits)), "corn_gpu_telofind")
        nrun = hits.n_run
        L.corn_gpu_hits_free(C.byref(hits))
        w = ctx.telowin(THR)
        return nrun, len(w)

    ctx.free(db)                                  # make room: the e2e path uploads its own copy
    e2e_step()
    barrier()
    ev0.record(stream)
    for _ in range(args.e2e_steps):
        nrun, nwin = e2e_step()
    ev1.record(stream)
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = float(n_bases) * world * args.e2e_steps / (float(t.item()) * 1e-3) / 1e9
    line["e2e"] = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": total_bytes,
                   "d2h_bytes_per_step": int(nrun) * 16 + int(nwin) * 16, "steps": args.e2e_steps,
                   "ms_per_step": float(t.item()) / args.e2e_steps,
                   "bound": "PCIe: the pinned H2D copy of 1 byte per base dominates (the scan kernel itself runs ~100x faster)"}

    # ---- CPU baseline: the reference binary, one thread (as shipped), bounded sample ----
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        binary, kind = ref_binary()
        seq = np.frombuffer((C.c_uint8 * total_bytes).from_address(cur), dtype=np.uint8)
        sample, off, got = [], 0, 0
        budget = 600_000_000 if args.workload == "c2" else 10**12
        for i, Lr in enumerate(lengths):
            if got < budget and (got + Lr <= budget or not sample):
                sample.append((f"chr{i + 1}", seq[off:off + Lr]))
                got += Lr
            off += (Lr + 1 + 31) // 32 * 32
        ours = os.path.join(ROOT, "cornetto_b200", "bin", "cornetto")
        with tempfile.TemporaryDirectory(prefix="corn_cpu_") as td:
            fa = os.path.join(td, "sample.fa")
            write_fasta(fa, sample)
            tf, tw = cpu_pipeline_seconds(binary, fa, td)
            # the drop-in binary on the same file, same two commands (process start, CUDA start-up, file read,
            # device-side parsing, scan and text output all inside the wall clock)
            runs3 = [cpu_pipeline_seconds(ours, fa, td) for _ in range(3)]      # CUDA start-up of a fresh process is noisy: best of 3
            of, ow = min(r[0] for r in runs3), min(r[1] for r in runs3)
            cli = {"sample": {"bases": int(got), "telofind_s": of, "telowin_s": ow, "gbases_per_s": got / (of + ow) / 1e9,
                              "reference_telofind_s": tf, "reference_telowin_s": tw}}
            if not args.no_cli_full:
                full = os.path.join(td, "full.fa")
                allc, off2 = [], 0
                for i, Lr in enumerate(lengths):
                    allc.append((f"chr{i + 1}", seq[off2:off2 + Lr]))
                    off2 += (Lr + 1 + 31) // 32 * 32
                write_fasta(full, allc)
                runs2 = [cpu_pipeline_seconds(ours, full, td) for _ in range(2)]
                ff, fw = min(r[0] for r in runs2), min(r[1] for r in runs2)
                cli["full"] = {"bases": int(n_bases), "fasta_bytes": os.path.getsize(full), "telofind_s": ff, "telowin_s": fw,
                               "gbases_per_s": n_bases / (ff + fw) / 1e9}
        line["cpu_baseline"] = {"value": got / (tf + tw) / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": f"{len(sample)} contigs ({got / 1e6:.0f} Mb) of the same assembly, FASTA on tmpfs/disk, page-cache warm",
                                "telofind_s": tf, "telowin_s": tw}
        cli["what"] = ("wall clock of the drop-in `cornetto telofind` + `cornetto telowin` commands on FASTA files (parse included), "
                       "bound by CUDA start-up (0.4-2 s per process on these boxes), the file read and text formatting")
        line["cli"] = cli
    L.corn_hbatch_destroy(hb)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
