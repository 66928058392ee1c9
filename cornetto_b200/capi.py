"""ctypes binding of include/corn_gpu.h and include/corn_bench.h (tests / bench plumbing).

Mirrors the C structs one to one; nothing here computes anything.  Loading fails loudly if the
shared library is missing -- there is no Python or CPU substitute for the kernels.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import bench_lib_path, lib_path

CORN_ALIGN = 32
CORN_OK = 0
CORN_E_NOGPU = -1

RUN_DTYPE = np.dtype([("rec", "<u4"), ("strand", "<u4"), ("start", "<u4"), ("end", "<u4")])
WIN_DTYPE = np.dtype([("rec", "<u4"), ("start", "<u4"), ("end", "<u4"), ("car", "<u4")])
DEPTHWIN_DTYPE = np.dtype([("ctg", "<u4"), ("st", "<u4"), ("end", "<u4"), ("depth", "<i4"), ("mq_depth", "<i4")])
FEAT_DTYPE = np.dtype([("rec", "<u4"), ("start", "<u4"), ("len", "<u4"), ("kind", "<u4"), ("period", "<u4"),
                       ("seed", "<u4"), ("p_variant", "<f4"), ("unit", "u1", (8,)), ("_pad", "<u4")])


class Batch(C.Structure):
    _fields_ = [("seq", C.c_void_p), ("offset", C.c_void_p), ("length", C.c_void_p),
                ("n_rec", C.c_uint32), ("total_bytes", C.c_uint64)]


class Hits(C.Structure):
    _fields_ = [("run", C.c_void_p), ("n_run", C.c_uint64), ("_owner", C.c_void_p)]


class Contigs(C.Structure):
    _fields_ = [("length", C.c_void_p), ("n", C.c_uint32)]


class Windows(C.Structure):
    _fields_ = [("win", C.c_void_p), ("n_win", C.c_uint64), ("_owner", C.c_void_p)]


class Intervals(C.Structure):
    _fields_ = [("iv", C.c_void_p), ("rec_first", C.c_void_p), ("n_iv", C.c_uint64), ("n_rec", C.c_uint32),
                ("_owner", C.c_void_p)]


class Ingest(C.Structure):
    _fields_ = [("db", C.c_void_p), ("n_rec", C.c_uint32), ("hdr_off", C.c_void_p), ("length", C.c_void_p),
                ("consumed", C.c_uint64), ("irregular", C.c_int), ("_owner", C.c_void_p)]


class DepthBatch(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("mq_depth", C.c_void_p), ("offset", C.c_void_p), ("length", C.c_void_p),
                ("n_ctg", C.c_uint32), ("n_total", C.c_uint64)]


class DepthParams(C.Structure):
    _fields_ = [("window_size", C.c_int), ("window_inc", C.c_int), ("thresh_low_depth", C.c_int), ("thresh_high_depth", C.c_int),
                ("low_mq_cov_thresh", C.c_float), ("edge_len", C.c_int), ("min_ctg_len", C.c_int), ("boring", C.c_int)]


class DepthWindows(C.Structure):
    _fields_ = [("win", C.c_void_p), ("n_win", C.c_uint64), ("_owner", C.c_void_p)]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("scan_ms", C.c_float), ("post_ms", C.c_float), ("d2h_ms", C.c_float),
                ("launches", C.c_uint32), ("out_bytes", C.c_uint64)]


# every symbol the two headers declare (tests/test_abi.py checks the library exports them all)
SYMBOLS = [
    "corn_gpu_device_count", "corn_gpu_init", "corn_gpu_destroy", "corn_gpu_strerror", "corn_gpu_last_error",
    "corn_gpu_set_stream",
    "corn_hbatch_create", "corn_hbatch_pin", "corn_hbatch_destroy", "corn_hbatch_reset", "corn_hbatch_room", "corn_hbatch_cursor",
    "corn_hbatch_commit", "corn_hbatch_add", "corn_hbatch_view",
    "corn_gpu_upload", "corn_gpu_dbatch_free", "corn_gpu_dbatch_seq_ptr", "corn_gpu_dbatch_bytes",
    "corn_gpu_dbatch_alloc", "corn_gpu_dbatch_download",
    "corn_gpu_telofind", "corn_gpu_telofind_dev", "corn_gpu_hits_free",
    "corn_gpu_telowin", "corn_gpu_windows_free",
    "corn_gpu_sdust", "corn_gpu_sdust_dev", "corn_gpu_intervals_free",
    "sdust", "sdust_buf_init", "sdust_buf_destroy", "sdust_core",
    "corn_gpu_depthwin", "corn_gpu_depth_windows_free",
    "corn_gpu_ingest", "corn_gpu_ingest_free", "corn_gpu_host_register", "corn_gpu_host_unregister",
    "corn_gpu_last_timing", "corn_gpu_total_launches",
    "corn_shard_plan", "corn_shard_local_index", "corn_shard_merge_runs", "corn_shard_merge_intervals",
    "corn_bench_fill_random", "corn_bench_fill_random_rec", "corn_bench_apply_features", "corn_bench_download_all", "corn_bench_flush_l2",
]
BENCH_SYMBOLS = [s for s in SYMBOLS if s.startswith("corn_bench_")]       # exported by libcorn_bench.so, not by the product library

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA library is the only implementation; there is no fallback)")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    # the synthetic-input helpers (include/corn_bench.h) live in a library of their own; their entry points are
    # attached to the same handle object so that callers keep writing L.corn_bench_*
    if os.path.exists(bench_lib_path()):
        B = C.CDLL(bench_lib_path())
        for name in BENCH_SYMBOLS:
            setattr(L, name, getattr(B, name))
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.corn_gpu_device_count.restype = i32
    L.corn_gpu_init.argtypes = [i32, C.POINTER(vp)]
    L.corn_gpu_destroy.argtypes = [vp]
    L.corn_gpu_destroy.restype = None
    L.corn_gpu_strerror.argtypes = [i32]
    L.corn_gpu_strerror.restype = C.c_char_p
    L.corn_gpu_last_error.argtypes = [vp]
    L.corn_gpu_last_error.restype = C.c_char_p
    L.corn_gpu_set_stream.argtypes = [vp, vp]
    L.corn_hbatch_create.argtypes = [u64, u32, C.POINTER(vp)]
    L.corn_hbatch_pin.argtypes = [vp]
    L.corn_hbatch_destroy.argtypes = [vp]
    L.corn_hbatch_destroy.restype = None
    L.corn_hbatch_reset.argtypes = [vp]
    L.corn_hbatch_reset.restype = None
    L.corn_hbatch_room.argtypes = [vp]
    L.corn_hbatch_room.restype = u64
    L.corn_hbatch_cursor.argtypes = [vp]
    L.corn_hbatch_cursor.restype = vp
    L.corn_hbatch_commit.argtypes = [vp, u64]
    L.corn_hbatch_add.argtypes = [vp, vp, u64]
    L.corn_hbatch_view.argtypes = [vp, C.POINTER(Batch)]
    L.corn_hbatch_view.restype = None
    L.corn_gpu_upload.argtypes = [vp, C.POINTER(Batch), C.POINTER(vp)]
    L.corn_gpu_dbatch_free.argtypes = [vp, vp]
    L.corn_gpu_dbatch_free.restype = None
    L.corn_gpu_dbatch_seq_ptr.argtypes = [vp]
    L.corn_gpu_dbatch_seq_ptr.restype = vp
    L.corn_gpu_dbatch_bytes.argtypes = [vp]
    L.corn_gpu_dbatch_bytes.restype = u64
    L.corn_gpu_dbatch_alloc.argtypes = [vp, vp, u32, C.POINTER(vp)]
    L.corn_gpu_dbatch_download.argtypes = [vp, vp, u32, vp]
    L.corn_gpu_telofind.argtypes = [vp, C.POINTER(Batch), C.c_char_p, C.POINTER(Hits)]
    L.corn_gpu_telofind_dev.argtypes = [vp, vp, C.c_char_p, C.POINTER(Hits)]
    L.corn_gpu_hits_free.argtypes = [C.POINTER(Hits)]
    L.corn_gpu_hits_free.restype = None
    L.corn_gpu_telowin.argtypes = [vp, C.POINTER(Hits), C.POINTER(Contigs), C.c_double, C.POINTER(Windows)]
    L.corn_gpu_windows_free.argtypes = [C.POINTER(Windows)]
    L.corn_gpu_windows_free.restype = None
    L.corn_gpu_sdust.argtypes = [vp, C.POINTER(Batch), i32, i32, C.POINTER(Intervals)]
    L.corn_gpu_sdust_dev.argtypes = [vp, vp, i32, i32, C.POINTER(Intervals)]
    L.corn_gpu_intervals_free.argtypes = [C.POINTER(Intervals)]
    L.corn_gpu_intervals_free.restype = None
    # the reference's own library interface (src/sdust/sdust.h:16-21)
    L.sdust.argtypes = [vp, C.c_char_p, i32, i32, i32, C.POINTER(i32)]
    L.sdust.restype = C.POINTER(C.c_uint64)
    L.sdust_buf_init.argtypes = [vp]
    L.sdust_buf_init.restype = vp
    L.sdust_buf_destroy.argtypes = [vp]
    L.sdust_buf_destroy.restype = None
    L.sdust_core.argtypes = [C.c_char_p, i32, i32, i32, C.POINTER(i32), vp]
    L.sdust_core.restype = C.POINTER(C.c_uint64)
    L.corn_gpu_depthwin.argtypes = [vp, C.POINTER(DepthBatch), C.POINTER(DepthParams), C.POINTER(DepthWindows)]
    L.corn_gpu_depth_windows_free.argtypes = [C.POINTER(DepthWindows)]
    L.corn_gpu_depth_windows_free.restype = None
    L.corn_gpu_ingest.argtypes = [vp, vp, u64, i32, C.POINTER(Ingest)]
    L.corn_gpu_ingest_free.argtypes = [C.POINTER(Ingest)]
    L.corn_gpu_ingest_free.restype = None
    L.corn_gpu_host_register.argtypes = [vp, u64]
    L.corn_gpu_host_unregister.argtypes = [vp]
    L.corn_gpu_host_unregister.restype = None
    L.corn_gpu_last_timing.argtypes = [vp, C.POINTER(Timing)]
    L.corn_gpu_total_launches.argtypes = [vp]
    L.corn_gpu_total_launches.restype = u64
    L.corn_shard_plan.argtypes = [vp, u32, u32, vp]
    L.corn_shard_local_index.argtypes = [vp, u32, u32, vp, vp]
    L.corn_shard_merge_runs.argtypes = [vp, vp, vp, u32, u32, vp]
    L.corn_shard_merge_intervals.argtypes = [vp, vp, vp, u32, u32, vp, vp]
    L.corn_bench_fill_random.argtypes = [vp, vp, u64]
    L.corn_bench_fill_random_rec.argtypes = [vp, vp, u64, vp]
    L.corn_bench_apply_features.argtypes = [vp, vp, vp, u32]
    L.corn_bench_download_all.argtypes = [vp, vp, vp]
    L.corn_bench_flush_l2.argtypes = [vp]
    _lib = L
    return L


class CornError(RuntimeError):
    pass


# ---- record sharding over several GPUs (csrc/shard.cu; pure host arithmetic, no device needed) -----------
def shard_plan(lengths, n_shards: int) -> np.ndarray:
    """corn_shard_plan: shard_of[r] for every record (longest first onto the least loaded shard)."""
    L = load()
    lens = np.ascontiguousarray(lengths, dtype=np.uint32)
    out = np.zeros(len(lens), dtype=np.uint32)
    _check(None, L.corn_shard_plan(lens.ctypes.data if len(lens) else None, len(lens), n_shards, out.ctypes.data if len(lens) else None), "corn_shard_plan")
    return out


def shard_records(shard_of, shard: int) -> np.ndarray:
    """global record numbers of one shard, in file order (= the order of the records inside its batch)."""
    return np.flatnonzero(np.asarray(shard_of) == shard).astype(np.uint32)


def shard_merge_runs(per_shard_runs, shard_of) -> np.ndarray:
    """corn_shard_merge_runs: per-shard run lists (rec = index inside the shard) -> one list in file order."""
    L = load()
    shard_of = np.ascontiguousarray(shard_of, dtype=np.uint32)
    parts = [np.ascontiguousarray(r, dtype=RUN_DTYPE) for r in per_shard_runs]
    n = np.array([len(p) for p in parts], dtype=np.uint64)
    ptrs = (C.c_void_p * len(parts))(*[p.ctypes.data if len(p) else None for p in parts])
    out = np.zeros(int(n.sum()), dtype=RUN_DTYPE)
    _check(None, L.corn_shard_merge_runs(ptrs, n.ctypes.data, shard_of.ctypes.data if len(shard_of) else None, len(shard_of), len(parts),
                                         out.ctypes.data if len(out) else None), "corn_shard_merge_runs")
    return out


def shard_merge_intervals(per_shard, shard_of):
    """corn_shard_merge_intervals: per-shard (iv, rec_first) -> (iv, rec_first) in file order."""
    L = load()
    shard_of = np.ascontiguousarray(shard_of, dtype=np.uint32)
    ivs = [np.ascontiguousarray(a, dtype=np.uint64) for a, _ in per_shard]
    firsts = [np.ascontiguousarray(b, dtype=np.uint64) for _, b in per_shard]
    p_iv = (C.c_void_p * len(ivs))(*[a.ctypes.data if len(a) else None for a in ivs])
    p_first = (C.c_void_p * len(firsts))(*[b.ctypes.data for b in firsts])
    out = np.zeros(sum(len(a) for a in ivs), dtype=np.uint64)
    first = np.zeros(len(shard_of) + 1, dtype=np.uint64)
    _check(None, L.corn_shard_merge_intervals(p_iv, p_first, shard_of.ctypes.data if len(shard_of) else None, len(shard_of), len(ivs),
                                              out.ctypes.data if len(out) else None, first.ctypes.data), "corn_shard_merge_intervals")
    return out, first


def _check(ctx, rc, what):
    if rc != CORN_OK:
        L = load()
        detail = L.corn_gpu_last_error(ctx).decode() if ctx else ""
        raise CornError(f"{what}: {L.corn_gpu_strerror(rc).decode()} [{rc}] {detail}")


def _copy_struct_array(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_uint8 * (n * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


class HostBatch:
    """A batch in the CORN_ALIGN layout, built with the library's own builder (pinned when a GPU exists)."""

    def __init__(self, records, capacity=None):
        L = load()
        records = [bytes(r) if not isinstance(r, np.ndarray) else r.tobytes() for r in records]
        need = sum((len(r) + 1 + CORN_ALIGN - 1) // CORN_ALIGN * CORN_ALIGN for r in records) + CORN_ALIGN
        self.h = C.c_void_p()
        _check(None, L.corn_hbatch_create(capacity or need, max(1, len(records)), C.byref(self.h)), "corn_hbatch_create")
        for r in records:
            _check(None, L.corn_hbatch_add(self.h, r, len(r)), "corn_hbatch_add")
        self.view = Batch()
        L.corn_hbatch_view(self.h, C.byref(self.view))
        self.n_rec = len(records)
        self.lengths = np.array([len(r) for r in records], dtype=np.uint32)

    def close(self):
        if self.h:
            load().corn_hbatch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One GPU context (include/corn_gpu.h).  Raises CornError(CORN_E_NOGPU) without a device."""

    def __init__(self, device: int = -1):
        L = load()
        self.L = L
        self.ctx = C.c_void_p()
        _check(None, L.corn_gpu_init(device, C.byref(self.ctx)), "corn_gpu_init")

    def close(self):
        if self.ctx:
            self.L.corn_gpu_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        _check(self.ctx, self.L.corn_gpu_set_stream(self.ctx, C.c_void_p(cuda_stream)), "corn_gpu_set_stream")

    # ---- batches ---------------------------------------------------------------------------------
    def upload(self, hb: HostBatch):
        self.L.corn_hbatch_pin(hb.h)
        db = C.c_void_p()
        _check(self.ctx, self.L.corn_gpu_upload(self.ctx, C.byref(hb.view), C.byref(db)), "corn_gpu_upload")
        return db

    def alloc(self, lengths):
        arr = np.ascontiguousarray(lengths, dtype=np.uint32)
        db = C.c_void_p()
        _check(self.ctx, self.L.corn_gpu_dbatch_alloc(self.ctx, arr.ctypes.data, len(arr), C.byref(db)), "corn_gpu_dbatch_alloc")
        return db

    def free(self, db):
        self.L.corn_gpu_dbatch_free(self.ctx, db)

    def download(self, db, rec: int, length: int) -> np.ndarray:
        out = np.empty(length, dtype=np.uint8)
        _check(self.ctx, self.L.corn_gpu_dbatch_download(self.ctx, db, rec, out.ctypes.data), "corn_gpu_dbatch_download")
        return out

    # ---- operators -------------------------------------------------------------------------------
    def telofind(self, hb: HostBatch, motif: str = "TTAGGG") -> np.ndarray:
        self.L.corn_hbatch_pin(hb.h)
        h = Hits()
        _check(self.ctx, self.L.corn_gpu_telofind(self.ctx, C.byref(hb.view), motif.encode(), C.byref(h)), "corn_gpu_telofind")
        out = _copy_struct_array(h.run, h.n_run, RUN_DTYPE)
        self.L.corn_gpu_hits_free(C.byref(h))
        return out

    def telofind_dev(self, db, motif: str = "TTAGGG", fetch: bool = True):
        h = Hits()
        _check(self.ctx, self.L.corn_gpu_telofind_dev(self.ctx, db, motif.encode(), C.byref(h) if fetch else None), "corn_gpu_telofind_dev")
        if not fetch:
            return None
        out = _copy_struct_array(h.run, h.n_run, RUN_DTYPE)
        self.L.corn_gpu_hits_free(C.byref(h))
        return out

    def telowin(self, threshold_adj: float, runs: np.ndarray | None = None, lengths=None) -> np.ndarray:
        w = Windows()
        if runs is None:
            rc = self.L.corn_gpu_telowin(self.ctx, None, None, threshold_adj, C.byref(w))
        else:
            runs = np.ascontiguousarray(runs, dtype=RUN_DTYPE)
            lens = np.ascontiguousarray(lengths, dtype=np.uint32)
            h = Hits(runs.ctypes.data if len(runs) else None, len(runs), None)
            c = Contigs(lens.ctypes.data if len(lens) else None, len(lens))
            rc = self.L.corn_gpu_telowin(self.ctx, C.byref(h), C.byref(c), threshold_adj, C.byref(w))
        _check(self.ctx, rc, "corn_gpu_telowin")
        out = _copy_struct_array(w.win, w.n_win, WIN_DTYPE)
        self.L.corn_gpu_windows_free(C.byref(w))
        return out

    def _intervals(self, iv: Intervals):
        n_rec = iv.n_rec
        first = np.frombuffer((C.c_uint8 * (8 * (n_rec + 1))).from_address(iv.rec_first), dtype=np.uint64).copy() if iv.rec_first else np.zeros(n_rec + 1, np.uint64)
        vals = np.frombuffer((C.c_uint8 * (8 * iv.n_iv)).from_address(iv.iv), dtype=np.uint64).copy() if iv.n_iv else np.zeros(0, np.uint64)
        self.L.corn_gpu_intervals_free(C.byref(iv))
        return vals, first

    def sdust(self, hb: HostBatch, T: int = 20, W: int = 64):
        self.L.corn_hbatch_pin(hb.h)
        iv = Intervals()
        _check(self.ctx, self.L.corn_gpu_sdust(self.ctx, C.byref(hb.view), T, W, C.byref(iv)), "corn_gpu_sdust")
        return self._intervals(iv)

    def sdust_dev(self, db, T: int = 20, W: int = 64):
        iv = Intervals()
        _check(self.ctx, self.L.corn_gpu_sdust_dev(self.ctx, db, T, W, C.byref(iv)), "corn_gpu_sdust_dev")
        return self._intervals(iv)

    def ingest(self, text, final: bool = True, keep_db: bool = False, pin: bool = False):
        """corn_gpu_ingest on a block of FASTA/FASTQ text (bytes or a uint8 array).  Returns a dict: irregular,
        consumed, hdr_off, length and -- unless keep_db -- the record bytes (`seq`, a list) downloaded again;
        with keep_db the resident batch handle is returned as `db` (free with free_dbatch)."""
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else np.ascontiguousarray(text, dtype=np.uint8)
        if pin and len(buf):
            self.L.corn_gpu_host_register(buf.ctypes.data, len(buf))
        ing = Ingest()
        try:
            _check(self.ctx, self.L.corn_gpu_ingest(self.ctx, buf.ctypes.data if len(buf) else None, len(buf), int(final), C.byref(ing)),
                   "corn_gpu_ingest")
        finally:
            if pin and len(buf):
                self.L.corn_gpu_host_unregister(buf.ctypes.data)
        n = int(ing.n_rec)
        res = {"irregular": bool(ing.irregular), "consumed": int(ing.consumed), "n_rec": n,
               "hdr_off": np.ctypeslib.as_array(C.cast(ing.hdr_off, C.POINTER(C.c_uint64)), (n,)).copy() if n else np.zeros(0, np.uint64),
               "length": np.ctypeslib.as_array(C.cast(ing.length, C.POINTER(C.c_uint32)), (n,)).copy() if n else np.zeros(0, np.uint32)}
        db = ing.db
        self.L.corn_gpu_ingest_free(C.byref(ing))
        if keep_db:
            res["db"] = db
            return res
        if db:
            flat = self.download_all(db)
            off = 0
            seqs = []
            for ln in res["length"]:
                ln = int(ln)
                seqs.append(flat[off:off + ln].tobytes())
                span = (ln + 1 + 31) // 32 * 32
                assert not flat[off + ln:off + span].any(), "padding not zero"
                off += span
            assert off == len(flat)
            res["seq"] = seqs
            self.free(db)
        else:
            res["seq"] = []
        return res

    def depthwin(self, depth_list, mq_list, window_size=2500, window_inc=50, lo=0, hi=1 << 30, mq_thr=0.4, edge_len=100000, min_ctg_len=1000000, boring=0):
        """corn_gpu_depthwin on per-contig uint16 arrays -> structured array of the selected windows"""
        d = np.ascontiguousarray(np.concatenate(depth_list), dtype=np.uint16)
        q = np.ascontiguousarray(np.concatenate(mq_list), dtype=np.uint16)
        lens = np.array([len(x) for x in depth_list], dtype=np.uint32)
        off = np.concatenate([[0], np.cumsum(lens[:-1], dtype=np.uint64)]).astype(np.uint64)
        b = DepthBatch(d.ctypes.data, q.ctypes.data, off.ctypes.data, lens.ctypes.data, len(lens), len(d))
        p = DepthParams(window_size, window_inc, lo, hi, mq_thr, edge_len, min_ctg_len, boring)
        w = DepthWindows()
        _check(self.ctx, self.L.corn_gpu_depthwin(self.ctx, C.byref(b), C.byref(p), C.byref(w)), "corn_gpu_depthwin")
        out = _copy_struct_array(w.win, w.n_win, DEPTHWIN_DTYPE)
        self.L.corn_gpu_depth_windows_free(C.byref(w))
        return out

    def timing(self) -> dict:
        t = Timing()
        self.L.corn_gpu_last_timing(self.ctx, C.byref(t))
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    def total_launches(self) -> int:
        return int(self.L.corn_gpu_total_launches(self.ctx))

    # ---- bench helpers (include/corn_bench.h) ------------------------------------------------------
    def fill_random(self, db, seed: int, rec_id=None):
        if rec_id is None:
            _check(self.ctx, self.L.corn_bench_fill_random(self.ctx, db, seed), "corn_bench_fill_random")
        else:
            ids = np.ascontiguousarray(rec_id, dtype=np.uint32)
            _check(self.ctx, self.L.corn_bench_fill_random_rec(self.ctx, db, seed, ids.ctypes.data), "corn_bench_fill_random_rec")

    def apply_features(self, db, feats: np.ndarray):
        feats = np.ascontiguousarray(feats, dtype=FEAT_DTYPE)
        _check(self.ctx, self.L.corn_bench_apply_features(self.ctx, db, feats.ctypes.data, len(feats)), "corn_bench_apply_features")

    def download_all(self, db) -> np.ndarray:
        n = int(self.L.corn_gpu_dbatch_bytes(db))
        out = np.empty(n, dtype=np.uint8)
        _check(self.ctx, self.L.corn_bench_download_all(self.ctx, db, out.ctypes.data), "corn_bench_download_all")
        return out

    def flush_l2(self):
        _check(self.ctx, self.L.corn_bench_flush_l2(self.ctx), "corn_bench_flush_l2")
