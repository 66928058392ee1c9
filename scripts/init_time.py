#!/usr/bin/env python3
"""Where the start-up time of a fresh process goes (ctypes, no torch): driver init, context, first launches."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = C.CDLL(os.path.join(ROOT, "cornetto_b200", "lib", "libcorn_gpu.so"))
t0 = time.perf_counter()
L.corn_gpu_device_count.restype = C.c_int
n = L.corn_gpu_device_count()
t1 = time.perf_counter()
ctx = C.c_void_p()
L.corn_gpu_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
r = L.corn_gpu_init(0, C.byref(ctx))
t2 = time.perf_counter()
# first real work: a tiny telofind through the host-buffer entry point
hb = C.c_void_p()
L.corn_hbatch_create.argtypes = [C.c_uint64, C.c_uint32, C.POINTER(C.c_void_p)]
L.corn_hbatch_create(1 << 20, 16, C.byref(hb))
L.corn_hbatch_add.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
L.corn_hbatch_add(hb, b"ACGTTAGGGTTAGGGTTAGGGACGT" * 100, 2500)


class Batch(C.Structure):
    _fields_ = [("seq", C.c_void_p), ("offset", C.c_void_p), ("length", C.c_void_p), ("n_rec", C.c_uint32), ("total_bytes", C.c_uint64)]


class Hits(C.Structure):
    _fields_ = [("run", C.c_void_p), ("n_run", C.c_uint64), ("_owner", C.c_void_p)]


v, h = Batch(), Hits()
L.corn_hbatch_view.argtypes = [C.c_void_p, C.POINTER(Batch)]
L.corn_hbatch_view(hb, C.byref(v))
L.corn_gpu_telofind.argtypes = [C.c_void_p, C.POINTER(Batch), C.c_char_p, C.POINTER(Hits)]
r2 = L.corn_gpu_telofind(ctx, C.byref(v), b"TTAGGG", C.byref(h))
t3 = time.perf_counter()
r3 = L.corn_gpu_telofind(ctx, C.byref(v), b"TTAGGG", C.byref(h))
t4 = time.perf_counter()
print(f"devices {n}: device_count {t1 - t0:.3f} s, corn_gpu_init {t2 - t1:.3f} s (rc {r}), first telofind {t3 - t2:.3f} s (rc {r2}, {h.n_run} runs), second {t4 - t3:.4f} s")
