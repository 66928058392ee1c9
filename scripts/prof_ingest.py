#!/usr/bin/env python3
"""Runs corn_gpu_ingest twice on synthetic 60-column FASTA text (for ncu / timing).  usage: prof_ingest.py [Mb]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cornetto_b200 import capi  # noqa: E402

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 800
ctx = capi.Context(0)
print(bench.bench_ingest(ctx, 6550.7, mbases=mb))
