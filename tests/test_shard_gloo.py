"""N > 1 host logic on CPU: world_size-2 gloo run of the record sharding + ordered gather.

The scan itself needs a GPU, so here every rank runs the ORACLE on its shard (checker code used
as a stand-in for the device call -- test infrastructure only); what is under test is
the library's record sharding (csrc/shard.cu through ctypes: corn_shard_plan, corn_shard_merge_runs,
corn_shard_merge_intervals -- the code bench.py --gpus N runs on): byte-balanced assignment, one gather of the sparse
results, file-order merge.""" 
import os
import socket
import sys

import numpy as np
import pytest

from util import ROOT

WORKER = r'''
import os, sys, ctypes as C
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import synth
from cornetto_b200 import capi
from cornetto_b200.capi import RUN_DTYPE
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
recs = synth.assembly(77, [90_000, 5_000, 60_000, 0, 33, 20_000, 41_000, 7, 15_000], n_gaps=2, telo=(50, 300), microsat_per_mb=2000.0)
lengths = [len(s) for _, s in recs]
shard_of = capi.shard_plan(lengths, world)
plan = [[int(x) for x in capi.shard_records(shard_of, r)] for r in range(world)]
L = C.CDLL(os.path.join(sys.argv[1], "oracle", "_build", "liboracle.so"))
class Run(C.Structure):
    _fields_ = [("strand", C.c_uint32), ("start", C.c_uint64), ("end", C.c_uint64)]
L.orc_telofind.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.POINTER(C.POINTER(Run))]
L.orc_telofind.restype = C.c_size_t
L.orc_sdust.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
L.orc_sdust.restype = C.POINTER(C.c_uint64)
def scan(indices):
    rows, ivs, first = [], [], [0]
    for k, g in enumerate(indices):
        b = bytes(recs[g][1]); p = C.POINTER(Run)()
        n = L.orc_telofind(b, len(b), b"TTAGGG", C.byref(p))
        rows += [(k, p[i].strand, p[i].start, p[i].end) for i in range(n)]
        m = C.c_int(); q = L.orc_sdust(b, len(b), 20, 64, C.byref(m))
        ivs += [q[i] for i in range(m.value)]; first.append(len(ivs))
    return np.array(rows, dtype=RUN_DTYPE), (np.array(ivs, dtype=np.uint64), np.array(first, dtype=np.uint64))
mine = scan(plan[rank])
gathered = [None] * world
dist.gather_object(mine, gathered if rank == 0 else None, dst=0)
if rank == 0:
    runs = capi.shard_merge_runs([g[0] for g in gathered], shard_of)
    iv, first = capi.shard_merge_intervals([g[1] for g in gathered], shard_of)
    want_runs, (want_iv, want_first) = scan(list(range(len(recs))))
    assert sorted(sum(plan, [])) == list(range(len(recs)))
    loads = [sum(lengths[i] for i in p) for p in plan]
    assert max(loads) - min(loads) <= max(lengths)
    assert len(runs) == len(want_runs) and (runs == want_runs).all(), "runs differ"
    assert (iv == want_iv).all() and (first == want_first).all(), "intervals differ"
    print("OK", len(runs), len(iv))
dist.barrier()
dist.destroy_process_group()
'''


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world2_shard_and_gather(tmp_path, oracle_bin):
    import subprocess
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(w), ROOT]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0 and "OK" in p.stdout, p.stdout[-3000:]


def test_plan_is_balanced_and_complete():
    from cornetto_b200 import capi
    sys.path.insert(0, ROOT)
    import bench
    for lengths, bound in ((bench.CHM13, 1.08), (bench.workload_lengths("c3"), 1.03)):
        for world in (1, 2, 4, 8):
            shard_of = capi.shard_plan(lengths, world)
            assert (capi.shard_plan(lengths, world) == shard_of).all()                 # deterministic
            plan = [list(capi.shard_records(shard_of, r)) for r in range(world)]
            assert sorted(int(x) for p in plan for x in p) == list(range(len(lengths)))
            loads = [sum(lengths[i] for i in p) for p in plan]
            assert max(loads) / (sum(loads) / world) < bound, (world, loads)


def test_merge_rejects_unordered_lists():
    from cornetto_b200 import capi
    shard_of = capi.shard_plan([10, 20, 30], 2)
    bad = np.zeros(2, dtype=capi.RUN_DTYPE)
    bad["rec"] = [1, 0]                                                               # not in record order
    with pytest.raises(capi.CornError):
        capi.shard_merge_runs([bad, bad[:0]], shard_of)
