import os, sys, subprocess, time
import numpy as np
sys.path.insert(0, os.getcwd())
import bench
rng = np.random.default_rng(3)
fa = "/tmp/p600.fa"
bench.write_fasta(fa, [(f"chr{i+1}", bench.host_random_contig(rng, 600_000_000 // 4)) for i in range(4)])
subprocess.run(["cat", fa], stdout=subprocess.DEVNULL)
ours = "./cornetto_b200/bin/cornetto"
def go(tag):
    for i in range(2):
        t0 = time.perf_counter()
        p = subprocess.run([ours, "telofind", fa], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=dict(os.environ, CORNETTO_TRACE="1"))
        dt = time.perf_counter() - t0
        init = [l for l in p.stderr.decode().splitlines() if "corn_gpu_init" in l][:1]
        print(tag, f"{dt:.3f} s", init)
go("no parent context")
import torch
x = torch.zeros(1, device="cuda")
go("parent holds a torch context")
y = torch.empty(3_000_000_000, dtype=torch.uint8).pin_memory()
go("parent + 3 GB pinned")
z = torch.empty(20_000_000_000, dtype=torch.uint8, device="cuda")
go("parent + 3 GB pinned + 20 GB device")
