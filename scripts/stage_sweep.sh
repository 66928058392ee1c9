python - <<'PY'
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import bench
rng = np.random.default_rng(3)
bench.write_fasta("/tmp/t3000.fa", [(f"chr{i+1}", bench.host_random_contig(rng, 3000*1_000_000//8)) for i in range(8)])
PY
cat /tmp/t3000.fa > /dev/null
for nt in 8 4 12 16 8 16; do
  for i in 1 2; do
    echo "threads=$nt $(CORNETTO_STAGE_THREADS=$nt CORNETTO_TRACE=1 ./cornetto_b200/bin/cornetto telofind /tmp/t3000.fa 2>&1 >/dev/null | grep -E 'corn_gpu_ingest' | sed 's/.*corn_gpu_ingest/ingest/')"
  done
done
