python - <<'PY'
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import bench
rng = np.random.default_rng(3)
bench.write_fasta("/tmp/t3000.fa", [(f"chr{i+1}", bench.host_random_contig(rng, 3000*1_000_000//8)) for i in range(8)])
PY
cat /tmp/t3000.fa > /dev/null
./cornetto_b200/bin/cornetto telofind /tmp/t3000.fa > /tmp/o_ref 2>/dev/null
for mbs in 0 2048 1024 512; do
  for cmd in telofind sdust; do
    for i in 1 2 3; do
      if [ $mbs = 0 ]; then unset CORNETTO_BATCH_MB; else export CORNETTO_BATCH_MB=$mbs; fi
      s=$(date +%s.%N); ./cornetto_b200/bin/cornetto $cmd /tmp/t3000.fa > /tmp/o_x 2>/tmp/err_x; e=$(date +%s.%N); echo "block=$mbs $cmd wall $(echo "$e - $s" | bc) s; $(grep "Real time" /tmp/err_x)"
    done
  done
done
