python - <<'PY'
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import bench
rng = np.random.default_rng(3)
for mb in (3000,):
    bench.write_fasta(f"/tmp/t{mb}.fa", [(f"chr{i+1}", bench.host_random_contig(rng, mb*1_000_000//8)) for i in range(8)])
PY
for mb in 3000; do
cat /tmp/t$mb.fa > /dev/null
for i in 1 2 3; do ( time CORNETTO_TRACE=1 ./cornetto_b200/bin/cornetto telofind /tmp/t$mb.fa >/tmp/o1 ) 2>&1 | grep -v "^$\|user\|sys" ; done
for i in 1 2; do ( time CORNETTO_INGEST=0 ./cornetto_b200/bin/cornetto telofind /tmp/t$mb.fa > /tmp/o2 ) 2>&1 | grep -E "real|Real" ; done
cmp /tmp/o1 /tmp/o2 && echo SAME
for i in 1 2; do ( time CORNETTO_FAST_EXIT=0 ./cornetto_b200/bin/cornetto telofind /tmp/t$mb.fa > /tmp/o2 ) 2>&1 | grep -E "real|Real" ; done
for i in 1 2; do ( time ./cornetto_b200/bin/cornetto sdust /tmp/t$mb.fa > /tmp/o2 ) 2>&1 | grep -E "real|Real" ; done
done
