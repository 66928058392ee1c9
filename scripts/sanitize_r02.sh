# compute-sanitizer over the round-2 kernels through the drop-in binary (small inputs; orderly teardown):
# two-phase sdust (scout, item table, dense + sparse item kernels, fold) on sequence with gaps, microsatellites and the
# stale-window corpus; the wide sdust instance; telostats (fused telofind + telowin: tile prefix look-back, hot-bin windows);
# the depth-window kernel (shared-memory adds).      usage: bash scripts/sanitize_r02.sh [tag=r02]
tag=${1:-r02}
log=gpurun_out/${tag}_san.log
: > $log
set -u
python - <<'PY'
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import synth
rng = np.random.default_rng(5)
recs = synth.assembly(31, [150_000, 40_000, 999, 33, 0, 1], n_gaps=3, iupac_per_mb=100.0, p_lower=0.05)
recs += [(f"ms{k}", synth.make_contig(rng, L, telo=None, n_its=0, microsat_per_mb=6000.0, n_gaps=g, gap_len=(1, 400))) for k, (L, g) in enumerate(((60_000, 2), (9_000, 0)))]
recs += synth.stale_window_records(7, 3)
open("/tmp/san2.fa", "wb").write(synth.fasta_bytes(recs, width=60))
named = [(nm, np.minimum(d, 65535), np.minimum(q, 65535)) for nm, d, q in synth.depth_arrays(3, [60_000, 2500, 2549, 7, 33_333])]
open("/tmp/san2_t.bg", "wb").write(synth.bedgraph_bytes(named, 1))
open("/tmp/san2_q.bg", "wb").write(synth.bedgraph_bytes(named, 2))
PY
export CORNETTO_FAST_EXIT=0
B=./cornetto_b200/bin/cornetto
run() {   # tool, label, command...
  tool=$1; what=$2; shift 2
  timeout 240 compute-sanitizer --tool $tool --log-file gpurun_out/san_tmp.log "$@" > /tmp/san2.out 2>/tmp/san2.err
  echo "$tool $what: exit $? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_tmp.log)" >> $log
}
run memcheck  "sdust (two-phase)"            $B sdust /tmp/san2.fa
$B sdust /tmp/san2.fa 2>/dev/null | cmp - /tmp/san2.out && echo "  same output without the tool" >> $log
run racecheck "sdust (two-phase)"            $B sdust /tmp/san2.fa
CORNETTO_SDUST_CLASSIC=1 run memcheck "sdust (chunk grid)" $B sdust /tmp/san2.fa
run memcheck  "sdust -w 200 (wide instance)" $B sdust -w 200 /tmp/san2.fa
( cd /tmp && rm -rf tmp_san2_telostats && CORNETTO_FAST_EXIT=0 timeout 240 compute-sanitizer --tool memcheck --log-file $OLDPWD/gpurun_out/san_tmp.log $OLDPWD/cornetto_b200/bin/cornetto telostats /tmp/san2.fa > /tmp/san2.out 2>/tmp/san2.err; echo "memcheck telostats: exit $? $(grep -E 'ERROR SUMMARY' $OLDPWD/gpurun_out/san_tmp.log)" >> $OLDPWD/$log )
run memcheck  "telofind"                     $B telofind /tmp/san2.fa
run racecheck "telofind"                     $B telofind /tmp/san2.fa
run memcheck  "noboringbits"                 $B noboringbits /tmp/san2_t.bg -q /tmp/san2_q.bg -m 2000 -e 100
run racecheck "noboringbits"                 $B noboringbits /tmp/san2_t.bg -q /tmp/san2_q.bg -m 2000 -e 100
run memcheck  "boringbits -w 777 -i 13"      $B boringbits /tmp/san2_t.bg -q /tmp/san2_q.bg -m 2000 -e 100 -w 777 -i 13
run memcheck  "noboringbits -i 7"            $B noboringbits /tmp/san2_t.bg -q /tmp/san2_q.bg -m 2000 -e 100 -w 100 -i 7
rm -f gpurun_out/san_tmp.log
cat $log
