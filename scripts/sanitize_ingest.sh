# compute-sanitizer memcheck over the device parser through the drop-in binary (small inputs; orderly teardown)
set -u
python - <<'PY'
import sys, os
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import synth
recs = synth.assembly(31, [120_000, 40_000, 999, 33, 0, 1], n_gaps=2, iupac_per_mb=100.0, p_lower=0.05)
open("/tmp/san.fa", "wb").write(synth.fasta_bytes(recs, width=60))
open("/tmp/san1.fa", "wb").write(synth.fasta_bytes(recs, width=0))
open("/tmp/san.fq", "wb").write(synth.fastq_bytes(synth.reads(32, 300, n50=2000)))
open("/tmp/san_irr.fa", "wb").write(synth.fasta_bytes(recs[:2], width=70) + b">odd\nACGT\n+\nIIII\n>after\nTTAGGGTTAGGGTTAGGG\n")
PY
export CORNETTO_FAST_EXIT=0
for f in /tmp/san.fa /tmp/san1.fa /tmp/san.fq /tmp/san_irr.fa; do
  for bb in 0 50000; do
    if [ $bb = 0 ]; then unset CORNETTO_BATCH_BYTES; else export CORNETTO_BATCH_BYTES=$bb; fi
    compute-sanitizer --tool memcheck --log-file gpurun_out/san_tmp.log ./cornetto_b200/bin/cornetto telofind $f > /tmp/san.out 2>/dev/null
    echo "telofind $f batch=$bb: $(grep -E 'ERROR SUMMARY' gpurun_out/san_tmp.log)" >> gpurun_out/r01_san_memcheck_ingest.log
    CORNETTO_INGEST=0 ./cornetto_b200/bin/cornetto telofind $f 2>/dev/null | cmp - /tmp/san.out && echo "  output identical to the serial reader" >> gpurun_out/r01_san_memcheck_ingest.log
  done
done
unset CORNETTO_BATCH_BYTES
compute-sanitizer --tool memcheck --log-file gpurun_out/san_tmp.log ./cornetto_b200/bin/cornetto sdust /tmp/san.fa > /dev/null 2>&1
echo "sdust /tmp/san.fa: $(grep -E 'ERROR SUMMARY' gpurun_out/san_tmp.log)" >> gpurun_out/r01_san_memcheck_ingest.log
compute-sanitizer --tool racecheck --log-file gpurun_out/san_tmp.log ./cornetto_b200/bin/cornetto sdust /tmp/san.fa > /dev/null 2>&1
echo "racecheck sdust /tmp/san.fa: $(grep -E 'RACECHECK SUMMARY' gpurun_out/san_tmp.log)" >> gpurun_out/r01_san_memcheck_ingest.log
cat gpurun_out/r01_san_memcheck_ingest.log
