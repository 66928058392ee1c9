# Collects the per-round evidence under gpurun_out/ (copy what should be judged into profiles/).
# usage: bash scripts/evidence.sh <tag>      e.g. r02_s        (one B200; ~12 minutes)
tag=${1:-r02_x}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${tag}_gpu_tests.log
python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' >> gpurun_out/${tag}_gpu_tests.log 2>&1
python bench.py --steps 100 --warmup 5 > gpurun_out/${tag}_bench_c2.jsonl 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference_arm.jsonl 2>> gpurun_out/${tag}_bench.err
# launch list of the fused step (one host context, so that the six launches of a step are in order)
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${tag}_launches_bench_c2.csv python bench.py --profile-only --no-pipeline --steps 3 --warmup 3 > /dev/null 2>&1
# full captures: the dominant kernel, and the three kernels of the two-phase sdust path
ncu --set full --clock-control none --import-source on -k regex:k_telofind_scan -s 3 -c 1 -o gpurun_out/scan_full -f python bench.py --profile-only --no-pipeline --steps 2 --warmup 3 > /dev/null 2>&1
ncu -i gpurun_out/scan_full.ncu-rep --page raw --csv > gpurun_out/scan_full_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/scan_full_raw.csv gpurun_out/${tag}_telofind_scan_ncu_full.json "ncu --set full --clock-control none, one launch of k_telofind_scan<6,TTAGGG> on the c2 workload (3.117 Gb resident)"
for k in k_sdust_scout k_sdust_dense k_sdust_scan; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/sd_$k -f python scripts/prof_sdust.py 1200 2 > /dev/null 2>&1
  ncu -i gpurun_out/sd_$k.ncu-rep --page raw --csv > gpurun_out/sd_${k}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/sd_${k}_raw.csv gpurun_out/${tag}_${k}_ncu_full.json "ncu --set full --clock-control none, second launch of $k, sdust -w 64 -t 20 on a 1.2 Gb feature-rich batch"
done
ncu --set full --clock-control none --import-source on -k regex:k_depth_windows -s 1 -c 1 -o gpurun_out/depth_full -f python scripts/prof_depthwin.py 1000 2 > /dev/null 2>&1
ncu -i gpurun_out/depth_full.ncu-rep --page raw --csv > gpurun_out/depth_full_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/depth_full_raw.csv gpurun_out/${tag}_k_depth_windows_ncu_full.json "ncu --set full --clock-control none, second launch of k_depth_windows<16> on 1.0 G bases x 2 uint16 arrays, window 2500 / 50"
python scripts/prof_depthwin.py 1000 3 > gpurun_out/${tag}_prof_depthwin.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${tag}_launches_sdust.csv python scripts/prof_sdust.py 1200 2 > /dev/null 2>&1
CORNETTO_TRACE=1 python scripts/prof_sdust.py 3000 3 > gpurun_out/${tag}_prof_sdust.txt 2>&1
CORNETTO_TRACE=1 python scripts/prof_sdust.py 3000 2 plain >> gpurun_out/${tag}_prof_sdust.txt 2>&1
CORNETTO_SDUST_CLASSIC=1 python scripts/prof_sdust.py 3000 2 >> gpurun_out/${tag}_prof_sdust.txt 2>&1
python scripts/cli_bench.py 3000 > gpurun_out/${tag}_cli_bench_fasta_3000Mb.json 2>> gpurun_out/${tag}_bench.err
python scripts/bits_cli_bench.py 40 300 > gpurun_out/${tag}_bits_cli.json 2>> gpurun_out/${tag}_bench.err
rm -f gpurun_out/*.ncu-rep
tail -c 300 gpurun_out/${tag}_bench_c2.jsonl; cat gpurun_out/${tag}_gpu_tests.log gpurun_out/${tag}_prof_sdust.txt gpurun_out/${tag}_cli_bench_fasta_3000Mb.json
