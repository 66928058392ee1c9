# Collects the per-round evidence under gpurun_out/ (copy what should be judged into profiles/).
# usage: bash scripts/evidence.sh <tag>      e.g. r01_q
tag=${1:-r01_x}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${tag}_gpu_tests.log
python bench.py > gpurun_out/${tag}_bench_c2.jsonl 2> gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.jsonl 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_c2.csv python bench.py --profile-only --steps 3 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_telofind_scan -s 3 -c 1 -o gpurun_out/scan_full -f python bench.py --profile-only --steps 2 --warmup 3 > /dev/null 2>&1
ncu -i gpurun_out/scan_full.ncu-rep --page raw --csv > gpurun_out/scan_full_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_sdust_scan -c 1 -o gpurun_out/sdust_full -f python scripts/prof_sdust.py 1200 1 > /dev/null 2>&1
ncu -i gpurun_out/sdust_full.ncu-rep --page raw --csv > gpurun_out/sdust_full_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_ing_copy -c 1 -o gpurun_out/ingest_full -f python scripts/prof_ingest.py 800 > /dev/null 2>&1
ncu -i gpurun_out/ingest_full.ncu-rep --page raw --csv > gpurun_out/ingest_full_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_ingest.csv python scripts/prof_ingest.py 800 > /dev/null 2>&1
python scripts/cli_bench.py 3000 > gpurun_out/${tag}_cli_bench_fasta_3000Mb.json 2>> gpurun_out/bench.err
python scripts/cli_bench_fastq.py 2000 > gpurun_out/${tag}_cli_bench_fastq_2000Mb.json 2>> gpurun_out/bench.err
tail -c 400 gpurun_out/${tag}_bench_c2.jsonl; cat gpurun_out/${tag}_gpu_tests.log gpurun_out/${tag}_cli_bench_fasta_3000Mb.json gpurun_out/${tag}_cli_bench_fastq_2000Mb.json
