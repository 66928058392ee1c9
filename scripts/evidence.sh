set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r01_p_gpu_tests.log
python bench.py > gpurun_out/r01_p_bench_c2.jsonl 2> gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_p_bench_reference_arm.jsonl 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_p_launches_bench_c2.csv python bench.py --profile-only --steps 3 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_telofind_scan -s 3 -c 1 -o gpurun_out/scan_full -f python bench.py --profile-only --steps 2 --warmup 3 > /dev/null 2>&1
ncu -i gpurun_out/scan_full.ncu-rep --page raw --csv > gpurun_out/scan_full_raw.csv 2>/dev/null
tail -c 600 gpurun_out/r01_p_bench_c2.jsonl; cat gpurun_out/r01_p_gpu_tests.log
