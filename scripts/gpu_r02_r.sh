#!/bin/bash
# round-2 GPU session R: even batches + slim claim (shipped), halved minimizer ring (k_skm_count at 8 CTAs/SM), 256-bit sweep, table
# clear on its own stream -- pass-1 / full-pipeline parity tests, then the N=1 bench line with a launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/r_*
timeout 900 python -m pytest tests/test_gpu_pass1.py tests/test_gpu_full.py tests/test_gpu_edge.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; tail -c 1500 gpurun_out/r_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 800 --csv --log-file gpurun_out/r_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r_ncu_bench.log 2>&1
python scripts/kern_times.py gpurun_out/r_launches.csv | sort | tail -25
