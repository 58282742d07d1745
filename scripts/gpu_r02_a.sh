#!/bin/bash
# round-2 GPU session A: correctness of the rewritten pass 1 (sanitizer on a small case, the whole -m gpu suite), then timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
tail -3 gpurun_out/a_smoke.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pass1.py -x -q -k "pe_fastq_k63 and aggregated" > gpurun_out/a_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/a_memcheck.log
tail -5 gpurun_out/a_memcheck.log
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -25 gpurun_out/a_pytest.log
PGB200_SKM_STATS=1 PGB200_VERBOSE=1 timeout 900 python bench.py --steps 3 --warmup 2 --write-digest > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/a_bench.json
tail -5 gpurun_out/a_bench.err
PGB200_SKM=0 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/a_bench_direct.json 2> gpurun_out/a_bench_direct.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 800 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/a_ncu_bench.log 2>&1
python scripts/kern_times.py gpurun_out/a_launches.csv 2>/dev/null | head -30
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_skm_apply -s 1 -c 1 -o gpurun_out/a_apply_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/a_ncu_apply.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:k_skm_count|k_skm_scatter|k_decode_fast|k_line_index|k_nl_count|k_decode_fix|k_skm_publish" -s 14 -c 7 -o gpurun_out/a_front_full python bench.py --genome 20000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/a_ncu_front.log 2>&1
ls -la gpurun_out | tail -20
