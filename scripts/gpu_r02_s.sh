#!/bin/bash
# round-2 GPU session S: where the step's wall time goes on the host side (bench timeline, engine verbose timers)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/s_*
PGB200_BENCH_TIMELINE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
grep "\[bench\]" gpurun_out/s_bench.err
PGB200_BENCH_TIMELINE=1 PGB200_VERBOSE=2 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/s_bench_v.json 2> gpurun_out/s_bench_v.err
grep -E "\[bench\]|reset_pass1|sweeps|finish|flush" gpurun_out/s_bench_v.err | tail -20
grep "chunk" gpurun_out/s_bench_v.err | tail -22
