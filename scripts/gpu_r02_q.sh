#!/bin/bash
# round-2 GPU session Q: ncu capture (full set, source counters) of the instance-packed aggregation kernel, build tree vq
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/q_*
PGB200_BUILD=vq timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_skm_apply -s 1 -c 1 -o gpurun_out/q_applyq_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/q_ncu_apply.log 2>&1
tail -3 gpurun_out/q_ncu_apply.log
ls -la gpurun_out/q_*
