#!/bin/bash
# round-2 GPU session: K = 127 at N = 1 with the table sized from a realistic distinct estimate (2^29 slots, 2^19 buckets)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/k127_*
timeout 300 python bench.py --K 127 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/k127_bench.json 2> gpurun_out/k127_bench.err; tail -c 900 gpurun_out/k127_bench.json
