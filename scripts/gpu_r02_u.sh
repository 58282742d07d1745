#!/bin/bash
# round-2 GPU session U: end-of-pass sweeps fused into the aggregation flush -- parity tests, then the bench with and without it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/u_*
timeout 900 python -m pytest tests/test_gpu_pass1.py tests/test_gpu_full.py tests/test_gpu_edge.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/u_pytest.log
for mode in fused unfused; do
  if [ $mode = unfused ]; then export PGB200_NO_FUSED_SWEEP=1; fi
  PGB200_BENCH_TIMELINE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/u_bench_$mode.json 2> gpurun_out/u_bench_$mode.err
  grep "\[bench\]" gpurun_out/u_bench_$mode.err | tail -1
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/u_bench_$mode.json").read().strip().splitlines() if l.startswith("{")][-1])
print("$mode", "ms", round(d["ms_per_step"], 2), "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), d["config"]["parity"][:30])
PY
done
