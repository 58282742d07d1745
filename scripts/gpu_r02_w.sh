#!/bin/bash
# round-2 GPU session W: fused sweeps with the spilled-key list -- pass-1 / pipeline parity tests, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/w_*
timeout 900 python -m pytest tests/test_gpu_pass1.py tests/test_gpu_full.py tests/test_gpu_edge.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/w_pytest.log
PGB200_BENCH_TIMELINE=1 PGB200_SKM_STATS=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err
grep "\[bench\]" gpurun_out/w_bench.err | tail -1; grep "aggregated epoch" gpurun_out/w_bench.err | tail -2
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/w_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
print("ms", round(d["ms_per_step"], 2), "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), d["config"]["parity"][:40])
PY
