#!/bin/bash
# round-2 GPU session V: fused vs separate sweeps again, host timeline of the TIMED steps only
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/v_*
for mode in fused unfused fused2; do
  unset PGB200_NO_FUSED_SWEEP
  if [ $mode = unfused ]; then export PGB200_NO_FUSED_SWEEP=1; fi
  PGB200_BENCH_TIMELINE=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/v_bench_$mode.json 2> gpurun_out/v_bench_$mode.err
  grep "\[bench\]" gpurun_out/v_bench_$mode.err | tail -1
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/v_bench_$mode.json").read().strip().splitlines() if l.startswith("{")][-1])
print("$mode", "ms", round(d["ms_per_step"], 2), "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), "dec", round(d["roofline"]["decode_stream_ms_per_step"], 2))
PY
done
