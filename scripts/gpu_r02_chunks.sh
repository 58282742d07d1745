#!/bin/bash
# round-2 GPU session: chunk size of the device-resident feed (1 M / 2 M / 4 M reads per feed_text call)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/c_*
for c in 0 2000000 4000000; do
  PGB200_BENCH_TIMELINE=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --chunk-reads $c > gpurun_out/c_bench_$c.json 2> gpurun_out/c_bench_$c.err
  grep "\[bench\]" gpurun_out/c_bench_$c.err | tail -1
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/c_bench_$c.json").read().strip().splitlines() if l.startswith("{")][-1])
print("chunk $c", "ms", round(d["ms_per_step"], 2), "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), d["config"]["chunks"])
PY
done
