#!/bin/bash
# round-2: the default bench line of the final code (device-resident pass fed in 4 M-read chunks)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/fb_*
timeout 300 python bench.py > gpurun_out/fb_bench.json 2> gpurun_out/fb_bench.err; echo "rc=$?"; tail -c 300 gpurun_out/fb_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/fb_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value %.4e" % d["value"], "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "frac", round(d["roofline"]["frac"], 4), "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "launches", d["gpu_launches"], d["config"]["chunks"], d["config"]["e2e_chunk_reads"], d["clocks"])
PY
