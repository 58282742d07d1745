#!/bin/bash
# round-2 GPU session I: shared-memory table size / occupancy variants of k_skm_apply (build trees s10, s10b6, s9b8) vs the default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/i_*
for v in "" s10 s10b6 s9b8; do
  tag=${v:-base}
  PGB200_BUILD=$v timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/i_bench_$tag.json 2> gpurun_out/i_bench_$tag.err
done
python - <<'PY'
import json
for f in ("i_bench_base", "i_bench_s10", "i_bench_s10b6", "i_bench_s9b8"):
    try:
        txt = open(f"gpurun_out/{f}.json").read().strip().splitlines()
        d = json.loads([l for l in txt if l.startswith("{")][-1])
        print(f, "ms", round(d["ms_per_step"], 2), "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), d["config"]["parity"][:30])
    except Exception as e:
        print(f, "failed", e)
PY
tail -2 gpurun_out/i_bench_s9b8.err
