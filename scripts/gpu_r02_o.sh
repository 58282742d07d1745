#!/bin/bash
# round-2 GPU session O: k_skm_apply variants -- even batches (records of a bucket split evenly over the warps), slim claim (empty
# shared-memory slots pre-initialised, claimers update with everybody else), no record prefetch -- build trees v0 va vb vc vd
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/o_*
run() {  # tag, build, extra env
  env PGB200_BUILD=$2 $3 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/o_bench_$1.json 2> gpurun_out/o_bench_$1.err
}
run v0 v0 ""
run va va ""
run vb vb ""
run vc vc ""
run vd vd ""
run vb_b20 vb "PGB200_SKM_BUCKETS=1048576"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/o_bench_*.json")):
    try:
        txt = open(f).read().strip().splitlines()
        d = json.loads([l for l in txt if l.startswith("{")][-1])
        print(f, "ms", round(d["ms_per_step"], 2), "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), d["config"]["parity"][:30])
    except Exception as e:
        print(f, "failed", e)
PY
PGB200_BUILD=vb timeout 900 python -m pytest tests/test_gpu_pass1.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/o_pytest_vb.log
