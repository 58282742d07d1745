#!/bin/bash
# round-2 GPU session P: the instance-packed aggregation kernel (k_skm_apply_q: CTA-wide quad packing, rolling k-mers) -- build trees
# vq (6 CTAs/SM), vq5 (5 CTAs/SM, 48 registers) against vb (per-warp batches, even split, slim claim)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/p_*
PGB200_BUILD=vq timeout 400 python -m pytest tests/test_gpu_pass1.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/p_pytest_vq.log
if ! grep -q passed gpurun_out/p_pytest_vq.log || grep -q failed gpurun_out/p_pytest_vq.log; then echo "vq pass-1 tests not green: no bench"; exit 0; fi
run() {
  env PGB200_BUILD=$2 $3 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/p_bench_$1.json 2> gpurun_out/p_bench_$1.err
}
run vb vb ""
run vq vq ""
run vq5 vq5 ""
run vq_b20 vq "PGB200_SKM_BUCKETS=1048576"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/p_bench_*.json")):
    try:
        txt = open(f).read().strip().splitlines()
        d = json.loads([l for l in txt if l.startswith("{")][-1])
        print(f, "ms", round(d["ms_per_step"], 2), "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), d["config"]["parity"][:30])
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".json", ".err")).read()[-600:])
PY
