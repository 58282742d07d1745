#!/bin/bash
# round-2 GPU session H: branch-free payload update, idle-driven aggregation of host text (e2e), prefetch A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/h_*
timeout 900 python -m pytest tests/test_gpu_pass1.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py -q -p no:cacheprovider > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -5 gpurun_out/h_pytest.log
PGB200_SKM_STATS=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
PGB200_BUILD=nopf timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/h_bench_nopf.json 2> gpurun_out/h_bench_nopf.err
python - <<'PY'
import json
for f in ("h_bench", "h_bench_nopf"):
    try:
        txt = open(f"gpurun_out/{f}.json").read().strip().splitlines()
        d = json.loads([l for l in txt if l.startswith("{")][-1])
        print(f, "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2) if d.get("e2e") else None, "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), d["config"]["parity"][:30])
    except Exception as e:
        print(f, "failed", e)
PY
grep -c "aggregated epoch" gpurun_out/h_bench.err; tail -6 gpurun_out/h_bench.err
