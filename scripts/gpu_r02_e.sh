#!/bin/bash
# round-2 GPU session E: the warp-synchronous one-warp-per-bucket kernel (build tree `next`) against the CTA kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/e_*
export PGB200_BUILD=next
timeout 900 python -m pytest tests/test_gpu_pass1.py tests/test_gpu_full.py tests/test_gpu_edge.py -q -p no:cacheprovider > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -8 gpurun_out/e_pytest.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/e_bench_cta.json 2> gpurun_out/e_bench_cta.err
PGB200_SKM_WARP=1 PGB200_SKM_STATS=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/e_bench_warp.json 2> gpurun_out/e_bench_warp.err
for ev in 3 5 7; do
  PGB200_SKM_WARP=1 PGB200_SKM_FLUSH_EVERY=$ev timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/e_bench_warp_ev$ev.json 2> gpurun_out/e_bench_warp_ev$ev.err
done
PGB200_SKM=0 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/e_bench_direct.json 2> gpurun_out/e_bench_direct.err
python - <<'PY'
import json
for f in ("e_bench_cta", "e_bench_warp", "e_bench_warp_ev3", "e_bench_warp_ev5", "e_bench_warp_ev7", "e_bench_direct"):
    try:
        txt = open(f"gpurun_out/{f}.json").read().strip().splitlines()
        d = json.loads([l for l in txt if l.startswith("{")][-1])
        print(f, "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2) if d.get("e2e") else None, "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), "slots", d["config"]["table_slots_per_gpu"], d["config"]["parity"][:30])
    except Exception as e:
        print(f, "failed", e)
PY
tail -4 gpurun_out/e_bench_warp.err
PGB200_SKM_WARP=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_skm_apply -s 1 -c 1 -o gpurun_out/e_applyw_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/e_ncu_applyw.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 800 --csv --log-file gpurun_out/e_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/e_ncu_bench.log 2>&1
python scripts/kern_times.py gpurun_out/e_launches.csv 2>/dev/null | head -30
ls -la gpurun_out | grep " e_"
