#!/bin/bash
# round-2 GPU session C: full -m gpu suite after the tips fix + dynamic batches + queued line index; e2e flush cadence; launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/c_*
timeout 2000 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -22 gpurun_out/c_pytest.log
PGB200_SKM_STATS=1 timeout 900 python bench.py --steps 3 --warmup 2 --write-digest > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
for ev in 0 7 10; do
  PGB200_SKM_FLUSH_EVERY=$ev timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c_bench_ev$ev.json 2> gpurun_out/c_bench_ev$ev.err
done
PGB200_SKM_WARP=1 PGB200_SKM_STATS=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/c_bench_warp.json 2> gpurun_out/c_bench_warp.err
PGB200_SKM_WARP=1 PGB200_SKM_FLUSH_EVERY=0 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c_bench_warp_ev0.json 2> gpurun_out/c_bench_warp_ev0.err
python - <<'PY'
import json
for f in ("c_bench", "c_bench_ev0", "c_bench_ev7", "c_bench_ev10", "c_bench_warp", "c_bench_warp_ev0"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2) if d.get("e2e") else None, "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), "parity", d["config"]["parity"][:40])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/c_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 800 --csv --log-file gpurun_out/c_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/c_ncu_bench.log 2>&1
python scripts/kern_times.py gpurun_out/c_launches.csv 2>/dev/null | head -30
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_skm_apply -s 1 -c 1 -o gpurun_out/c_apply_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/c_ncu_apply.log 2>&1
ls -la gpurun_out | grep " c_"
PGB200_SKM_WARP=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_skm_apply -s 1 -c 1 -o gpurun_out/c_applyw_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/c_ncu_applyw.log 2>&1
tail -3 gpurun_out/c_bench_warp.err
