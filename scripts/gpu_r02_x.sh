#!/bin/bash
# round-2 GPU session X: verification of the round's FINAL code on one B200 -- whole -m gpu suite, smoke, default bench, ncu evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/x_*
timeout 2400 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x_pytest.log
tail -14 gpurun_out/x_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/x_smoke.log 2>&1; tail -1 gpurun_out/x_smoke.log
timeout 900 python bench.py > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
txt = open("gpurun_out/x_bench.json").read().strip().splitlines()
d = json.loads([l for l in txt if l.startswith("{")][-1])
print("value %.4e" % d["value"], "ms", round(d["ms_per_step"], 2), "e2e", d.get("e2e"), "cpu", d.get("cpu_baseline", {}).get("value"))
print("   roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "traffic", "issue_slots_pct", "insert_kernel_ms_per_step", "apply_kernel_ms_per_step")}, d["config"]["parity"][:40])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 800 --csv --log-file gpurun_out/x_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/x_ncu_bench.log 2>&1
python scripts/kern_times.py gpurun_out/x_launches.csv 2>/dev/null | sort | tail -22
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_skm_apply -s 1 -c 1 -o gpurun_out/x_apply_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/x_ncu_apply.log 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:k_skm_count|k_skm_scatter|k_line_index|k_decode_fast|k_nl_count" -s 25 -c 5 -o gpurun_out/x_front_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/x_ncu_front.log 2>&1
timeout 300 python bench.py --K 127 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/x_bench_k127.json 2> gpurun_out/x_bench_k127.err; tail -c 600 gpurun_out/x_bench_k127.json
ls -la gpurun_out | grep " x_"
