#!/bin/bash
# round-2 GPU session J: verification of the round's final code on one B200 -- whole -m gpu suite, bench (both arms), ncu evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/j_*
timeout 2400 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.log
tail -16 gpurun_out/j_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_smoke.log 2>&1; tail -1 gpurun_out/j_smoke.log
timeout 900 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "bench rc=$?"
( time timeout 1500 python bench.py --impl reference ) > gpurun_out/j_bench_ref.json 2> gpurun_out/j_bench_ref.err; echo "reference arm rc=$?"
python - <<'PY'
import json
for f in ("j_bench", "j_bench_ref"):
    try:
        txt = open(f"gpurun_out/{f}.json").read().strip().splitlines()
        d = json.loads([l for l in txt if l.startswith("{")][-1])
        print(f, "value %.4e" % d["value"], "ms", round(d["ms_per_step"], 2), "e2e", d.get("e2e"), "cpu", d.get("cpu_baseline"))
        if "roofline" in d: print("   roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "traffic", "issue_slots_pct", "insert_kernel_ms_per_step", "apply_kernel_ms_per_step")}, d["config"]["parity"][:40])
    except Exception as e:
        print(f, "failed", e)
PY
tail -4 gpurun_out/j_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 800 --csv --log-file gpurun_out/j_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/j_ncu_bench.log 2>&1
python scripts/kern_times.py gpurun_out/j_launches.csv 2>/dev/null | head -30
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_skm_apply -s 1 -c 1 -o gpurun_out/j_apply_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/j_ncu_apply.log 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:k_skm_count|k_skm_scatter|k_line_index|k_decode_fast|k_nl_count" -s 25 -c 5 -o gpurun_out/j_front_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/j_ncu_front.log 2>&1
ls -la gpurun_out | grep " j_"
