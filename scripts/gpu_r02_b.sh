#!/bin/bash
# round-2 GPU session B: the `all` hang, the whole -m gpu suite with the warp-autonomous apply kernel + 3-stream pipeline, timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/hang && rm -f gpurun_out/b_*
python - <<'PY' > gpurun_out/b_hang.log 2>&1
import sys; sys.path.insert(0, '.')
from soapdenovo2_b200 import synth
print(synth.scenario_pe_fastq('/tmp/hang', genome_len=30000, n_pairs=4000))
PY
for mode in "-a 1" ""; do
  echo "== pregraph K=31 -p 1 -R $mode" >> gpurun_out/b_hang.log
  PGB200_VERBOSE=2 timeout 90 soapdenovo2_b200/bin/pregraph-b200-63mer pregraph -s /tmp/hang/pe.cfg -K 31 -p 1 -R $mode -o /tmp/hang/x >> gpurun_out/b_hang.log 2>&1; echo "rc=$?" >> gpurun_out/b_hang.log
done
echo "== all through the drop-in binary" >> gpurun_out/b_hang.log
timeout 120 oracle/_ref/SOAPdenovo-63mer-b200 all -s /tmp/hang/pe.cfg -K 31 -p 1 -a 1 -R -o /tmp/hang/y > /dev/null 2>> gpurun_out/b_hang.log; echo "rc=$?" >> gpurun_out/b_hang.log
grep -E "^==|rc=|Time spent|error|Error" gpurun_out/b_hang.log | tail -30
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pass1.py -x -q -k "pe_fastq_k63 and aggregated or stress" > gpurun_out/b_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/b_memcheck.log
tail -4 gpurun_out/b_memcheck.log
timeout 2000 python -m pytest tests -m gpu -q --durations=20 -p no:cacheprovider > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -40 gpurun_out/b_pytest.log
PGB200_SKM_STATS=1 timeout 900 python bench.py --steps 3 --warmup 2 --write-digest > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"
tail -c 2600 gpurun_out/b_bench.json; tail -3 gpurun_out/b_bench.err
PGB200_BENCH_SLOTS_MULT=0.5 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/b_bench_half.json 2> gpurun_out/b_bench_half.err
python - <<'PY'
import json
for f in ("b_bench", "b_bench_half"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2) if d.get("e2e") else None, "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), "dec", round(d["roofline"]["decode_ms_per_step"], 2), "slots", d["config"]["table_slots_per_gpu"])
    except Exception as e:
        print(f, "failed", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 800 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu_bench.log 2>&1
python scripts/kern_times.py gpurun_out/b_launches.csv 2>/dev/null | head -30
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_skm_apply -s 1 -c 1 -o gpurun_out/b_apply_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu_apply.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:k_skm_count|k_line_index" -s 20 -c 2 -o gpurun_out/b_front_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu_front.log 2>&1
ls -la gpurun_out | grep " b_"
