#!/bin/bash
# round-2 GPU session T: two-step k_skm_count (change list -> runs), build tree vm, against the main tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/t_*
PGB200_BUILD=vm timeout 600 python -m pytest tests/test_gpu_pass1.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/t_pytest_vm.log
for v in "" vm; do
  tag=${v:-main}
  PGB200_BUILD=$v PGB200_BENCH_TIMELINE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t_bench_$tag.json 2> gpurun_out/t_bench_$tag.err
  grep "\[bench\]" gpurun_out/t_bench_$tag.err | tail -1
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/t_bench_$tag.json").read().strip().splitlines() if l.startswith("{")][-1])
print("$tag", "ms", round(d["ms_per_step"], 2), "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), d["config"]["parity"][:30])
PY
done
PGB200_BUILD=vm timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/t_launches_vm.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/t_ncu.log 2>&1
python scripts/kern_times.py gpurun_out/t_launches_vm.csv | sort | grep -E "count|scatter|line_index|decode|nl_count|sweep|apply"
