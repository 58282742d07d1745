#!/bin/bash
# round-2 GPU session L (gpurun --gpus 8): final code -- 2-GPU parity tests, then the strong-scaling curve of configs[1] at N = 2, 4, 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/l_*
timeout 900 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider > gpurun_out/l_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest_multi.log
tail -4 gpurun_out/l_pytest_multi.log
run() {  # run <tag> <nproc> <args...>
  tag=$1; n=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $n "$@" > gpurun_out/l_$tag.json 2> gpurun_out/l_$tag.err
  echo "$tag rc=$?"
}
run c2_n8 8 --steps 5 --warmup 3 --no-cpu-baseline
run c2_n4 4 --steps 5 --warmup 3 --no-cpu-baseline
run c2_n2 2 --steps 5 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json
for f in ("l_c2_n2", "l_c2_n4", "l_c2_n8"):
    try:
        txt = open(f"gpurun_out/{f}.json").read().strip().splitlines()
        d = json.loads([l for l in txt if l.startswith("{")][-1])
        print(f, "N", d["n_gpus"], "ms", round(d["ms_per_step"], 2), "value %.4e" % d["value"], "e2e", round(d["e2e"]["ms_per_step"], 2), "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), d["config"]["parity"][:30])
    except Exception as e:
        print(f, "failed", e)
PY
