#!/bin/bash
# round-2 GPU session Y (gpurun --gpus 2): final code -- 2-GPU parity tests (two processes / one process / CLI), strong-scaling point N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/y_*
timeout 600 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider > gpurun_out/y_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/y_pytest_multi.log
tail -4 gpurun_out/y_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/y_c2_n2.json 2> gpurun_out/y_c2_n2.err; echo "n2 rc=$?"
python - <<'PY'
import json
txt = open("gpurun_out/y_c2_n2.json").read().strip().splitlines()
d = json.loads([l for l in txt if l.startswith("{")][-1])
print("N", d["n_gpus"], "ms", round(d["ms_per_step"], 2), "value %.4e" % d["value"], "e2e", round(d["e2e"]["ms_per_step"], 2), "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), d["config"]["parity"][:30])
PY
