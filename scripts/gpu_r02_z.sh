#!/bin/bash
# round-2 GPU session Z (gpurun --gpus 8): final code -- strong-scaling point N = 8 on configs[1]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/z_*
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/z_c2_n8.json 2> gpurun_out/z_c2_n8.err; echo "n8 rc=$?"
python - <<'PY'
import json
txt = open("gpurun_out/z_c2_n8.json").read().strip().splitlines()
d = json.loads([l for l in txt if l.startswith("{")][-1])
print("N", d["n_gpus"], "ms", round(d["ms_per_step"], 2), "value %.4e" % d["value"], "e2e", round(d["e2e"]["ms_per_step"], 2), "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), d["config"]["parity"][:30])
PY
