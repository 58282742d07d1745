#!/bin/bash
# round-2 (gpurun --gpus 2): N = 2 with 2.5 M-read chunks (2 per mate and rank) instead of 1 M (5 per mate and rank)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/n2c_*
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --chunk-reads 2500000 > gpurun_out/n2c.json 2> gpurun_out/n2c.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/n2c.json").read().strip().splitlines() if l.startswith("{")][-1])
print("N", d["n_gpus"], "ms", round(d["ms_per_step"], 2), "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), d["config"]["chunks"], d["config"]["parity"][:30])
PY
