#!/bin/bash
# round-2 GPU session D (gpurun --gpus 2): 2-GPU parity (two processes / one process / CLI), bench at N=2 (strong), N=1 beside it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/d_*
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/d_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/d_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider > gpurun_out/d_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest_multi.log
tail -30 gpurun_out/d_pytest_multi.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/d_bench_n2.json 2> gpurun_out/d_bench_n2.err; echo "bench n2 rc=$?"
tail -c 1800 gpurun_out/d_bench_n2.json; tail -5 gpurun_out/d_bench_n2.err
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/d_bench_n1.json 2> gpurun_out/d_bench_n1.err
python - <<'PY'
import json
for f in ("d_bench_n1", "d_bench_n2"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2) if d.get("e2e") else None, "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), "distinct", d["config"]["distinct_kmers"], d["config"]["parity"][:50])
    except Exception as e:
        print(f, "failed", e)
PY
ls -la gpurun_out | grep " d_"
