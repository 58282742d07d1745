#!/bin/bash
# round-2 GPU session F (gpurun --gpus 8): the scaling curve on configs[1] (strong) and the human-scale targets configs[2] / configs[3]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/f_*
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/f_gpus.txt 2>&1
run() {  # run <tag> <nproc> <args...>
  tag=$1; n=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $n "$@" > gpurun_out/f_$tag.json 2> gpurun_out/f_$tag.err
  echo "$tag rc=$?"
}
run c2_n8 8 --steps 3 --warmup 2 --no-cpu-baseline
run c2_n4 4 --steps 3 --warmup 2 --no-cpu-baseline
PGB200_VERBOSE=1 run c3_n8 8 --weak-genome 3000000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
PGB200_BENCH_SLOTS_MULT=0.25 PGB200_VERBOSE=1 run c4_n8 8 --weak-genome 3000000000 --K 127 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
python - <<'PY'
import json
for f in ("f_c2_n8", "f_c2_n4", "f_c3_n8", "f_c4_n8"):
    try:
        txt = open(f"gpurun_out/{f}.json").read().strip().splitlines()
        d = json.loads([l for l in txt if l.startswith("{")][-1])
        print(f, "N", d["n_gpus"], "ms", round(d["ms_per_step"], 2), "value %.3e" % d["value"], "inst/s %.3e" % d["config"]["instances_per_s"], "e2e", round(d["e2e"]["ms_per_step"], 2) if d.get("e2e") else None,
              "ins", round(d["roofline"]["insert_kernel_ms_per_step"], 2), "apply", round(d["roofline"]["apply_kernel_ms_per_step"], 2), "distinct", d["config"]["distinct_kmers"], "frac", round(d["roofline"]["frac"], 3), d["config"]["parity"][:30])
    except Exception as e:
        print(f, "failed", e)
PY
tail -5 gpurun_out/f_c3_n8.err; tail -5 gpurun_out/f_c4_n8.err
nvidia-smi --query-gpu=index,memory.used --format=csv >> gpurun_out/f_gpus.txt 2>&1
ls -la gpurun_out | grep " f_"
