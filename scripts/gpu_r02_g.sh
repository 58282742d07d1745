#!/bin/bash
# round-2 GPU session G (gpurun --gpus 2): weak-scaling code path at a small size before the 8-GPU run (configs[2] recipe, 200 Mbp)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/g_*
for k in 63 127; do
  extra=""; [ $k = 127 ] && export PGB200_BENCH_SLOTS_MULT=0.5
  PGB200_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --weak-genome 200000000 --K $k --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/g_weak_k$k.json 2> gpurun_out/g_weak_k$k.err; echo "weak K=$k rc=$?"
done
unset PGB200_BENCH_SLOTS_MULT
timeout 600 python bench.py --genome 200000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/g_n1_200m.json 2> gpurun_out/g_n1_200m.err
python - <<'PY'
import json
for f in ("g_weak_k63", "g_weak_k127", "g_n1_200m"):
    try:
        txt = open(f"gpurun_out/{f}.json").read().strip().splitlines()
        d = json.loads([l for l in txt if l.startswith("{")][-1])
        print(f, "N", d["n_gpus"], "ms", round(d["ms_per_step"], 2), "value %.3e" % d["value"], "distinct", d["config"]["distinct_kmers"], "instances", d["config"]["kmer_instances"], "chunks", d["config"]["chunks"], d["config"]["chunk_reads"], "slots", d["config"]["table_slots_per_gpu"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -4 gpurun_out/g_weak_k63.err; tail -4 gpurun_out/g_weak_k127.err
