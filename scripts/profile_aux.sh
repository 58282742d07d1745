#!/bin/bash
# ncu --set full captures of the kernels besides k_chop_insert (one GPU, short workload)
cd "$(dirname "$0")/.."
D=/tmp/pgb200_prof; mkdir -p $D gpurun_out
PGB200_BUCKET=1 ncu --set full --clock-control none --import-source on -k regex:"k_decode_pack|k_bucket_count|k_bucket_scatter|k_apply_tuples|k_sweep" -c 6 -f -o gpurun_out/aux_pass1_r01 \
  python bench.py --genome 10000000 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python - <<PY
import sys, torch
sys.path.insert(0, '.')
import bench
G = 10000000; n = int(G * 30 / 300)
t1, t2 = bench.gen_pe_fastq_gpu(torch, 'cuda', G, n, 42)
t1.cpu().numpy().tofile('$D/p_1.fq'); t2.cpu().numpy().tofile('$D/p_2.fq')
open('$D/p.cfg','w').write("max_rd_len=150\n[LIB]\navg_ins=300\nreverse_seq=0\nasm_flags=3\nrank=1\nq1=$D/p_1.fq\nq2=$D/p_2.fq\n")
PY
ncu --set full --clock-control none --import-source on -k regex:"k_pass2|k_edge_walk|k_edge_emit|k_thin_walk|k_layout_place|k_layout_resolve|k_minor_decide" -c 8 -f -o gpurun_out/aux_graph_r01 \
  soapdenovo2_b200/bin/pregraph-b200-63mer pregraph -s $D/p.cfg -K 63 -p 8 -a 4 -R -o $D/out > /dev/null 2>&1
ls -la gpurun_out/aux_*.ncu-rep
