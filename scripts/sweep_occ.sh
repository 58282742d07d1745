#!/bin/bash
run() { python bench.py --genome 30000000 --steps 2 --warmup 1 --chunk-reads 2000000 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 step_ms', round(d['ms_per_step'],2), 'insert_ms', round(d['roofline']['insert_kernel_ms_per_step'],2), 'slots', d['config']['table_slots_per_gpu'])"; }
run base
PGB200_BENCH_SLOTS_MULT=2 run slots_x2
PGB200_BENCH_SLOTS_MULT=4 run slots_x4
for mb in 6 8; do
  mv soapdenovo2_b200/lib soapdenovo2_b200/lib_keep; cp -r soapdenovo2_b200/lib_mb$mb soapdenovo2_b200/lib
  run minblocks_$mb
  rm -rf soapdenovo2_b200/lib; mv soapdenovo2_b200/lib_keep soapdenovo2_b200/lib
done
