#!/bin/bash
# sweep insert-kernel knobs on a 30 Mbp workload
for cfg in "0 0" "1 0" "2 0" "3 0" "4 0" "8 0" "0 1" "2 1" "4 1"; do
  set -- $cfg
  PGB200_PF=$1 PGB200_L2GRAN=$2 python bench.py --genome 30000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pf=$1 l2gran=$2 step_ms', round(d['ms_per_step'],2), 'insert_ms', round(d['roofline']['insert_kernel_ms_per_step'],2), 'frac', round(d['roofline']['frac'],3))"
done
