#!/bin/bash
# round-2 GPU session N (1 GPU): where the CLI's wall time goes at full configs[1] size (PGB200_VERBOSE timeline), with the .edge.gz
# and with the sidecar alone; then the drop-in tests (sidecar written by the stage, "only" mode)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/n_*
timeout 600 bash scripts/cli_full.sh 100000000 > gpurun_out/n_cli_gen.log 2>&1; echo "gen+run rc=$?"
D=/tmp/pgb200_cli
for mode in gz only; do
  rm -f $D/gpu.*
  if [ $mode = only ]; then export PGB200_EDGE_SIDECAR=only; fi
  s=$(date +%s.%N)
  PGB200_VERBOSE=2 soapdenovo2_b200/bin/pregraph-b200-63mer pregraph -s $D/c2.cfg -K 63 -p 8 -a 16 -R -o $D/gpu 2> gpurun_out/n_cli_${mode}_stderr.log
  e=$(date +%s.%N)
  python -c "print('CLI wall ($mode): %.2f s' % ($e - $s))" | tee -a gpurun_out/n_cli_walls.log
  grep -E "stage wall|reading the files|pass 1:" gpurun_out/n_cli_${mode}_stderr.log | tee -a gpurun_out/n_cli_walls.log
  ls -la $D/gpu.* >> gpurun_out/n_cli_walls.log
done
unset PGB200_EDGE_SIDECAR
grep "chunk" gpurun_out/n_cli_gz_stderr.log | head -30
timeout 900 python -m pytest tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/n_pytest_dropin.log
