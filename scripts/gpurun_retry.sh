#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing is charged for those): gpurun_retry.sh <gpus> <timeout> <command>
G=$1; T=$2; shift 2
for i in $(seq 1 12); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_try.log 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > /tmp/gpurun_try.log 2>&1; fi
  if grep -q "status=transient\|status=busy" /tmp/gpurun_try.log; then echo "[retry $i] $(grep -o 'status=[a-z]*' /tmp/gpurun_try.log | head -1)"; sleep 150; continue; fi
  break
done
tail -120 /tmp/gpurun_try.log
