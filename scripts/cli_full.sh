#!/bin/bash
# Full-size configs[1] through the CLI (files on local scratch), timing every phase; optional reference run beside it.
set -e
cd "$(dirname "$0")/.."
G=${1:-100000000}
D=/tmp/pgb200_cli; mkdir -p $D
python - <<PY
import sys, torch, os
sys.path.insert(0, '.')
import bench
n = int($G * 30 / 300)
t1, t2 = bench.gen_pe_fastq_gpu(torch, 'cuda', $G, n, 42)
t1.cpu().numpy().tofile('$D/c2_1.fq'); t2.cpu().numpy().tofile('$D/c2_2.fq')
open('$D/c2.cfg','w').write("max_rd_len=150\n[LIB]\navg_ins=300\nreverse_seq=0\nasm_flags=3\nrank=1\nq1=$D/c2_1.fq\nq2=$D/c2_2.fq\n")
print('generated', n, 'pairs')
PY

s=$(date +%s.%N)
PGB200_VERBOSE=1 soapdenovo2_b200/bin/pregraph-b200-63mer pregraph -s $D/c2.cfg -K 63 -p 8 -a ${A:-16} -R -o $D/gpu 2> $D/gpu.log || { tail -20 $D/gpu.log; exit 1; }
e=$(date +%s.%N)
python -c "print(\"GPU CLI wall: %.2f s\" % ($e - $s))"
grep -E "pgb200|Time spent|node\(s\)|edge\(s\)|tip\(s\)|pre-arc|vertex" $D/gpu.log
ls -la $D/gpu.*
if [ -n "$REF" ]; then
  s=$(date +%s.%N)
  oracle/_ref/SOAPdenovo-63mer pregraph -s $D/c2.cfg -K 63 -p 8 -a ${A:-16} -R -o $D/ref 2> $D/ref.log
  e=$(date +%s.%N)
  python -c "print(\"REF wall: %.2f s\" % ($e - $s))"
  grep -E "Time spent" $D/ref.log
  for x in kmerFreq vertex preGraphBasic preArc edge.gz markOnEdge path; do cmp $D/gpu.$x $D/ref.$x && echo "$x identical"; done
fi
