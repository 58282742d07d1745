#!/bin/bash
# round-2 GPU session M (gpurun --gpus 2): FULL-SIZE configs[1] through the CLI, pass 1 sharded over 2 GPUs, next to the unmodified
# reference on the same box: all seven files must be identical; then the same on one GPU (timing of the round's final code)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out && rm -f gpurun_out/m_*
export PGB200_EDGE_SIDECAR=1
PGB200_GPUS=2 REF=1 timeout 1500 bash scripts/cli_full.sh 100000000 > gpurun_out/m_cli_full_2gpu.log 2>&1; echo "2-GPU CLI rc=$?"
tail -30 gpurun_out/m_cli_full_2gpu.log
cp /tmp/pgb200_cli/gpu.log gpurun_out/m_cli_full_2gpu_stderr.log 2>/dev/null
mkdir -p /tmp/pgb200_cli/keep && cp /tmp/pgb200_cli/ref.* /tmp/pgb200_cli/keep/ 2>/dev/null
s=$(date +%s.%N)
PGB200_VERBOSE=1 soapdenovo2_b200/bin/pregraph-b200-63mer pregraph -s /tmp/pgb200_cli/c2.cfg -K 63 -p 8 -a 16 -R -o /tmp/pgb200_cli/gpu1 2> gpurun_out/m_cli_full_1gpu_stderr.log
e=$(date +%s.%N)
python -c "print('1-GPU CLI wall: %.2f s' % ($e - $s))" | tee gpurun_out/m_cli_full_1gpu.log
for x in kmerFreq vertex preGraphBasic preArc edge.gz markOnEdge path; do cmp /tmp/pgb200_cli/gpu1.$x /tmp/pgb200_cli/ref.$x && echo "1-GPU $x identical" | tee -a gpurun_out/m_cli_full_1gpu.log; done
grep -E "pgb200|Time spent" gpurun_out/m_cli_full_1gpu_stderr.log | tee -a gpurun_out/m_cli_full_1gpu.log
ls -la /tmp/pgb200_cli/gpu1.edge.b200 /tmp/pgb200_cli/gpu1.edge.gz | tee -a gpurun_out/m_cli_full_1gpu.log
# contig through the sidecar (reference contig, linked with contig_sidecar.c) vs the reference's own contig on its own files
( time oracle/_ref/SOAPdenovo-63mer contig -g /tmp/pgb200_cli/ref -R ) > gpurun_out/m_contig_ref.log 2>&1
rm -f /tmp/pgb200_cli/gpu1.edge.gz
( time oracle/_ref/SOAPdenovo-63mer-b200 contig -g /tmp/pgb200_cli/gpu1 -R ) > gpurun_out/m_contig_sidecar.log 2>&1
for x in contig Arc updated.edge ContigIndex; do cmp /tmp/pgb200_cli/gpu1.$x /tmp/pgb200_cli/ref.$x && echo "contig $x identical (sidecar, no .edge.gz)" | tee -a gpurun_out/m_cli_full_1gpu.log; done
grep real gpurun_out/m_contig_ref.log gpurun_out/m_contig_sidecar.log | tee -a gpurun_out/m_cli_full_1gpu.log
