#!/bin/bash
# Link-level drop-in check: the reference's OWN objects (oracle/_ref/o{63,127}/*.o, compiled unmodified
# from /root/reference by oracle/Makefile) minus the six files the engine replaces, plus pregraph_shim.o and libpregraph_b200.so.
# The result, oracle/_ref/SOAPdenovo-{63,127}mer-b200, is the reference's main() / contig / map / scaff around the GPU pregraph:
#   SOAPdenovo-63mer-b200 pregraph ... | contig ... | all ...      (tests/test_gpu_dropin.py compares it with the unmodified binary)
set -e
cd "$(dirname "$0")/.."
REF=oracle/_ref
LIB=soapdenovo2_b200/lib
REFLIB=${REFLIBDIR:-/root/reference/sparsePregraph/inc}
[ -f $LIB/libpregraph_b200.so ] || make -s -j8 -C soapdenovo2_b200/csrc
for fl in 63 127; do
  [ -d $REF/o$fl ] || { echo "link_dropin: $REF/o$fl missing (run make -C oracle ref where /root/reference exists)"; exit 2; }
  objs=$(ls $REF/o$fl/*.o | grep -v -E '/(pregraph|prlHashReads|cutTipPreGraph|node2edge|prlRead2path|output_pregraph)\.o$')
  gcc -O2 -c -DPGB_FLAVOUR127=$([ $fl = 127 ] && echo 1 || echo 0) soapdenovo2_b200/csrc/pregraph_shim.c -o $REF/o$fl/pregraph_shim_b200.o.tmp
  # libbam.a (b= inputs of the other stages) ships with the reference; it travels to the GPU box inside the already linked binary
  g++ -no-pie $objs $REF/o$fl/pregraph_shim_b200.o.tmp -L$REFLIB -L$LIB -lpregraph_b200 -Wl,-rpath,'$ORIGIN/../../soapdenovo2_b200/lib' \
      -pthread -lz -lm -lbam -lrt -o $REF/SOAPdenovo-${fl}mer-b200
  rm -f $REF/o$fl/pregraph_shim_b200.o.tmp
  echo "linked $REF/SOAPdenovo-${fl}mer-b200"
done
