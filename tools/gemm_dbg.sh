#!/bin/bash
# GEMM bottleneck experiments: time shape 0 (K=256 BRES) and shape 1 (K=2048) with parts of the kernel disabled
for d in ${DBG_MODES:-0 1 2 4 5 6}; do
  for s in ${DBG_SHAPES:-0 1}; do
    echo "debug=$d shape=$s: $(MQDET_GEMM_DEBUG=$d timeout 60 python tools/prof_gemm.py timeonly $s 2>&1 | tail -1)"
  done
done
