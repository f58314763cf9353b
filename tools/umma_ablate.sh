#!/bin/bash
# per-launch CUDA-event times of the tcgen05 training step under the knock-outs of NCF_UMMA_ABLATE
for ab in ${ABS:-0 1 2 4 8 16 3 7 12 31}; do
  echo "== ablate=$ab"
  NCF_UMMA_ABLATE=$ab NCF_UMMA_TIMING=1 STEPS=3 timeout 120 python tools/umma_step.py 2>&1 | grep "\[umma\]" | tail -8 | tr '\n' ' '
  echo
done
