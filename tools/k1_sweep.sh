#!/bin/bash
# run on the GPU box: build K1 variants (only k1 object differs; other objects are rebuilt too but in parallel)
for v in "" "-DK1_THREADS_N=320" "-DK1_THREADS_N=768 -DK1_CTAS_PER_SM=1"; do
  AVB_NVCC_EXTRA="$v" python tools/kernel_times.py Rat 2>&1 | tail -1
done
