#!/bin/bash
python -m pytest tests/test_gpu_mstpp.py -x -q 2>&1 | tail -2
for v in "" "-DDW_CPT=8"; do
  echo "variant: $v"; AVB_NVCC_EXTRA="$v" python tools/mstpp_bench.py 1 482 512 2>&1 | head -4
  AVB_NVCC_EXTRA="$v" python tools/mstpp_bench.py 8 482 512 2>&1 | head -1
done
