#!/bin/bash
for v in "" "-DPW_CTAS=3" "-DPW_CTAS=10"; do
  echo "variant: $v"; AVB_NVCC_EXTRA="$v" timeout 200 python tools/mstpp_bench.py 1 482 512 --check 2>&1 | grep -E "MST|gemm|parity"
done
