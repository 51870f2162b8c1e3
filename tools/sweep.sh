#!/bin/bash
for v in "-DUV_MINB=4 -DUV_RH_ROWS=32" "-DUV_MINB=4 -DUV_RH_ROWS=48" "-DUV_MINB=5 -DUV_RH_ROWS=32"; do
  AVB_NVCC_EXTRA="$v" python tools/bee_kernels.py HoneyBee 2>&1 | tail -1
done
for v in "" "-DG_MINB3_R=8"; do
  for sp in Squirrel Bear Raccoon Lion; do AVB_NVCC_EXTRA="$v" python tools/bee_kernels.py $sp 2>&1 | tail -1; done
done
