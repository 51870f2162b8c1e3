#!/bin/bash
python -m pytest tests/test_gpu_honeybee.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
for v in "" "-DUV_SQRT_RN"; do AVB_NVCC_EXTRA="$v" python tools/kernel_times.py HoneyBee 2>&1 | tail -1; done
