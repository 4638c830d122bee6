#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tma.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -4
python scripts/kbench.py 7 | cut -c1-700
UG4B200_NO_COMPRESS=1 python scripts/kbench.py 7 | cut -c1-400
