#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv1_tma_kernel -s 5 -c 1 -f -o gpurun_out/prof_vi python scripts/kbench.py 7 > gpurun_out/ncu_vi.log 2>&1; tail -2 gpurun_out/ncu_vi.log
