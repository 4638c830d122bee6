#!/bin/bash
# One gpurun call: GPU test-suite, smoke, bench, ncu launch list (+ optional full capture).
# usage: scripts/gpu_check.sh [tests|bench|ncu|all]
set -u
what=${1:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_info.csv 2>&1
nproc > gpurun_out/host_info.txt; free -g >> gpurun_out/host_info.txt; lscpu | head -20 >> gpurun_out/host_info.txt
if [[ $what == tests || $what == all ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
fi
if [[ $what == bench || $what == all ]]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
if [[ $what == ncu || $what == all ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  tail -2 gpurun_out/ncu_bench.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv1_kernel -s 40 -c 6 -f -o gpurun_out/prof_spmv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log
fi
