#!/bin/bash
# one call: GPU tests, smoke, bench (with CPU baseline), reference arm, ncu launch list + full capture
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_info.csv 2>&1
nproc > gpurun_out/host_info.txt; free -g >> gpurun_out/host_info.txt; lscpu | head -20 >> gpurun_out/host_info.txt
t0=$SECONDS
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -45 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
echo "tests: $((SECONDS-t0)) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-3500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "bench: $((SECONDS-t0)) s"
UG4B200_NO_COMPRESS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err; cut -c1-300 gpurun_out/bench_plain.json; tail -3 gpurun_out/bench_plain.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-600 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
echo "ref: $((SECONDS-t0)) s"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv1 -s 60 -c 8 -f -o gpurun_out/prof_spmv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
echo "total: $((SECONDS-t0)) s"
