#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vi.json 2> gpurun_out/bench_vi.err; cut -c1-2500 gpurun_out/bench_vi.json; tail -3 gpurun_out/bench_vi.err
UG4B200_NO_COMPRESS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err; cut -c1-400 gpurun_out/bench_plain.json; tail -3 gpurun_out/bench_plain.err
