#!/bin/bash
# tests + bench A/B (PDL off/on)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; cut -c1-700 gpurun_out/bench_nopdl.json; tail -3 gpurun_out/bench_nopdl.err
UG4B200_PDL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; cut -c1-700 gpurun_out/bench_pdl.json; tail -3 gpurun_out/bench_pdl.err
UG4B200_PDL=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_pdl.log; tail -5 gpurun_out/pytest_gpu_pdl.log
