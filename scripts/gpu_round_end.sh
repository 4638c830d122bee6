#!/bin/bash
# round-end style pass: GPU tests, smoke, bench (both arms)
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-3800 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
echo "total: $((SECONDS-t0)) s"
