#!/bin/bash
# closing validation of round 2 after the last clean-ups (LU launch bounds, 32 run slots of the x-staged stream):
# the driver's GPU test command, smoke, configs[1], and the hierarchical order with Cuthill-McKee at upload (x-staged stream now)
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/last2_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/last2_smoke.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_last2.json 2> gpurun_out/bench_last2.err; cut -c1-200 gpurun_out/bench_last2.json; tail -2 gpurun_out/bench_last2.err
timeout 300 python bench.py --order hier --reorder cmk --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_hier_cmk2.json 2> gpurun_out/bench_hier_cmk2.err; cut -c1-200 gpurun_out/bench_hier_cmk2.json; tail -2 gpurun_out/bench_hier_cmk2.err
echo "total: $((SECONDS-t0)) s"
