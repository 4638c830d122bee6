#!/bin/bash
# last GPU call of round 2: exactly what the driver runs at round end (GPU tests with -x, smoke, bench), plus the top-level
# ncu capture of gs_color_kernel that r02a missed (it caught coarse-level launches)
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/last_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/last_smoke.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; cut -c1-260 gpurun_out/bench_last.json; tail -2 gpurun_out/bench_last.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gs_color_kernel -s 24 -c 8 -o gpurun_out/gs_color_top \
  python scripts/kbench_ws.py convdiff 7 > /dev/null 2>&1
echo "total: $((SECONDS-t0)) s"
