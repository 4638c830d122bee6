#!/bin/bash
# x-staged SpMV kernel: parity first, then kernel timings (default build, UG4B200_NO_XSTAGE=1, library variants) at
# 129^3 and 65^3, then the bench.
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 600 python -m pytest tests/test_gpu_tma.py tests/test_gpu_kernels.py tests/test_gpu_solver.py tests/test_gpu_p2p.py -m gpu -q -x --timeout 120 2>&1 | tail -15 | tee gpurun_out/tests_xs.log
{
for r in 7 6; do
  echo "--- default refs $r"; timeout 120 python scripts/kbench.py $r 2>&1 | tail -1 | cut -c1-700
  echo "--- no_xstage refs $r"; UG4B200_NO_XSTAGE=1 timeout 120 python scripts/kbench.py $r 2>&1 | tail -1 | cut -c1-700
  for d in gpurun_variants/*/; do
    n=$(basename $d); [ -f $d/libug4b200.so ] || continue
    echo "--- $n refs $r"; UG4B200_LIBDIR=$PWD/gpurun_variants/$n timeout 120 python scripts/kbench.py $r 2>&1 | tail -1 | cut -c1-700
  done
done
} | tee gpurun_out/variants_xs.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_xs.json 2> gpurun_out/bench_xs.err; cut -c1-600 gpurun_out/bench_xs.json; tail -3 gpurun_out/bench_xs.err
timeout 800 python bench.py --scaling strong --steps 5 --warmup 3 > gpurun_out/bench_strong_n1.json 2> gpurun_out/bench_strong_n1.err; cut -c1-300 gpurun_out/bench_strong_n1.json; tail -3 gpurun_out/bench_strong_n1.err
echo "total: $((SECONDS-t0)) s"
