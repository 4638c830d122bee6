#!/bin/bash
# kbench for the default build and every library variant under gpurun_variants/; optional full test + bench
set -u
mkdir -p gpurun_out
echo "--- default"; python scripts/kbench.py 7 | cut -c1-330
for d in gpurun_variants/*/; do
  n=$(basename $d)
  [ -f $d/libug4b200.so ] || continue
  echo "--- $n"; UG4B200_LIBDIR=$PWD/gpurun_variants/$n timeout 120 python scripts/kbench.py 7 2>&1 | cut -c1-330
done | tee gpurun_out/variants.txt
if [[ "${1:-}" == full ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nocpu.json 2> gpurun_out/bench_nocpu.err; cut -c1-1200 gpurun_out/bench_nocpu.json; tail -3 gpurun_out/bench_nocpu.err
fi
