#!/bin/bash
# 1-GPU pass of round 2: complete GPU suite with the strict history errors recorded, the benches BASELINE.md names
# for one GPU (configs[1], configs[2] grid on one GPU, hierarchical order, CMK), ncu launch list + --set full capture.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_r02_single.sh'
set -u
mkdir -p gpurun_out
t0=$SECONDS
export UG4B200_RECORD_HIST_ERR=$PWD/gpurun_out/hist_err_1gpu.jsonl
rm -f $UG4B200_RECORD_HIST_ERR
timeout 600 python -m pytest tests -m gpu -q -rf --timeout 120 2>&1 | tail -40 | tee gpurun_out/tests_1gpu.log
unset UG4B200_RECORD_HIST_ERR
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; cut -c1-400 gpurun_out/bench_$name.json; tail -3 gpurun_out/bench_$name.err; }
run poisson --steps 10 --warmup 3
run poisson_hier --order hier --steps 5 --warmup 3 --no-cpu-baseline
run poisson_hier_cmk --order hier --reorder cmk --steps 5 --warmup 3 --no-cpu-baseline
run poisson_hier_rcmk --order hier --reorder rcmk --steps 5 --warmup 3 --no-cpu-baseline
run poisson_strong_n1 --scaling strong --steps 5 --warmup 3 ${STRONG_CPU:---no-cpu-baseline}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_poisson.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv1_vi_kernel -s 40 -c 3 -o gpurun_out/spmv_vi \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
echo "total: $((SECONDS-t0)) s"
