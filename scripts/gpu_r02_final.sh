#!/bin/bash
# Final 1-GPU pass of round 2: complete GPU suite (strict history errors recorded), the bench lines of BASELINE.md §3 that run
# on one GPU, launch list + ncu --set full captures of the two dominant kernels.
set -u
mkdir -p gpurun_out
t0=$SECONDS
export UG4B200_RECORD_HIST_ERR=$PWD/gpurun_out/hist_err_1gpu.jsonl; rm -f $UG4B200_RECORD_HIST_ERR
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 180 2>&1 | tail -25 | tee gpurun_out/tests_1gpu.log
unset UG4B200_RECORD_HIST_ERR
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; cut -c1-300 gpurun_out/bench_$name.json; tail -3 gpurun_out/bench_$name.err; }
run poisson --steps 10 --warmup 3
run elasticity_97 --workload elasticity --base-mult 3 --refs 5 --steps 5 --warmup 3
run convdiff_257 --workload convdiff --scaling strong --steps 3 --warmup 3 --no-cpu-baseline
timeout 300 python scripts/time_s1.py 2>&1 | tail -1 | tee gpurun_out/s1_2d.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference_arm.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_poisson.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv1_vi_kernel -s 12 -c 2 -o gpurun_out/spmv_vi_top \
  python scripts/kbench.py 7 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmv1_xs_kernel -s 8 -c 1 -o gpurun_out/spmv_xs_257 \
  python scripts/kbench.py 8 > /dev/null 2>&1
echo "total: $((SECONDS-t0)) s"
