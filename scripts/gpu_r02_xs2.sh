#!/bin/bash
# x-staged stream as the fall-back for wide column windows: parity, 257^3 kernel timings, configs[1] and configs[2]-on-one-GPU benches
set -u
mkdir -p gpurun_out
t0=$SECONDS
export UG4B200_RECORD_HIST_ERR=$PWD/gpurun_out/hist_err_1gpu.jsonl; rm -f $UG4B200_RECORD_HIST_ERR
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 120 2>&1 | tail -30 | tee gpurun_out/tests_1gpu.log
unset UG4B200_RECORD_HIST_ERR
{
echo "--- refs 8 (257^3) default"; timeout 300 python scripts/kbench.py 8 2>&1 | tail -1 | cut -c1-900
echo "--- refs 8 (257^3) UG4B200_NO_XSTAGE=1 (plain stream)"; UG4B200_NO_XSTAGE=1 timeout 300 python scripts/kbench.py 8 2>&1 | tail -1 | cut -c1-900
echo "--- refs 7 UG4B200_XSTAGE=1"; UG4B200_XSTAGE=1 timeout 120 python scripts/kbench.py 7 2>&1 | tail -1 | cut -c1-900
echo "--- refs 7 default"; timeout 120 python scripts/kbench.py 7 2>&1 | tail -1 | cut -c1-900
echo "--- refs 7 default UG4B200_PDL=1"; UG4B200_PDL=1 timeout 120 python scripts/kbench.py 7 2>&1 | tail -1 | cut -c1-900
} | tee gpurun_out/kbench_257.txt
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; cut -c1-400 gpurun_out/bench_$name.json; tail -3 gpurun_out/bench_$name.err; }
run poisson --steps 10 --warmup 3 --no-cpu-baseline
UG4B200_PDL=1 run poisson_pdl --steps 10 --warmup 3 --no-cpu-baseline
run convdiff --workload convdiff --steps 5 --warmup 3 --no-cpu-baseline
run poisson_strong_n1 --scaling strong --steps 5 --warmup 3
echo "total: $((SECONDS-t0)) s"
