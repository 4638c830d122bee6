#!/bin/bash
# profiles: ncu launch list of one bench run + full capture of the dominant kernel (and the batch kernel)
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv1_vi_kernel -s 5 -c 2 -f -o gpurun_out/prof_spmv python scripts/kbench.py 7 > gpurun_out/ncu_vi.log 2>&1; tail -1 gpurun_out/ncu_vi.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 20 -c 1 -f -o gpurun_out/prof_batch python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_batch.log 2>&1; tail -1 gpurun_out/ncu_batch.log | cut -c1-200
python scripts/kbench.py 7 | tee gpurun_out/kbench7.json | cut -c1-700
python scripts/kbench.py 6 | tee gpurun_out/kbench6.json | cut -c1-700
