#!/bin/bash
# level-cost sweep (ms per solve for hierarchies of 3..7 refinements) + full ncu capture of the top-level kernels
set -u
mkdir -p gpurun_out
for r in 3 4 5 6 7; do
  timeout 300 python bench.py --refs $r --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('refs',$r,'ms',round(d['ms_per_step'],4),'its',d['config']['iterations'],'launches',d['gpu_launches_per_step'],'us/it',round(1e3*d['ms_per_step']/(d['config']['iterations']+1),1))"
done | tee gpurun_out/level_sweep.txt
python scripts/kbench.py 7 | tee gpurun_out/kbench7.json | cut -c1-600
python scripts/kbench.py 6 | tee gpurun_out/kbench6.json | cut -c1-600
python scripts/kbench.py 5 | tee gpurun_out/kbench5.json | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv1_tma_kernel -s 5 -c 2 -f -o gpurun_out/prof_vi python scripts/kbench.py 7 > gpurun_out/ncu_vi.log 2>&1; tail -2 gpurun_out/ncu_vi.log
