#!/bin/bash
# batch kernel: GPU tests, level sweep with recording on/off, bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
sweep() {
for r in 4 5 6 7; do
  timeout 300 python bench.py --refs $r --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('refs',$r,'ms',round(d['ms_per_step'],4),'its',d['config']['iterations'],'launches',d['gpu_launches_per_step'],'us/it',round(1e3*d['ms_per_step']/(d['config']['iterations']+1),1))"
done
}
echo "== batch on (default max rows)"; sweep | tee gpurun_out/level_sweep_batch.txt
echo "== batch off"; UG4B200_BATCH=0 sweep | tee gpurun_out/level_sweep_nobatch.txt
echo "== batch on + PDL"; UG4B200_PDL=1 sweep
