#!/bin/bash
# 8-GPU pass: one parity test at world 8, then the weak-scaling bench N = 1, 2, 4, 8
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
t0=$SECONDS
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "8-0-1--1 or 8-0-2-0" 2>&1 | tail -3
echo "tests: $((SECONDS-t0)) s"
run() { # n tag env...
  local n=$1 tag=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n${n}${tag}.json 2> gpurun_out/bench_n${n}${tag}.err
  tail -1 gpurun_out/bench_n${n}${tag}.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print('N',d['n_gpus'],'$tag','ms',round(d['ms_per_step'],3),'MDoF/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'its',d['config']['iterations'],'launches',d['gpu_launches_per_step'])
except Exception as e: print('bench failed', e)"
  grep -iE "error|Traceback" gpurun_out/bench_n${n}${tag}.err | head -3
}
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').readline()); print('N 1 ms',round(d['ms_per_step'],3),'MDoF/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1))"
for n in 2 4 8; do
  [ $n -le $NG ] || continue
  run $n "" A=1
done
echo "total: $((SECONDS-t0)) s"
