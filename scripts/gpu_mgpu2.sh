#!/bin/bash
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_p2p.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) scripts/p2p_bench.py 2>/dev/null | grep P2PBENCH
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
run $NG "" A=1
run $NG _push UG4B200_FUSED_PUSH=1
run $NG _gather5 UG4B200_GATHER_LEVEL=5
run $NG _push_gather5 UG4B200_FUSED_PUSH=1 UG4B200_GATHER_LEVEL=5
