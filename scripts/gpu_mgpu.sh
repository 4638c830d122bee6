#!/bin/bash
# multi-GPU pass: parity tests for the world sizes this box has, then the bench at every N <= #GPUs
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
t0=$SECONDS
if [[ "${1:-tests}" == tests ]]; then
  timeout 1200 python -m pytest tests/test_multi_gpu.py tests/test_gpu_p2p.py -m gpu -x -q 2>&1 | tail -6
  echo "tests: $((SECONDS-t0)) s"
fi
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
for n in 2 4 8; do
  [ $n -le $NG ] || continue
  run $n "" A=1
  run $n _push UG4B200_FUSED_PUSH=1
  if [[ "${2:-}" == sweep ]]; then
    run $n _nopush_w128 UG4B200_NO_FUSED_PUSH=1 UG4B200_P2P_CTA_WORK=128
    run $n _nopush_w256 UG4B200_NO_FUSED_PUSH=1 UG4B200_P2P_CTA_WORK=256
    run $n _push_w128 UG4B200_P2P_CTA_WORK=128
    run $n _pdl UG4B200_PDL=1
    run $n _nopush_pdl UG4B200_NO_FUSED_PUSH=1 UG4B200_PDL=1
  fi
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('N 1 ms',round(d['ms_per_step'],3),'MDoF/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1))"
echo "total: $((SECONDS-t0)) s"
