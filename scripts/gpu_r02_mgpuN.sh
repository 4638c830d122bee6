#!/bin/bash
# Multi-GPU measurement pass on whatever GPU count the box has (run with gpurun --gpus N):
#   strong scaling (257^3 on N GPUs), configs[3] / configs[4] at N GPUs, a few N-GPU parity tests.
#   MGPU_TESTS: pytest -k expression (default: the random-rhs cases of this GPU count); MGPU_EXTRA=1: convdiff + elasticity benches
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
t0=$SECONDS
tr() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $NG "$@" > gpurun_out/bench_${name}_n${NG}.json 2> gpurun_out/bench_${name}_n${NG}.err
      cut -c1-500 gpurun_out/bench_${name}_n${NG}.json; grep -v "destroy_process_group\|^\*\*\*\|^$\|OMP_NUM" gpurun_out/bench_${name}_n${NG}.err | tail -3; }
tr poisson_strong --scaling strong --steps 10 --warmup 3
if [ "${MGPU_EXTRA:-0}" = "1" ]; then
  tr poisson_weak --steps 10 --warmup 3
  tr convdiff --workload convdiff --steps 5 --warmup 3
  tr elasticity --workload elasticity --base-mult 3 --refs 5 --steps 5 --warmup 3
fi
export UG4B200_RECORD_HIST_ERR=$PWD/gpurun_out/hist_err_mgpu${NG}.jsonl; rm -f $UG4B200_RECORD_HIST_ERR
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -rf --timeout 200 -k "${MGPU_TESTS:-random_rhs and ${NG}-}" 2>&1 | tail -25 | tee gpurun_out/mgpu${NG}_tests.log
echo "total: $((SECONDS-t0)) s"
