#!/bin/bash
# Multi-GPU parity pass (2 / 4 / 8 GPUs — whatever the box has): tests/test_multi_gpu.py complete (-rf, no -x), the
# solver re-init tests, then one bench line per workload at the box's GPU count.
#   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/gpu_r02_mgpu.sh'
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
t0=$SECONDS
export UG4B200_RECORD_HIST_ERR=$PWD/gpurun_out/hist_err_mgpu${NG}.jsonl
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -rf --timeout 150 ${MGPU_K:+-k "$MGPU_K"} 2>&1 | tail -40 | tee gpurun_out/mgpu${NG}_tests.log
timeout 200 python -m pytest tests/test_gpu_solver.py -m gpu -q -rf --timeout 60 -k reinit 2>&1 | tail -15 | tee gpurun_out/reinit_tests.log
if [ "${1:-}" = "bench" ]; then
  for w in poisson convdiff elasticity; do
    extra=""; [ $w = elasticity ] && extra="--refs 6"
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $NG --workload $w $extra --steps 5 --warmup 3 > gpurun_out/bench_${w}_n${NG}.json 2> gpurun_out/bench_${w}_n${NG}.err
    cut -c1-700 gpurun_out/bench_${w}_n${NG}.json; tail -2 gpurun_out/bench_${w}_n${NG}.err
  done
fi
echo "total: $((SECONDS-t0)) s"
