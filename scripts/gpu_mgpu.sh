#!/bin/bash
# multi-GPU pass: parity tests for the world sizes this box has, then the bench at N = 2 (and 4, 8 if present)
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
t0=$SECONDS
timeout 1200 python -m pytest tests/test_multi_gpu.py tests/test_gpu_p2p.py -m gpu -x -q 2>&1 | tail -6
echo "tests: $((SECONDS-t0)) s"
for n in 2 4 8; do
  [ $n -le $NG ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  tail -1 gpurun_out/bench_n$n.json | cut -c1-900; tail -2 gpurun_out/bench_n$n.err | cut -c1-300
  UG4B200_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n${n}_nccl.json 2> gpurun_out/bench_n${n}_nccl.err
  tail -1 gpurun_out/bench_n${n}_nccl.json | cut -c1-300
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-300
echo "total: $((SECONDS-t0)) s"
