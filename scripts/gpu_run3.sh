#!/bin/bash
# multi-GPU: parity tests + bench at N=$1 with P2P on/off
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_solver.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_mgpu.log; tail -8 gpurun_out/pytest_mgpu.log
for p2p in 1 0; do
  UG4B200_P2P=$p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_p2p${p2p}.json 2> gpurun_out/bench_n${N}_p2p${p2p}.err
  cut -c1-400 gpurun_out/bench_n${N}_p2p${p2p}.json; tail -3 gpurun_out/bench_n${N}_p2p${p2p}.err
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json
