#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_p2p.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_mgpu.log; tail -5 gpurun_out/pytest_mgpu.log
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 "$@" 2>&1 | grep -E "P2PBENCH|^\{|Error|error" | cut -c1-900; }
echo "--- p2p single release"; run scripts/p2p_bench.py

echo "--- nccl"; UG4B200_P2P=0 run scripts/p2p_bench.py
echo "--- bench p2p"; run bench.py --gpus $N --steps 10 --warmup 3 | cut -c1-330
echo "--- bench p2p pdl"; UG4B200_PDL=1 run bench.py --gpus $N --steps 10 --warmup 3 | cut -c1-330
echo "--- bench p2p gather0"; UG4B200_GATHER_LEVEL=0 run bench.py --gpus $N --steps 10 --warmup 3 | cut -c1-330
echo "--- bench nccl"; UG4B200_P2P=0 run bench.py --gpus $N --steps 10 --warmup 3 | cut -c1-330
