#!/bin/bash
set -u
NG=$(nvidia-smi -L | wc -l)
for w in 128 512 32; do
  UG4B200_P2P_CTA_WORK=$w timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) scripts/p2p_bench.py 2>/dev/null | grep P2PBENCH | sed "s/^/w=$w /"
done
UG4B200_P2P=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) scripts/p2p_bench.py 2>/dev/null | grep P2PBENCH | sed "s/^/nccl /"
