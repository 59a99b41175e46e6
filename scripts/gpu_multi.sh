#!/bin/bash
# N-GPU bench through torchrun, exactly as the driver launches it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 2 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_multi_${N}.log | cut -c1-900
