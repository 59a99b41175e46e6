#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "sdpa" 2>&1 | tail -4
timeout 600 python scripts/microbench.py 2>&1 | grep attn | tail -4
} 2>&1 | tee gpurun_out/round9.log
