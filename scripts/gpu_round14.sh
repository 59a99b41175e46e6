#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_vae_quant_gpu.py -m gpu -q -x -k "linear_quant" 2>&1 | tail -5
timeout 600 python scripts/microbench_quant.py 2>&1 | tail -14
} 2>&1 | tee gpurun_out/round14.log
