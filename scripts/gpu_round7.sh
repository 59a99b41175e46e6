#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "=== pair-mode tests"
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_vae_quant_gpu.py -m gpu -q -x -k "linear or conv2d or vae_decode" 2>&1 | tail -8
echo "=== pair-mode microbench"
timeout 600 python scripts/microbench.py 2>&1 | grep -v attn | tail -12
echo "=== single-CTA microbench"
FLUXB200_GEMM_SINGLE_CTA=1 timeout 600 python scripts/microbench.py 2>&1 | grep -v attn | tail -12
} 2>&1 | tee gpurun_out/round7.log
