#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "=== full gpu suite (except slow)"
timeout 1500 python -m pytest tests -m "gpu and not slow" -q -x -s 2>&1 | grep -E "passed|failed|error|DiT step|denoise|fused vs|720x1280|VAE decode|u8 image|nf4|q4k" | tail -30
echo "=== C1 full depth"
timeout 1200 python -m pytest tests/test_dit_gpu.py -m gpu -q -x -s -k c1_schnell 2>&1 | grep -E "passed|failed|error|C1" | tail -5
echo "=== microbench"
timeout 600 python scripts/microbench.py 2>&1 | tail -13
} 2>&1 | tee gpurun_out/round11.log
