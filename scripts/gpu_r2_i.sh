#!/bin/bash
# round-2 run I: the two-accumulator tile with the sub-tiles side by side along N (A shared), as cuBLAS's nvjet 192x256 does
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "=== big-tile tests"
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k big_tiles 2>&1 | tail -3
echo "=== microbench gemm_big=3 (long K)"
FLUXB200_GEMM_BIG=3 timeout 300 python scripts/microbench.py gemm 2>&1 | grep -E "12288|15360|8192" | tail -6
echo "=== microbench gemm_big=0"
FLUXB200_GEMM_BIG=0 timeout 300 python scripts/microbench.py gemm 2>&1 | grep -E "12288|15360|8192" | tail -6
FLUXB200_GEMM_TRACE=1 python -m diffusion_rs_b200.build --force > /dev/null 2>&1; echo trace build rc=$?
echo "=== trace gemm_big=3"
FLUXB200_GEMM_TRACE=1 FLUXB200_GEMM_BIG=3 timeout 300 python scripts/gemm_trace.py 2>&1 | tail -5 | cut -c1-700
} 2>&1 | tee gpurun_out/r2i.log
