#!/bin/bash
# round-2 run G: where does the 512x256 kernel lose?  clock64 trace of the MMA thread, small vs BIG tiles
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
FLUXB200_GEMM_TRACE=1 python -m diffusion_rs_b200.build --force > /dev/null 2>&1; echo build rc=$?
echo "=== 256x256 tiles"
FLUXB200_GEMM_TRACE=1 FLUXB200_GEMM_BIG=0 timeout 300 python scripts/gemm_trace.py 2>&1 | tail -6
echo "=== 512x256 tiles (long K)"
FLUXB200_GEMM_TRACE=1 FLUXB200_GEMM_BIG=1 timeout 300 python scripts/gemm_trace.py 2>&1 | tail -6
echo "=== 512x256 tiles (every GEMM)"
FLUXB200_GEMM_TRACE=1 FLUXB200_GEMM_BIG=2 timeout 300 python scripts/gemm_trace.py 2>&1 | tail -6
} 2>&1 | tee gpurun_out/r2g.log
