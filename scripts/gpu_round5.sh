#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 600 python scripts/microbench.py 2>&1 | tail -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/prof_gemm_qkv -f python scripts/prof_shapes.py gemm 4608 9216 3072 plain 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/prof_gemm_gelu -f python scripts/prof_shapes.py gemm 4608 12288 3072 gelu 2>&1 | tail -2
} 2>&1 | tee gpurun_out/round5.log
