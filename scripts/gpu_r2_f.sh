#!/bin/bash
# round-2 run F: A/B of the stage-full barrier wait scope in the GEMM main loop (CTA vs acquire.cluster), GroupNorm apply
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "=== pytest ops + vae (current build: CTA-scope wait)"
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_vae_quant_gpu.py -m "gpu and not slow" -q -x -k "linear or conv or groupnorm or vae_decode or packed or sdpa" 2>&1 | tail -3
echo "=== gemm microbench: CTA-scope wait (default build)"
timeout 300 python scripts/microbench.py gemm 2>&1 | tail -13
echo "=== bench C2 (default build)"
timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/r2f_bench.json; python scripts/show_bench.py gpurun_out/r2f_bench.json
echo "=== rebuild with FLUXB200_FULL_WAIT_CLUSTER=1 (round-1 behaviour)"
FLUXB200_FULL_WAIT_CLUSTER=1 python -m diffusion_rs_b200.build --force > /dev/null 2>&1; echo rc=$?
echo "=== gemm microbench: acquire.cluster wait"
FLUXB200_FULL_WAIT_CLUSTER=1 timeout 300 python scripts/microbench.py gemm 2>&1 | tail -13
echo "=== bench C2 (acquire.cluster wait)"
FLUXB200_FULL_WAIT_CLUSTER=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2f_bench_clusterwait.json; python scripts/show_bench.py gpurun_out/r2f_bench_clusterwait.json | head -12
} 2>&1 | tee gpurun_out/r2f.log
