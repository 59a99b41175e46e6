#!/bin/bash
# round-2 run D: epilogue residual prefetch, per-image weight cache for the quantised configs, ncu evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "=== pytest (ops, DiT, quantised step)"
timeout 1200 python -m pytest tests/test_ops_gpu.py tests/test_dit_gpu.py "tests/test_vae_quant_gpu.py::test_quantised_dit_step" -m "gpu and not slow" -q -s -x > gpurun_out/r2d_pytest.log 2>&1; tail -5 gpurun_out/r2d_pytest.log
echo "=== gemm microbench"
timeout 300 python scripts/microbench.py gemm 2>&1 | tail -13
echo "=== gemm microbench FLUXB200_GEMM_BIG=1 (long-K rows)"
FLUXB200_GEMM_BIG=1 timeout 300 python scripts/microbench.py gemm 2>&1 | grep -E "12288|15360|8192" | tail -6
echo "=== bench C2"
timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/r2d_bench.json; python scripts/show_bench.py gpurun_out/r2d_bench.json
echo "=== bench C3 nf4 (per-image weight cache, default)"
timeout 900 python bench.py --steps 2 --warmup 2 --quant nf4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2d_bench_nf4.json; python scripts/show_bench.py gpurun_out/r2d_bench_nf4.json > gpurun_out/tmp.txt; head -3 gpurun_out/tmp.txt; grep -E "dequant|clocks|gemm_tc|e2e" gpurun_out/tmp.txt
echo "=== bench C3 nf4 FLUXB200_DEQUANT_MODE=1 (staged per layer, pipelined)"
FLUXB200_DEQUANT_MODE=1 timeout 900 python bench.py --steps 2 --warmup 2 --quant nf4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2d_bench_nf4_staged.json; python scripts/show_bench.py gpurun_out/r2d_bench_nf4_staged.json > gpurun_out/tmp.txt; head -3 gpurun_out/tmp.txt; grep -E "dequant|clocks|gemm_tc" gpurun_out/tmp.txt
echo "=== bench C3 nf4 FLUXB200_DEQUANT_MODE=2 (fused producer)"
FLUXB200_DEQUANT_MODE=2 timeout 900 python bench.py --steps 1 --warmup 1 --quant nf4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2d_bench_nf4_fused.json; python scripts/show_bench.py gpurun_out/r2d_bench_nf4_fused.json > gpurun_out/tmp.txt; head -3 gpurun_out/tmp.txt; grep -E "dequant|clocks|gemm_tc" gpurun_out/tmp.txt
echo "=== bench C5 slice q4k batch 4"
timeout 900 python bench.py --steps 1 --warmup 2 --quant q4k --batch 4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2d_bench_q4k_b4.json; python scripts/show_bench.py gpurun_out/r2d_bench_q4k_b4.json > gpurun_out/tmp.txt; head -3 gpurun_out/tmp.txt; grep -E "dequant|clocks|gemm_tc|e2e" gpurun_out/tmp.txt
echo "=== bench C4 geometry 720x1280"
timeout 900 python bench.py --steps 2 --warmup 2 --height 720 --width 1280 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2d_bench_720.json; python scripts/show_bench.py gpurun_out/r2d_bench_720.json | head -4
echo "=== ncu profiles"
bash scripts/gpu_r2_profiles.sh r2 > /dev/null 2>&1; tail -25 gpurun_out/profiles_r2.log
} 2>&1 | tee gpurun_out/r2d.log
