#!/bin/bash
# round-2 run B: BIG-tile GEMM, expansion pipeline, full parity suite (logged), C2 / C3 benches with A/B flags
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "=== pytest -m 'gpu and not slow'"
timeout 1200 python -m pytest tests -m "gpu and not slow" -x -q -s > gpurun_out/r2b_pytest_fast.log 2>&1; tail -4 gpurun_out/r2b_pytest_fast.log
grep -E "block parity|rel err|DiT step|VAE decode|u8 image|720x1280|fused vs" gpurun_out/r2b_pytest_fast.log | cut -c1-220
echo "=== gemm microbench (big on / off)"
timeout 300 python scripts/microbench.py gemm 2>&1 | tail -10
FLUXB200_GEMM_BIG=0 timeout 300 python scripts/microbench.py gemm 2>&1 | grep -E "12288\)?,|15360|K': 12288|K': 15360|8192" | tail -4
FLUXB200_GEMM_BIG=2 timeout 300 python scripts/microbench.py gemm 2>&1 | tail -10
echo "=== bench C2 (gemm_big=1 default)"
timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/r2b_bench.json; python scripts/show_bench.py gpurun_out/r2b_bench.json
echo "=== bench C2 FLUXB200_GEMM_BIG=0"
FLUXB200_GEMM_BIG=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2b_bench_big0.json; python scripts/show_bench.py gpurun_out/r2b_bench_big0.json > gpurun_out/tmp.txt; head -8 gpurun_out/tmp.txt; grep gemm_tcgen05 gpurun_out/tmp.txt
echo "=== bench C2 FLUXB200_GEMM_BIG=2"
FLUXB200_GEMM_BIG=2 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2b_bench_big2.json; python scripts/show_bench.py gpurun_out/r2b_bench_big2.json > gpurun_out/tmp.txt; head -8 gpurun_out/tmp.txt; grep gemm_tcgen05 gpurun_out/tmp.txt
echo "=== bench C3 nf4 (overlap on)"
timeout 900 python bench.py --steps 2 --warmup 2 --quant nf4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2b_bench_nf4.json; python scripts/show_bench.py gpurun_out/r2b_bench_nf4.json > gpurun_out/tmp.txt; cat gpurun_out/tmp.txt
echo "=== bench C3 nf4 FLUXB200_DEQUANT_OVERLAP=0"
FLUXB200_DEQUANT_OVERLAP=0 timeout 900 python bench.py --steps 2 --warmup 2 --quant nf4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2b_bench_nf4_noovl.json; python scripts/show_bench.py gpurun_out/r2b_bench_nf4_noovl.json > gpurun_out/tmp.txt; head -3 gpurun_out/tmp.txt
echo "=== pytest -m 'gpu and slow'"
timeout 1200 python -m pytest tests -m "gpu and slow" -q -s > gpurun_out/r2b_pytest_slow.log 2>&1; tail -4 gpurun_out/r2b_pytest_slow.log
grep -E "full depth|VAE decode|u8 image" gpurun_out/r2b_pytest_slow.log | cut -c1-300
} 2>&1 | tee gpurun_out/r2b.log
