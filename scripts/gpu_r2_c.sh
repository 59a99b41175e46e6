#!/bin/bash
# round-2 run C: cooperative-warp attention variants, persistent LN / dequant / GroupNorm, C2 + C3 + C5-slice benches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "=== pytest -m 'gpu and not slow'"
timeout 1200 python -m pytest tests -m "gpu and not slow" -q -s > gpurun_out/r2c_pytest_fast.log 2>&1; tail -6 gpurun_out/r2c_pytest_fast.log
echo "=== attention variants + trace"
timeout 600 python scripts/attn_variants.py 0 9 10 11 3 > gpurun_out/r2c_attn.log 2>&1; tail -7 gpurun_out/r2c_attn.log | cut -c1-470
echo "=== bench C2"
timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/r2c_bench.json; python scripts/show_bench.py gpurun_out/r2c_bench.json
for v in 9 10; do
echo "=== bench C2 --attn-variant $v"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --attn-variant $v 2>&1 | tail -1 > gpurun_out/r2c_bench_attn$v.json; python scripts/show_bench.py gpurun_out/r2c_bench_attn$v.json > gpurun_out/tmp.txt; head -3 gpurun_out/tmp.txt; grep -E "attention|clocks" gpurun_out/tmp.txt
done
echo "=== bench C3 nf4 (overlap on)"
timeout 900 python bench.py --steps 2 --warmup 2 --quant nf4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2c_bench_nf4.json; python scripts/show_bench.py gpurun_out/r2c_bench_nf4.json > gpurun_out/tmp.txt; head -3 gpurun_out/tmp.txt; grep -E "dequant|clocks|gemm_tc" gpurun_out/tmp.txt
echo "=== bench C3 nf4 FLUXB200_DEQUANT_OVERLAP=0"
FLUXB200_DEQUANT_OVERLAP=0 timeout 900 python bench.py --steps 2 --warmup 2 --quant nf4 --no-cpu-baseline --no-kernel-timing 2>&1 | tail -1 > gpurun_out/r2c_bench_nf4_noovl.json; python scripts/show_bench.py gpurun_out/r2c_bench_nf4_noovl.json 2>/dev/null | head -3
echo "=== bench C5 slice q4k batch 4"
timeout 900 python bench.py --steps 1 --warmup 2 --quant q4k --batch 4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2c_bench_q4k_b4.json; python scripts/show_bench.py gpurun_out/r2c_bench_q4k_b4.json > gpurun_out/tmp.txt; head -3 gpurun_out/tmp.txt; grep -E "dequant|clocks|gemm_tc" gpurun_out/tmp.txt
} 2>&1 | tee gpurun_out/r2c.log
