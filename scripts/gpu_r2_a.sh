#!/bin/bash
# round-2 run A: validate the step graph / hoisted modulations / new parity tests, first bench, attention variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "=== pytest -m 'gpu and not slow'"
timeout 900 python -m pytest tests -m "gpu and not slow" -x -q -s 2>&1 | grep -v "^$" | tail -60
echo "=== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "=== bench (C2, step graph)"
timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/r2a_bench.json; python scripts/show_bench.py gpurun_out/r2a_bench.json
echo "=== bench (C2, FLUXB200_STEP_GRAPH=0)"
FLUXB200_STEP_GRAPH=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2a_bench_nograph.json; python scripts/show_bench.py gpurun_out/r2a_bench_nograph.json | head -5
echo "=== attention variants + trace"
timeout 400 python scripts/attn_variants.py 2>&1 | tail -12 | cut -c1-420
echo "=== gemm microbench"
timeout 300 python scripts/microbench.py 2>&1 | tail -14
echo "=== pytest -m 'gpu and slow'"
timeout 1200 python -m pytest tests -m "gpu and slow" -x -q -s 2>&1 | grep -v "^$" | tail -30
} 2>&1 | tee gpurun_out/r2a.log
