#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 300 python scripts/microbench_ln.py 2>&1 | tail -10
echo "=== bench C2 FLUXB200_LN_REREAD=1"
FLUXB200_LN_REREAD=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2l_bench_reread.json; python scripts/show_bench.py gpurun_out/r2l_bench_reread.json | grep -E "^value|ln_modulate|clocks|gemm_tc"
echo "=== bench C2 FLUXB200_LN_REREAD=0"
FLUXB200_LN_REREAD=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2l_bench.json; python scripts/show_bench.py gpurun_out/r2l_bench.json | grep -E "^value|ln_modulate|clocks|gemm_tc"
} 2>&1 | tee gpurun_out/r2l.log
