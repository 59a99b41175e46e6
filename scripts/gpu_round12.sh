#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "=== C2 bf16 1024"
timeout 1500 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1e.json; python scripts/show_bench.py gpurun_out/bench_r1e.json
echo "=== C3 nf4 1024"
timeout 1500 python bench.py --steps 1 --warmup 3 --quant nf4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_r1_nf4.json; python scripts/show_bench.py gpurun_out/bench_r1_nf4.json
echo "=== C4 geometry 720x1280 bf16 (1 GPU)"
timeout 1500 python bench.py --steps 1 --warmup 3 --height 720 --width 1280 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_r1_720.json; python scripts/show_bench.py gpurun_out/bench_r1_720.json
} 2>&1 | tee gpurun_out/round12.log
