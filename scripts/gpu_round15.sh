#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "=== C3 nf4 1024 (staged dequant, default)"
timeout 1500 python bench.py --steps 1 --warmup 3 --quant nf4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_r1_nf4c.json; python scripts/show_bench.py gpurun_out/bench_r1_nf4c.json
echo "=== C5 per-GPU slice: q4k 1024 batch 4"
timeout 1500 python bench.py --steps 1 --warmup 3 --quant q4k --batch 4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_r1_q4k_b4.json; python scripts/show_bench.py gpurun_out/bench_r1_q4k_b4.json
} 2>&1 | tee gpurun_out/round15.log
