#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests/test_vae_quant_gpu.py -m gpu -q -x -s -k "quantised" 2>&1 | grep -E "passed|failed|rror|DiT step|assert" | tail -12
echo "=== C3 nf4 1024"
timeout 1500 python bench.py --steps 1 --warmup 3 --quant nf4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_r1_nf4b.json; python scripts/show_bench.py gpurun_out/bench_r1_nf4b.json
} 2>&1 | tee gpurun_out/round13.log
