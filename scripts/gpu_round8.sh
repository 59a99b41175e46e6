#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_dit_gpu.py -m gpu -q -x 2>&1 | tail -6
timeout 1500 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r1c.json | cut -c1-600
} 2>&1 | tee gpurun_out/round8.log
