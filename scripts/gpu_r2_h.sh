#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "=== sdpa tests"
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k sdpa 2>&1 | tail -2
echo "=== attention trace (production variant)"
timeout 300 python scripts/attn_variants.py 0 2>&1 | tail -2 | cut -c1-520
echo "=== bench C2"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2h_bench.json; python scripts/show_bench.py gpurun_out/r2h_bench.json
} 2>&1 | tee gpurun_out/r2h.log
