#!/bin/bash
# the complete GPU suite on the final tree (what the driver runs at round end), smoke, and the two bench arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_final3_bench.json; python scripts/show_bench.py gpurun_out/r2_final3_bench.json | head -12
} 2>&1 | tee gpurun_out/r2_final3.log
