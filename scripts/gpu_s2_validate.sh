#!/bin/bash
# session-2 re-validation: GPU parity tests, smoke, full-size bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "=== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== bench"
timeout 900 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_s2.json; python scripts/show_bench.py gpurun_out/bench_s2.json
} 2>&1 | tee gpurun_out/s2_validate.log
