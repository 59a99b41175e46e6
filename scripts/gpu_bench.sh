#!/bin/bash
# full-size bench (our arm) + reference arm; logs under gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1}
{
timeout 1500 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}.json
} 2>&1 | tee gpurun_out/bench_${TAG}.log
