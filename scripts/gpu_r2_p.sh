#!/bin/bash
# in-loop A/B of attention builds on one box: production (0), P in 4 instalments (2), no MUFU ping-pong (12)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for v in 0 2 12 0; do
  echo "=== --attn-variant $v"
  timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --attn-variant $v 2>&1 | tail -1 > gpurun_out/r2p_bench_attn$v.json
  python scripts/show_bench.py gpurun_out/r2p_bench_attn$v.json | grep -E "^value|attention|clocks"
done
} 2>&1 | tee gpurun_out/r2p.log
