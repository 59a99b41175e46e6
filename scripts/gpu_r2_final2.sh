#!/bin/bash
# re-validation of the final tree (ln_reread on by default, extra attention variants): fast GPU suite, smoke, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m "gpu and not slow" -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_final2_bench.json; python scripts/show_bench.py gpurun_out/r2_final2_bench.json
} 2>&1 | tee gpurun_out/r2_final2.log
