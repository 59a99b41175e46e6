#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_dit_gpu.py -m gpu -q -x -s 2>&1 | grep -E "passed|failed|DiT step|denoise|fused vs"
timeout 900 python -m pytest tests/test_vae_quant_gpu.py -m gpu -q -x -k "quantised" 2>&1 | tail -4
timeout 1500 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1d.json
python scripts/show_bench.py gpurun_out/bench_r1d.json
} 2>&1 | tee gpurun_out/round10.log
