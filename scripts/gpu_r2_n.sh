#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vae_quant_gpu.py -m gpu -q -s -k "fully_quantised" 2>&1 | tail -6 | tee gpurun_out/r2n.log
