#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "sdpa" 2>&1 | tail -4 | tee gpurun_out/r2q.log
