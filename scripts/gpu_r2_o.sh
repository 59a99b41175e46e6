#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dit_gpu.py -m "gpu and not slow" -q 2>&1 | tail -4 | tee gpurun_out/r2o.log
