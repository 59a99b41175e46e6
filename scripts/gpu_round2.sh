#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_dit_gpu.py -m gpu -q -x -s 2>&1 | tail -40 ) 2>&1 | tee gpurun_out/dit_tests.log
( timeout 600 python scripts/microbench.py 2>&1 | tail -40 ) 2>&1 | tee gpurun_out/microbench.log
