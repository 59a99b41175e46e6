#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/attn_variants.py 0 2 15 16 17 2>&1 | tail -7 | cut -c1-470 | tee gpurun_out/r2u.log
