#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 600 python scripts/attn_variants.py 0 12 13 14 1 2>&1 | tail -7 | cut -c1-520
} 2>&1 | tee gpurun_out/r2k.log
