#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 2 -c 1 -o gpurun_out/prof_attn_cur -f python scripts/prof_shapes.py attn 1 24 4608 2>&1 | tail -2
