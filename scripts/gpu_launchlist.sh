#!/bin/bash
# ncu launch list only (the cheap pass): one full-size DiT step + VAE decode, twice (value path and Pipeline.forward path)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1c}
KREGEX='regex:tcgen05|ln_modulate|qknorm|gemv_jobs|silu_kernel|euler|pe_table|timestep_emb|vec_combine|gn_stats|gn_apply|upsample2x|softmax_rows|transpose_kernel|unpack_latents|postprocess_u8|dequant'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv \
   python bench.py --steps 1 --warmup 0 --num-steps 1 --no-cpu-baseline --no-kernel-timing > gpurun_out/ncu_launch_run_${TAG}.log 2>&1
wc -l gpurun_out/launches_${TAG}.csv
