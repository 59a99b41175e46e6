#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KREGEX='regex:tcgen05|ln_modulate|qknorm|gemv_jobs|silu_kernel|euler|pe_table|timestep_emb|vec_combine|gn_stats|gn_apply|upsample2x|softmax_rows|transpose_kernel|unpack_latents|postprocess_u8|dequant'
{
for t in test_groupnorm_nhwc test_vae_packed_u8 test_bnb_4bit_ffi_symbols test_bnb_int8_ffi_symbols; do
  echo "=== $t"
  timeout 300 python -m pytest tests/test_vae_quant_gpu.py -m gpu -q -s -k "$t" 2>&1 | tail -8
done
echo "=== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== bench"
timeout 1500 python bench.py --steps 2 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_r1_first.json
echo "=== ncu launch list (1 DiT step + VAE, our kernels only)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 1500 --csv --log-file gpurun_out/launches_r1.csv \
   python bench.py --steps 1 --warmup 0 --num-steps 1 --no-cpu-baseline --no-kernel-timing > gpurun_out/ncu_launch_run.log 2>&1
tail -2 gpurun_out/ncu_launch_run.log
wc -l gpurun_out/launches_r1.csv
echo "=== ncu full: gemm + attention"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 30 -c 2 -o gpurun_out/prof_gemm_r1 -f \
   python bench.py --steps 1 --warmup 0 --num-steps 1 --layers 2 --single-layers 2 --no-cpu-baseline --no-kernel-timing > gpurun_out/ncu_full_gemm.log 2>&1
tail -2 gpurun_out/ncu_full_gemm.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 2 -c 1 -o gpurun_out/prof_attn_r1 -f \
   python bench.py --steps 1 --warmup 0 --num-steps 1 --layers 2 --single-layers 2 --no-cpu-baseline --no-kernel-timing > gpurun_out/ncu_full_attn.log 2>&1
tail -2 gpurun_out/ncu_full_attn.log
ls -la gpurun_out/
} 2>&1 | tee gpurun_out/round4.log
