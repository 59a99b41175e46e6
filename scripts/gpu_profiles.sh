#!/bin/bash
# ncu evidence for profiles/: launch list of one full-size DiT step + VAE decode, and --set full captures of the top kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1}
KREGEX='regex:tcgen05|ln_modulate|qknorm|gemv_jobs|silu_kernel|euler|pe_table|timestep_emb|vec_combine|gn_stats|gn_apply|upsample2x|softmax_rows|transpose_kernel|unpack_latents|postprocess_u8|dequant'
{
echo "=== launch list: bench.py --steps 1 --warmup 0 --num-steps 1 (1 DiT step + VAE, then the same through Pipeline.forward)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv \
   python bench.py --steps 1 --warmup 0 --num-steps 1 --no-cpu-baseline --no-kernel-timing > gpurun_out/ncu_launch_run_${TAG}.log 2>&1
wc -l gpurun_out/launches_${TAG}.csv
echo "=== ncu --set full: single-block fused GEMM 4608x21504x3072 (q|k|v -> QK-norm+RoPE, proj_mlp -> GELU)"
# in a --layers 1 --single-layers 1 run the gemm launches are: txt_in, img_in, dbl qkv, proj, mlp1, mlp2, sgl lin1, lin2, final
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 6 -c 2 -o gpurun_out/prof_gemm_single_${TAG} -f \
   python bench.py --steps 1 --warmup 0 --num-steps 1 --layers 1 --single-layers 1 --no-cpu-baseline --no-kernel-timing > gpurun_out/ncu_full_gemm_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_full_gemm_${TAG}.log | cut -c1-200
echo "=== ncu --set full: double-block GEMMs (qkv grouped, proj, mlp1 gelu, mlp2)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 4 -o gpurun_out/prof_gemm_double_${TAG} -f \
   python bench.py --steps 1 --warmup 0 --num-steps 1 --layers 1 --single-layers 1 --no-cpu-baseline --no-kernel-timing > gpurun_out/ncu_full_gemm2_${TAG}.log 2>&1
echo "=== ncu --set full: attention"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 1 -c 1 -o gpurun_out/prof_attn_${TAG} -f \
   python bench.py --steps 1 --warmup 0 --num-steps 1 --layers 1 --single-layers 1 --no-cpu-baseline --no-kernel-timing > gpurun_out/ncu_full_attn_${TAG}.log 2>&1
ls -la gpurun_out/*.ncu-rep
} 2>&1 | tee gpurun_out/profiles_${TAG}.log
