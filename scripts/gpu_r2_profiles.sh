#!/bin/bash
# round-2 ncu evidence for profiles/: launch list of full-size DiT steps + VAE, and --set full captures of every kernel
# class north_star lists (GEMMs incl. the BIG-tile and conv-mode launches, attention, ln_modulate, GroupNorm,
# quantised-weight expansion, fused-dequant GEMM).  One GPU; numbers printed under ncu are never bench values.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2}
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-kernel-timing"
NCU="ncu --set full --clock-control none --import-source on -f"
{
echo "=== launch list: 2 full-depth DiT steps + VAE decode, x3 images (value path, e2e warm-up, e2e timed)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${TAG}.csv \
   $B --num-steps 2 > gpurun_out/ncu_launch_run_${TAG}.log 2>&1
wc -l gpurun_out/launches_${TAG}.csv
S="$B --num-steps 1 --layers 1 --single-layers 1"
echo "=== GEMMs of one double + one single block (1 DiT step: txt_in, 77 modulation groups..., then the step graph)"
timeout 900 $NCU -k regex:gemm_tcgen05 -s 8 -c 8 -o gpurun_out/prof_gemm_${TAG} $S > gpurun_out/ncu_gemm_${TAG}.log 2>&1
echo "=== attention"
timeout 600 $NCU -k regex:attention_tcgen05 -s 0 -c 2 -o gpurun_out/prof_attn_${TAG} $S > gpurun_out/ncu_attn_${TAG}.log 2>&1
echo "=== ln_modulate"
timeout 600 $NCU -k regex:ln_modulate -s 0 -c 6 -o gpurun_out/prof_ln_${TAG} $S > gpurun_out/ncu_ln_${TAG}.log 2>&1
echo "=== GroupNorm + the other VAE passes"
timeout 900 $NCU -k "regex:gn_stats|gn_apply|upsample2x|softmax_rows" -s 40 -c 16 -o gpurun_out/prof_vae_hbm_${TAG} $S > gpurun_out/ncu_vae_${TAG}.log 2>&1
echo "=== conv-mode GEMM at 1024x1024 (the last VAE convolutions: 128 -> 128 channels, conv_out 128 -> 3)"
timeout 900 $NCU -k regex:gemm_tcgen05 -s 50 -c 7 -o gpurun_out/prof_conv_${TAG} $S > gpurun_out/ncu_conv_${TAG}.log 2>&1
echo "=== NF4: expansion kernel (staged path)"
timeout 900 $NCU -k regex:dequant_batch -s 4 -c 6 -o gpurun_out/prof_dequant_${TAG} $S --quant nf4 > gpurun_out/ncu_dequant_${TAG}.log 2>&1
echo "=== NF4: fused-dequant GEMM (operand producer)"
timeout 900 python scripts/prof_quant_gemm.py > /dev/null 2>&1
timeout 900 $NCU -k regex:gemm_tcgen05 -s 2 -c 2 -o gpurun_out/prof_gemm_fusedq_${TAG} python scripts/prof_quant_gemm.py > gpurun_out/ncu_fusedq_${TAG}.log 2>&1
echo "=== yardsticks (vendor kernels, for comparison only): cuBLAS GEMM, torch SDPA"
timeout 600 $NCU -k "regex:nvjet|cutlass|xmma|gemm|sm100|sm90" -s 1 -c 1 -o gpurun_out/prof_yard_gemm_${TAG} python scripts/prof_yardsticks.py gemm 4608 21504 3072 > gpurun_out/ncu_yard_gemm_${TAG}.log 2>&1
timeout 600 $NCU -k "regex:nvjet|cutlass|xmma|gemm|sm100|sm90" -s 1 -c 1 -o gpurun_out/prof_yard_gemm2_${TAG} python scripts/prof_yardsticks.py gemm 4608 3072 15360 > gpurun_out/ncu_yard_gemm2_${TAG}.log 2>&1
timeout 600 $NCU -k "regex:fmha|flash|cudnn|attention|sdpa|sm100|sm90" -s 1 -c 1 -o gpurun_out/prof_yard_attn_${TAG} python scripts/prof_yardsticks.py attn > gpurun_out/ncu_yard_attn_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_yard_gemm_${TAG}.log gpurun_out/ncu_yard_attn_${TAG}.log
ls -la gpurun_out/*_${TAG}.ncu-rep
echo "=== summaries (text) -> gpurun_out/profiles_${TAG}/ ; the .ncu-rep files stay on the box (gpurun_out is capped at 64 MiB)"
FLUXB200_PROFILE_OUT=gpurun_out/profiles_${TAG} python scripts/summarize_profiles.py ${TAG}
rm -f gpurun_out/*_${TAG}.ncu-rep
du -sh gpurun_out
} 2>&1 | tee gpurun_out/profiles_${TAG}.log
