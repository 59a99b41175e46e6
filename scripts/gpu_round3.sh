#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in test_conv2d_nhwc test_groupnorm_nhwc test_vae_decode_vs_oracle test_vae_packed_u8 test_bnb_4bit_ffi_symbols test_bnb_int8_ffi_symbols test_q4k_dequant test_quantised_dit_step; do
  echo "=== $t"
  timeout 600 python -m pytest tests/test_vae_quant_gpu.py -m gpu -q -s -k "$t" 2>&1 | tail -30
done 2>&1 | tee gpurun_out/vae_quant_tests.log
