#!/bin/bash
# Run every GPU op test in its own process (a trapped kernel kills the CUDA context) with a hard timeout.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader
for t in test_linear_plain test_linear_bias_after_round_and_nobias test_linear_gelu test_linear_gate_residual test_sdpa test_layernorm_modulate test_qknorm_rope; do
  echo "=== $t"
  timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "$t" -x 2>&1 | tail -25
done 2>&1 | tee gpurun_out/ops_groups.log
