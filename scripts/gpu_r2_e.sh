#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_r2_profiles.sh r2 2>&1 | tail -60
