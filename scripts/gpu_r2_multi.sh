#!/bin/bash
# round-2 multi-GPU run (gpurun --gpus 8): the BASELINE configs that are defined on 8 GPUs, measured as stated
#   C4  FLUX.1-dev 720x1280 50-step bf16, 8 prompts -> 1 per GPU
#   C5  FLUX.1-dev 1024x1024 50-step GGUF Q4_K, 32 prompts -> 4 per GPU
#   C2  at 8 GPUs (weak scaling of the headline config; the driver measures the full 1/2/4/8 curve itself)
# plus the two-devices-from-two-threads boundary test, which needs >= 2 GPUs in ONE process.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --no-cpu-baseline"
{
nvidia-smi --query-gpu=index,name,power.limit --format=csv,noheader | head -8
echo "=== two devices, two threads, one process"
timeout 600 python -m pytest tests/test_capi_and_host.py -m gpu -q -s 2>&1 | tail -3
echo "=== C4: 720x1280, 1 image per GPU, 8 GPUs"
timeout 900 $T --steps 2 --warmup 2 --height 720 --width 1280 2>&1 | tail -1 > gpurun_out/r2_bench_c4_8gpu.json; python scripts/show_bench.py gpurun_out/r2_bench_c4_8gpu.json | head -9
echo "=== C5: Q4_K, 4 images per GPU, 8 GPUs"
timeout 1200 $T --steps 1 --warmup 1 --quant q4k --batch 4 2>&1 | tail -1 > gpurun_out/r2_bench_c5_8gpu.json; python scripts/show_bench.py gpurun_out/r2_bench_c5_8gpu.json | head -9
echo "=== C2: 1024x1024 bf16, 1 image per GPU, 8 GPUs"
timeout 900 $T --steps 2 --warmup 2 2>&1 | tail -1 > gpurun_out/r2_bench_c2_8gpu.json; python scripts/show_bench.py gpurun_out/r2_bench_c2_8gpu.json | head -9
echo "=== C2 at 8 GPUs with FLUXB200_STEP_GRAPH=0 (what the step graph buys when 8 processes share the host)"
FLUXB200_STEP_GRAPH=0 timeout 900 $T --steps 2 --warmup 2 --no-kernel-timing 2>&1 | tail -1 > gpurun_out/r2_bench_c2_8gpu_nograph.json; python scripts/show_bench.py gpurun_out/r2_bench_c2_8gpu_nograph.json 2>/dev/null | head -3
} 2>&1 | tee gpurun_out/r2_multi.log
