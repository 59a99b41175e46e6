#!/bin/bash
# session-2 final evidence: GPU parity tests, smoke, both bench arms, the other BASELINE configs, microbenchmarks, trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "=== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "=== bench (C2)"
timeout 900 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_s2_final.json; python scripts/show_bench.py gpurun_out/bench_s2_final.json
echo "=== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_s2_reference.json; cut -c1-400 gpurun_out/bench_s2_reference.json
echo "=== C3 nf4"
timeout 900 python bench.py --steps 1 --warmup 3 --quant nf4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_s2_nf4.json; python scripts/show_bench.py gpurun_out/bench_s2_nf4.json | head -4
echo "=== C4 geometry 720x1280"
timeout 900 python bench.py --steps 1 --warmup 3 --height 720 --width 1280 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_s2_720.json; python scripts/show_bench.py gpurun_out/bench_s2_720.json | head -4
echo "=== C5 slice q4k batch 4"
timeout 900 python bench.py --steps 1 --warmup 2 --quant q4k --batch 4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_s2_q4k_b4.json; python scripts/show_bench.py gpurun_out/bench_s2_q4k_b4.json | head -4
echo "=== microbenchmarks"
for b in mma_rate exp_rate cluster_probe; do
  [ -x scripts/ubench/$b ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Idiffusion_rs_b200/csrc scripts/ubench/$b.cu -o scripts/ubench/$b
done
scripts/ubench/mma_rate > gpurun_out/ubench_mma_rate.txt 2>&1; tail -25 gpurun_out/ubench_mma_rate.txt
scripts/ubench/exp_rate > gpurun_out/ubench_exp_rate.txt 2>&1; cat gpurun_out/ubench_exp_rate.txt
scripts/ubench/cluster_probe > gpurun_out/ubench_cluster_probe.txt 2>&1; cat gpurun_out/ubench_cluster_probe.txt
echo "=== attention variants + trace"
timeout 300 python scripts/attn_variants.py 0 1 2 3 4 2>&1 | tail -8 | cut -c1-330
} 2>&1 | tee gpurun_out/s2_final.log
