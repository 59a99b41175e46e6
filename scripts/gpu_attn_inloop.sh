#!/bin/bash
# attention kernel variants measured INSIDE the DiT loop (power-capped clocks): short images (8 denoising steps)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "$@"; do
  echo "=== attn_variant $v"
  timeout 600 python bench.py --steps 1 --warmup 1 --num-steps 8 --no-cpu-baseline --attn-variant $v 2>&1 | tail -1 > gpurun_out/bench_attn_v$v.json
  python - <<PY
import json
j=json.load(open("gpurun_out/bench_attn_v$v.json"))
k=j["kernels"]
print("img ms", round(j["ms_per_step"],1), "profiled", round(j["profiled_image_ms"],1), "clk", j["clocks"]["sm_mhz"])
for n in ("gemm_tcgen05","attention_tcgen05"):
    v=k[n]; print(f"  {n:20s} ms={v['ms']:8.1f} TF/s={v['flops']/v['ms']/1e9:8.1f}")
PY
done 2>&1 | tee gpurun_out/attn_inloop.log
