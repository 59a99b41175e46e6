#!/bin/bash
# programmatic dependent launch A/B inside the DiT loop + correctness of the model-level tests with PDL on
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for pdl in 0 1; do
  echo "=== FLUXB200_PDL=$pdl"
  FLUXB200_PDL=$pdl timeout 600 python bench.py --steps 2 --warmup 1 --num-steps 10 --no-cpu-baseline --no-kernel-timing 2>&1 | tail -1 > gpurun_out/bench_pdl$pdl.json
  python -c "
import json; j=json.load(open('gpurun_out/bench_pdl$pdl.json')); print('img ms', round(j['ms_per_step'],1), 'e2e ms', round(j['e2e']['ms_per_step'],1), 'clk', j['clocks']['sm_mhz'])"
done
echo "=== tests (PDL on)"
timeout 900 python -m pytest tests/test_dit_gpu.py tests/test_ops_gpu.py -x -q -m gpu 2>&1 | tail -4
} 2>&1 | tee gpurun_out/pdl_ab.log
