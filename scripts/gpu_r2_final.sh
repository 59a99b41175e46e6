#!/bin/bash
# round-2 final validation: what the driver runs at round end (full GPU test suite, smoke, both bench arms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "=== pytest -m gpu (all, incl. slow)"
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_final_pytest.log 2>&1; tail -4 gpurun_out/r2_final_pytest.log
grep -E "block parity|rel err|DiT step|VAE decode|u8 image|720x1280|fused vs|full depth" gpurun_out/r2_final_pytest.log | cut -c1-260
echo "=== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "=== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/r2_final_reference.json; cut -c1-600 gpurun_out/r2_final_reference.json
echo "=== bench (C2)"
timeout 900 python bench.py --steps 4 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_final_bench.json; python scripts/show_bench.py gpurun_out/r2_final_bench.json
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_final_bench.json'))
print('rooflines_other', json.dumps(j.get('rooflines_other'))[:600])
print('step_graph', j.get('step_graph'), 'joules/image', j.get('joules_per_image_per_gpu'))
PY
} 2>&1 | tee gpurun_out/r2_final.log
