#!/bin/bash
# round-2 run M: C2 at 8 GPUs with every rank's own time (bench.py ms_per_step_by_rank) - is the 8-GPU loss one slow chip?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --no-cpu-baseline --no-kernel-timing"
{
timeout 900 $T --steps 3 --warmup 2 2>&1 | tail -1 > gpurun_out/r2m_bench_c2_8gpu.json
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2m_bench_c2_8gpu.json'))
print('value', j['value'], 'ms_per_step', j['ms_per_step'], 'e2e', j['e2e']['value'])
print('by rank', j.get('ms_per_step_by_rank'))
print('clocks', j.get('clocks'))
PY
nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu --format=csv,noheader
} 2>&1 | tee gpurun_out/r2m.log
