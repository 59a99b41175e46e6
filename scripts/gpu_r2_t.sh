#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python bench.py --steps 2 --warmup 2 --text-encoders --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2t_bench_text.json
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2t_bench_text.json'))
print('value', j['value'], 'e2e', j['e2e']['value'])
print('text', json.dumps(j['text_encoders']))
PY
} 2>&1 | tee gpurun_out/r2t.log
