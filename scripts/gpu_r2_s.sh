#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_vae_quant_gpu.py -m "gpu and not slow" -q -k "groupnorm or vae_decode or packed or conv" 2>&1 | tail -3
timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2s_bench.json; python scripts/show_bench.py gpurun_out/r2s_bench.json | grep -E "^value|groupnorm|clocks"
timeout 600 ncu --set full --clock-control none -f -k "regex:gn_apply" -s 24 -c 4 -o gpurun_out/prof_gn3_r2 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-kernel-timing --num-steps 1 --layers 1 --single-layers 1 > /dev/null 2>&1
ncu -i gpurun_out/prof_gn3_r2.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h,u=r[0],r[1]
c=lambda n: h.index(n)
for row in r[2:]:
    print(row[c('Kernel Name')][:30], row[c('gpu__time_duration.sum')], u[c('gpu__time_duration.sum')], 'rd', row[c('dram__bytes_read.sum')], u[c('dram__bytes_read.sum')], 'wr', row[c('dram__bytes_write.sum')], u[c('dram__bytes_write.sum')], 'xu%', row[c('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active')])
"
rm -f gpurun_out/prof_gn3_r2.ncu-rep
} 2>&1 | tee gpurun_out/r2s.log
