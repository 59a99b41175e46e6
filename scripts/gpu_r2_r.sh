#!/bin/bash
# round-2 run R: ncu of the rewritten gn_apply, and of ln_modulate with the caches left alone (--cache-control none: inside
# the step its input was just written by the previous GEMM and is L2-resident; the default ncu replay flushes it)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-kernel-timing --num-steps 1 --layers 1 --single-layers 1"
{
timeout 600 ncu --set full --clock-control none -f -k "regex:gn_stats|gn_apply" -s 44 -c 12 -o gpurun_out/prof_gn2_r2 $S > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none -f -k regex:ln_modulate -s 0 -c 6 -o gpurun_out/prof_ln_hot_r2 $S > /dev/null 2>&1
for t in gn2 ln_hot; do
  echo "===== $t"
  ncu -i gpurun_out/prof_${t}_r2.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h,u=r[0],r[1]
def col(n): return h.index(n)
for row in r[2:]:
    name=row[col('Kernel Name')][:40]
    t=float(row[col('gpu__time_duration.sum')].replace(',',''))
    tu=u[col('gpu__time_duration.sum')]
    rd=row[col('dram__bytes_read.sum')]; wr=row[col('dram__bytes_write.sum')]
    lts=row[col('lts__t_bytes.sum')] if 'lts__t_bytes.sum' in h else '?'
    print(f'{name:40s} {t:10.2f} {tu:8s} dram rd {rd:>10s} {u[col(\"dram__bytes_read.sum\")]:6s} wr {wr:>10s} {u[col(\"dram__bytes_write.sum\")]:6s} lts bytes {lts} {u[col(\"lts__t_bytes.sum\")] if \"lts__t_bytes.sum\" in h else \"\"} grid {row[col(\"launch__grid_size\")]}')
"
done
rm -f gpurun_out/prof_gn2_r2.ncu-rep gpurun_out/prof_ln_hot_r2.ncu-rep
} 2>&1 | tee gpurun_out/r2r.log
