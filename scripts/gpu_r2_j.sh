#!/bin/bash
# round-2 run J: ncu --set full of the same GEMM (4608x3072x15360) on the 256x256 kernel and on the 512x256 kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none -f"
{
FLUXB200_GEMM_BIG=0 timeout 300 $NCU -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/prof_big0_r2 python scripts/prof_shapes.py gemm 4608 3072 15360 > /dev/null 2>&1
FLUXB200_GEMM_BIG=1 timeout 300 $NCU -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/prof_big1_r2 python scripts/prof_shapes.py gemm 4608 3072 15360 > /dev/null 2>&1
for t in big0 big1; do
  echo "===== $t"
  ncu -i gpurun_out/prof_${t}_r2.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h,u,row=r[0],r[1],r[2]
for i,name in enumerate(h):
    n=name.lower()
    if any(k in n for k in ('time_duration','tensor','inst_executed.sum','bank_conflict','lts__t_bytes.sum','dram__bytes','cycles_active.avg','throughput.avg.pct','issue_active','grid_size','shared','smem','l1tex__data_pipe')):
        print(f'{name:96s} {row[i]:>18s} {u[i]}')
"
done
rm -f gpurun_out/prof_big*_r2.ncu-rep
} 2>&1 | tee gpurun_out/r2j.log
