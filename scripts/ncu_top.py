"""Summarise an .ncu-rep: key metrics + top stall instructions (reads `ncu -i` output; no GPU needed)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units = r[0], r[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct",
        "launch__registers_per_thread", "launch__grid_size", "lts__throughput.avg.pct", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform", "smsp__issue_active.avg.pct", "l1tex__throughput.avg.pct",
        "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_lsu",
        "smsp__average_warp", "smsp__warps_issue_stalled"]
for row in r[2:]:
    print("==", row[hdr.index("Kernel Name")])
    for i, h in enumerate(hdr):
        if any(h.startswith(w) for w in want) and "per_second" not in h:
            print(f"  {h} = {row[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
data = []
for x in rows[2:]:
    if len(x) < len(h):
        continue
    try:
        data.append((int(x[isamp]), int(x[iex]), x[isrc].strip()))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
print("instructions", len(data), "samples", tot)
top = sorted(range(len(data)), key=lambda i: -data[i][0])[:topn]
for i in sorted(top):
    s, e, t = data[i]
    print(f"{i:5d} {100 * s / tot:5.1f}% exec={e:9d}  {t[:100]}")
