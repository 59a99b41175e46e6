"""Where does the GEMM's MMA thread wait?  clock64 trace of scheduling unit 0 (fluxb200_debug_gemm_trace).

The instrumentation is compiled out of production builds: run as  FLUXB200_GEMM_TRACE=1 python scripts/gemm_trace.py
(and rebuild without the variable afterwards)."""
import os
os.environ.setdefault("FLUXB200_GEMM_TRACE", "1")
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_rs_b200 import build, lib as L, ops  # noqa: E402

build.build()
lib = L.load()
trace = torch.zeros(64 * 4, dtype=torch.int64, device="cuda")
BIG = int(os.environ.get("FLUXB200_GEMM_BIG", "0"))
shapes = [(4608, 3072, 15360, "plain"), (4096, 3072, 12288, "gate"), (4608, 21504, 3072, "plain"), (4608, 3072, 3072, "gate"),
          (8192, 8192, 8192, "plain")]
for M, N, K, mode in shapes:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16()
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    gate = torch.randn(1, N, device="cuda").bfloat16()

    def run():
        if mode == "gelu":
            ops.linear(x, w, b, act=ops.ACT_GELU, out=out)
        elif mode == "gate":
            ops.linear(x, w, b, gate=gate, rows_per_batch=M, res=out, out=out)
        else:
            ops.linear(x, w, b, out=out)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    trace.zero_()
    L.check(lib.fluxb200_debug_gemm_trace(trace.data_ptr()))
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run()
    e.record()
    torch.cuda.synchronize()
    L.check(lib.fluxb200_debug_gemm_trace(None))
    t = trace.cpu().view(64, 4).clone()
    n = int((t[:, 3] > 0).sum())
    t = t[:n]
    kb = K // 64
    if BIG and (BIG in (2, 4) or K >= 8192):  # the 512x256 kernel packs (item total x 4 + sub-tile count) into field 3
        nsub = t[:, 3] % 4
        t[:, 3] //= 4
        rows = [(int(nsub[i]), int(t[i, 1]), int(t[i, 2]), int(t[i, 3]), kb * 4 * 128 * int(nsub[i])) for i in range(n)]
        print(f"{M}x{N}x{K} {mode} BIG: {a.elapsed_time(e)*1e3:.1f} us, {2*M*N*K/a.elapsed_time(e)/1e9:.0f} TFLOP/s; unit 0 items "
              f"(sub-tiles, wait accumulator, wait TMA, total, MMA floor): {rows}")
        continue
    floor = kb * 4 * 128
    ms = a.elapsed_time(e)
    span = int(t[-1, 0] + t[-1, 3] - t[0, 0])
    print(f"{M}x{N}x{K} {mode}: {ms*1e3:.1f} us, {2*M*N*K/ms/1e9:.0f} TFLOP/s; unit 0 ran {n} tiles; MMA floor/tile {floor} clk; "
          f"tile total avg {t[1:, 3].float().mean():.0f} (first {int(t[0,3])}); wait accumulator avg {t[1:, 1].float().mean():.0f}; "
          f"wait TMA avg {t[1:, 2].float().mean():.0f} (first tile {int(t[0,2])}); span {span} clk vs {n*floor} floor "
          f"= {100*n*floor/span:.1f}%")
