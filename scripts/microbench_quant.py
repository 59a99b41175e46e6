"""Fused-dequant GEMM microbenchmark (CUDA-event timed, launches queued behind a spin kernel)."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_rs_b200 import build, ops, quantize as QZ  # noqa: E402

build.build()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=8, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda._sleep(2_000_000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


for M, N, K in [(4608, 9216, 3072), (4608, 21504, 3072), (4608, 3072, 15360), (512, 9216, 3072)]:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fl = 2 * M * N * K
    ms = timeit(lambda: ops.linear(x, w, None, out=out))
    print(f"dense  {M}x{N}x{K}: {ms*1e3:8.1f} us {fl/ms/1e9:8.1f} TFLOP/s", flush=True)
    packed, a8, code, nmax, off, lut = QZ.quantize_nf4(w)
    absmax = (code[a8.long()] * nmax.repeat_interleave(256)[:a8.numel()] + off).float().contiguous()
    ms = timeit(lambda: ops.linear_quant(x, packed, absmax, "nf4", N, out=out))
    print(f"nf4    {M}x{N}x{K}: {ms*1e3:8.1f} us {fl/ms/1e9:8.1f} TFLOP/s", flush=True)
    q4 = QZ.quantize_q4k(w)
    ms = timeit(lambda: ops.linear_quant(x, q4, None, "q4k", N, out=out))
    print(f"q4_k   {M}x{N}x{K}: {ms*1e3:8.1f} us {fl/ms/1e9:8.1f} TFLOP/s", flush=True)
