"""Kernel microbenchmarks on real FLUX shapes (CUDA-event timed, L2 flushed between iterations)."""
import json
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_rs_b200 import build, ops  # noqa: E402

build.build()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda._sleep(2_000_000)  # ~1 ms spin kernel: the launches below queue behind it, so a->b has no host gaps
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


res = []
only = sys.argv[1] if len(sys.argv) > 1 else None
for M, N, K in [(4608, 3072, 3072), (4096, 9216, 3072), (4608, 21504, 3072), (4096, 12288, 3072), (4096, 3072, 12288),
                (4608, 3072, 15360), (512, 9216, 3072), (8192, 8192, 8192)]:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = torch.randn(N, K, device="cuda").bfloat16() / math.sqrt(K)
    b = torch.randn(N, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: ops.linear(x, w, b, out=out))
    ms_t = timeit(lambda: torch.matmul(x, w.t(), out=out))
    fl = 2 * M * N * K
    res.append(dict(kind="gemm", M=M, N=N, K=K, ms=ms, tflops=fl / ms / 1e9, cublas_ms=ms_t, cublas_tflops=fl / ms_t / 1e9))
    print(res[-1], flush=True)
    if K >= 12288 or (M, N, K) == (4608, 3072, 3072):  # the shapes that run with the gate * x + residual epilogue (in place)
        gate = torch.randn(1, N, device="cuda").bfloat16()
        ms = timeit(lambda: ops.linear(x, w, b, gate=gate, rows_per_batch=M, res=out, out=out))
        res.append(dict(kind="gemm+gate+res", M=M, N=N, K=K, ms=ms, tflops=fl / ms / 1e9))
        print(res[-1], flush=True)
    if (M, N, K) == (4096, 12288, 3072):
        ms = timeit(lambda: ops.linear(x, w, b, act=ops.ACT_GELU, out=out))
        res.append(dict(kind="gemm+gelu", M=M, N=N, K=K, ms=ms, tflops=fl / ms / 1e9))
        print(res[-1], flush=True)

for B, H, L in ([] if only == "gemm" else [(1, 24, 4608), (1, 24, 4112), (1, 24, 512)]):
    q = torch.randn(B, H, L, 128, device="cuda").bfloat16()
    k = torch.randn(B, H, L, 128, device="cuda").bfloat16()
    v = torch.randn(B, H, L, 128, device="cuda").bfloat16()
    ms = timeit(lambda: ops.sdpa(q, k, v, 1 / math.sqrt(128)))
    ms_t = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
    fl = 4 * B * H * L * L * 128
    res.append(dict(kind="attn", B=B, H=H, L=L, ms=ms, tflops=fl / ms / 1e9, torch_sdpa_ms=ms_t, torch_tflops=fl / ms_t / 1e9))
    print(res[-1], flush=True)

Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/microbench.json").write_text(json.dumps(res, indent=1))
