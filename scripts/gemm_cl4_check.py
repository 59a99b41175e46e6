"""CL4 GEMM (pairs of pairs with W multicast): bit-exactness against the pair kernel on 512-aligned shapes + timing."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_rs_b200 import build, lib as L, ops  # noqa: E402

build.build()
lib = L.load()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10):
    for _ in range(3):
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


for M, N, K, mode in [(512, 256, 64, "plain"), (1024, 768, 512, "gelu"), (4608, 3072, 3072, "gate"), (4608, 21504, 3072, "plain"),
                      (4608, 3072, 15360, "gate"), (4608, 12288, 3072, "gelu"), (4096, 9216, 3072, "plain"),
                      (8192, 8192, 8192, "plain")]:
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    x = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    gate = torch.randn(1, N, device="cuda", generator=g).bfloat16()
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()

    def run(out):
        if mode == "gelu":
            ops.linear(x, w, b, act=ops.ACT_GELU, out=out)
        elif mode == "gate":
            out.copy_(res)
            ops.linear(x, w, b, gate=gate, rows_per_batch=M, res=out, out=out)
        else:
            ops.linear(x, w, b, out=out)

    outs, ms = [], []
    for cl4 in (0, 1):
        L.check(lib.fluxb200_set_flag(b"gemm_cl4", cl4))
        o = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        run(o)
        torch.cuda.synchronize()
        outs.append(o.clone())
        ms.append(timeit(lambda: run(o)))
    same = torch.equal(outs[0], outs[1])
    fl = 2 * M * N * K
    print(f"{M}x{N}x{K} {mode}: bit-equal={same}  pair {ms[0]*1e3:.1f} us ({fl/ms[0]/1e9:.0f} TF/s)  cl4 {ms[1]*1e3:.1f} us "
          f"({fl/ms[1]/1e9:.0f} TF/s)", flush=True)
    assert same
