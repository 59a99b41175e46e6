"""Run one op on a real FLUX shape a few times (for ncu) — usage: prof_shapes.py gemm M N K [gelu|gate] | attn B H L"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_rs_b200 import build, ops  # noqa: E402

build.build()
kind = sys.argv[1]
if kind == "gemm":
    M, N, K = map(int, sys.argv[2:5])
    mode = sys.argv[5] if len(sys.argv) > 5 else "plain"
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16()
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    gate = torch.randn(1, N, device="cuda").bfloat16()
    for _ in range(4):
        if mode == "gelu":
            ops.linear(x, w, b, act=ops.ACT_GELU, out=out)
        elif mode == "gate":
            ops.linear(x, w, b, gate=gate, rows_per_batch=M, res=out, out=out)
        else:
            ops.linear(x, w, b, out=out)
else:
    B, H, L = map(int, sys.argv[2:5])
    q = torch.randn(B, H, L, 128, device="cuda").bfloat16()
    k = torch.randn(B, H, L, 128, device="cuda").bfloat16()
    v = torch.randn(B, H, L, 128, device="cuda").bfloat16()
    for _ in range(4):
        ops.sdpa(q, k, v, 1 / math.sqrt(128))
torch.cuda.synchronize()
