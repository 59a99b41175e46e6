"""Yardstick kernels for ncu (NOT product code, nothing here is on any fluxb200 path): cuBLAS bf16 GEMM and torch's
fused SDPA (cuDNN / flash) on the two dominant FLUX shapes, so that their launch geometry (grid, cluster, shared memory,
registers) and pipe utilisation can be read next to ours.  usage: prof_yardsticks.py gemm M N K | attn"""
import math
import sys

import torch

kind = sys.argv[1]
if kind == "gemm":
    M, N, K = (int(a) for a in sys.argv[2:5])
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(x, w.t(), out=out)
else:
    q = torch.randn(1, 24, 4608, 128, device="cuda").bfloat16()
    k = torch.randn(1, 24, 4608, 128, device="cuda").bfloat16()
    v = torch.randn(1, 24, 4608, 128, device="cuda").bfloat16()
    for _ in range(3):
        torch.nn.functional.scaled_dot_product_attention(q, k, v)
torch.cuda.synchronize()
