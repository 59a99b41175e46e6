"""Fused-dequant GEMM (NF4 and Q4_K weights expanded inside the operand producer) on the single-block projection shape
4608 x 21504 x 3072 - for ncu (scripts/gpu_r2_profiles.sh) and a CUDA-event timing line."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_rs_b200 import build, ops  # noqa: E402
from diffusion_rs_b200 import quantize as QZ  # noqa: E402

build.build()
M, N, K = 4608, 21504, 3072
x = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
b = torch.randn(N, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for kind in ("nf4", "q4k"):
    if kind == "nf4":
        packed, a8, code8, nmax, offset, _ = QZ.quantize_nf4(w)
        # double-quantised absmax -> f32 (what fluxb200_model_finalize does once per model)
        aux = (code8[a8.long()] * nmax.repeat_interleave(256)[: a8.numel()] + offset).float().contiguous()
    else:
        packed, aux = QZ.quantize_q4k(w), None
    for _ in range(3):
        ops.linear_quant(x, packed, aux, kind, N, b, out=out)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        ops.linear_quant(x, packed, aux, kind, N, b, out=out)
    e.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 5
    print(kind, "fused-dequant GEMM", ms, "ms", 2 * M * N * K / ms / 1e9, "TFLOP/s")
