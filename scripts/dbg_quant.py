import math, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_rs_b200 import build, ops, quantize as QZ
build.build()
kind = sys.argv[1] if len(sys.argv) > 1 else "nf4"
M, N, K = 256, 256, 1024
x = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
if kind == "nf4":
    packed, a8, code, nmax, off, lut = QZ.quantize_nf4(w)
    absmax = (code[a8.long()] * nmax.repeat_interleave(256)[:a8.numel()] + off).float().contiguous()
    y = ops.linear_quant(x, packed, absmax, "nf4", N)
else:
    q4 = QZ.quantize_q4k(w)
    y = ops.linear_quant(x, q4, None, "q4k", N)
torch.cuda.synchronize()
print("ok", y.float().abs().mean().item())
