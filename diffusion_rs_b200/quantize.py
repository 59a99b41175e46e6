"""Manufacture synthetic *quantised* checkpoints (bitsandbytes NF4 with double quantisation, GGUF Q4_K) from bf16
tensors, on whatever device the tensor lives on.  Only used to build the C3 / C5 benchmark configurations: real
checkpoints arrive already quantised.  Tensor naming follows what `BnbLinear::linear_4bit` reads
(diffusion_rs_backend/src/bitsandbytes/mod.rs:137-222) and the Q4_K block layout of k_quants.rs:130-136."""
from __future__ import annotations

import json

import torch

DT_Q4K = 10

NF4_LUT = [-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
           -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
           0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0]


def is_block_linear_weight(name: str, t: torch.Tensor) -> bool:
    return name.endswith(".weight") and t.dim() == 2 and "transformer_blocks" in name and ".norm_" not in name


def quantize_nf4(w: torch.Tensor, blocksize: int = 64, nested_blocksize: int = 256):
    dev = w.device
    lut = torch.tensor(NF4_LUT, dtype=torch.float32, device=dev)
    flat = w.reshape(-1).float()
    blocks = flat.reshape(-1, blocksize)
    absmax = blocks.abs().amax(1)
    scaled = blocks / torch.where(absmax == 0, torch.ones_like(absmax), absmax)[:, None]
    idx = torch.empty(scaled.shape, dtype=torch.uint8, device=dev)
    for s in range(0, scaled.shape[0], 1 << 16):  # chunked nearest-code search
        e = min(scaled.shape[0], s + (1 << 16))
        idx[s:e] = (scaled[s:e, :, None] - lut[None, None, :]).abs().argmin(-1).to(torch.uint8)
    idx = idx.reshape(-1)
    packed = ((idx[0::2] << 4) | idx[1::2]).to(torch.uint8)
    # double quantisation of absmax (synthetic linear 256-entry code book)
    offset = float(absmax.mean())
    a = absmax - offset
    pad = (-a.numel()) % nested_blocksize
    ap = torch.cat([a, torch.zeros(pad, device=dev)]).reshape(-1, nested_blocksize)
    nmax = ap.abs().amax(1)
    code = torch.linspace(-1.0, 1.0, 256, device=dev, dtype=torch.float32)
    sc = ap / torch.where(nmax == 0, torch.ones_like(nmax), nmax)[:, None]
    q = torch.round((sc + 1.0) * 127.5).clamp(0, 255).to(torch.uint8).reshape(-1)[:a.numel()]
    return packed, q, code, nmax, offset, lut


def quantize_q4k(w: torch.Tensor) -> torch.Tensor:
    """bf16/f32 [N, K] (K % 256 == 0) -> u8 [N*K/256*144] in the BlockQ4K byte layout (simple min/max quantiser)."""
    dev = w.device
    x = w.float().reshape(-1, 8, 32)
    mn = torch.clamp(x.amin(2), max=0.0)
    mx = x.amax(2)
    scale_f = (mx - mn) / 15.0
    min_f = -mn
    d = (scale_f.amax(1) / 63.0).half().float()
    dmin = (min_f.amax(1) / 63.0).half().float()
    one = torch.ones_like(d)
    sc = torch.where(d[:, None] > 0, torch.round(scale_f / torch.where(d == 0, one, d)[:, None]), torch.zeros_like(scale_f))
    m = torch.where(dmin[:, None] > 0, torch.round(min_f / torch.where(dmin == 0, one, dmin)[:, None]), torch.zeros_like(min_f))
    sc = sc.clamp(0, 63).to(torch.uint8)
    m = m.clamp(0, 63).to(torch.uint8)
    eff = d[:, None] * sc.float()
    q = torch.where(eff[..., None] > 0,
                    torch.round((x + (dmin[:, None] * m.float())[..., None]) / torch.where(eff == 0, torch.ones_like(eff), eff)[..., None]),
                    torch.zeros_like(x)).clamp(0, 15).to(torch.uint8)
    nb = x.shape[0]
    out = torch.zeros(nb, 144, dtype=torch.uint8, device=dev)
    out[:, 0:2] = d.half().view(torch.uint8).reshape(nb, 2)
    out[:, 2:4] = dmin.half().view(torch.uint8).reshape(nb, 2)
    scb = torch.zeros(nb, 12, dtype=torch.uint8, device=dev)
    for j in range(4):
        scb[:, j] = sc[:, j] & 63
        scb[:, j + 4] = m[:, j] & 63
    for j in range(4, 8):
        scb[:, j + 4] = (sc[:, j] & 0xF) | ((m[:, j] & 0xF) << 4)
        scb[:, j - 4] |= (sc[:, j] >> 4) << 6
        scb[:, j] |= (m[:, j] >> 4) << 6
    out[:, 4:16] = scb
    for g in range(4):
        out[:, 16 + g * 32:16 + (g + 1) * 32] = q[:, 2 * g] | (q[:, 2 * g + 1] << 4)
    return out.reshape(-1)


def quantize_tensor(name: str, t: torch.Tensor, kind: str):
    """Yield (tensor_name, tensor, dtype_code|None, logical_shape|None) for one checkpoint tensor."""
    if not is_block_linear_weight(name, t):
        yield name, t, None, None
        return
    N, K = t.shape
    if kind == "q4k":
        yield name, quantize_q4k(t), DT_Q4K, (N, K)
    elif kind == "nf4":
        packed, a8, code, nmax, offset, lut = quantize_nf4(t)
        yield name, packed.reshape(-1, 1), None, None
        yield name + ".absmax", a8, None, None
        yield name + ".quant_map", lut, None, None
        yield name + ".nested_absmax", nmax, None, None
        yield name + ".nested_quant_map", code, None, None
        js = json.dumps({"blocksize": 64, "shape": [N, K], "dtype": "bfloat16", "nested_blocksize": 256,
                         "nested_offset": offset, "nested_dtype": "float32"}).encode()
        yield name + ".quant_state.bitsandbytes__nf4", torch.tensor(list(js), dtype=torch.uint8, device=t.device), None, None
    else:
        raise ValueError(f"unknown quantisation kind {kind!r}")
