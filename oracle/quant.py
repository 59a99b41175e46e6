"""ORACLE (test infrastructure, not product code) — numpy restatement of the quantised-weight formats on the path.

* bitsandbytes NF4 / FP4 / blockwise-int8 / LLM.int8 de-quantisation: the reference's CUDA kernel semantics
  (diffusion_rs_backend/kernels/bitsandbytes/dequant.cu:12-170, 205-214), which SURVEY N4 names as ground truth
  (the reference's CPU path mis-indexes absmax).  Nested (double-quant) absmax: bitsandbytes/mod.rs:230-239.
* GGUF Q4_K: BlockQ4K layout k_quants.rs:130-136, to_float :1568-1599, get_scale_min_k4 quantized/utils.rs:49-59.
  Cross-checked against the independent `gguf` Python package in tests/test_oracle_golden.py and pinned by the
  reference's own round-trip KAT (core/tests/quantized_tests.rs:567-612).
The quantisers here (NF4/FP4 nearest-code, Q4_K via gguf.quants) only manufacture synthetic test weights; the oracle
weight is always dequant(quant(W)).
"""
from __future__ import annotations

import json

import numpy as np

NF4_LUT = np.array([-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
                    -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
                    0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941,
                    0.7229568362236023, 1.0], dtype=np.float32)  # dDequantizeNF4 dequant.cu:39-92
# dDequantizeFP4Tree dequant.cu:12-37 : bit3 sign, bits 2..0 select the magnitude
_FP4_MAG = np.array([0.0, 5.208333333e-03, 0.66666667, 1.0, 0.33333333, 0.5, 0.16666667, 0.25], dtype=np.float32)
FP4_LUT = np.concatenate([_FP4_MAG, -_FP4_MAG]).astype(np.float32)


def bf16_round(x: np.ndarray) -> np.ndarray:
    """f32 -> bf16 (round to nearest even) -> f32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)


def dequant_4bit(packed: np.ndarray, absmax: np.ndarray, blocksize: int, n: int, kind: str) -> np.ndarray:
    """kDequantizeBlockwise<T, ..., FP4|NF4> (dequant.cu:94-170): high nibble first; f32 result (cast by the caller)."""
    lut = NF4_LUT if kind == "nf4" else FP4_LUT
    p = packed.reshape(-1).astype(np.uint8)
    out = np.empty(p.size * 2, dtype=np.float32)
    am = absmax.astype(np.float32)[np.arange(p.size) // (blocksize // 2)]
    out[0::2] = lut[p >> 4] * am
    out[1::2] = lut[p & 15] * am
    return out[:n]


def dequant_blockwise_int8(code: np.ndarray, q: np.ndarray, absmax: np.ndarray, blocksize: int) -> np.ndarray:
    """General8bit branch (dequant.cu:135-140): code[q] * absmax[i / blocksize]."""
    q = q.reshape(-1).astype(np.uint8)
    return code.astype(np.float32)[q] * absmax.astype(np.float32)[np.arange(q.size) // blocksize]


def dequant_int8_rowwise(w: np.ndarray, scb: np.ndarray) -> np.ndarray:
    """dequantize_8bit_kernel (dequant.cu:205-214): w * SCB[row] / 127."""
    return (w.astype(np.float32) * scb.astype(np.float32)[:, None]) / np.float32(127.0)


def nested_absmax(absmax_u8, nested_code, nested_absmax_f32, nested_blocksize, offset):
    """BnbLinear::dequantize_4bit (bitsandbytes/mod.rs:230-239): int8-blockwise dequant of absmax, + offset (f32)."""
    a = dequant_blockwise_int8(nested_code, absmax_u8, nested_absmax_f32, nested_blocksize)
    return (a + np.float32(offset)).astype(np.float32)


# ---- synthetic quantisers (test-weight manufacture only) -------------------------------------------------
def quantize_4bit(w: np.ndarray, blocksize: int, kind: str):
    lut = NF4_LUT if kind == "nf4" else FP4_LUT
    flat = w.reshape(-1).astype(np.float32)
    assert flat.size % blocksize == 0
    blocks = flat.reshape(-1, blocksize)
    absmax = np.abs(blocks).max(1).astype(np.float32)
    scaled = blocks / np.where(absmax == 0, 1, absmax)[:, None]
    idx = np.abs(scaled[..., None] - lut[None, None, :]).argmin(-1).astype(np.uint8).reshape(-1)
    packed = ((idx[0::2] << 4) | idx[1::2]).astype(np.uint8)
    return packed, absmax


def quantize_absmax_nested(absmax: np.ndarray, nested_blocksize: int = 256):
    """Double quantisation of absmax into u8 codes over a 256-entry linear code book (synthetic)."""
    offset = float(absmax.mean())
    a = absmax - np.float32(offset)
    pad = (-a.size) % nested_blocksize
    ap = np.concatenate([a, np.zeros(pad, np.float32)]).reshape(-1, nested_blocksize)
    nmax = np.abs(ap).max(1).astype(np.float32)
    code = np.linspace(-1.0, 1.0, 256).astype(np.float32)
    scaled = ap / np.where(nmax == 0, 1, nmax)[:, None]
    q = np.rint((scaled + 1.0) * 127.5).clip(0, 255).astype(np.uint8).reshape(-1)[:a.size]
    return q, code, nmax, offset


def quant_state_json(blocksize, shape, nested_blocksize=None, nested_offset=None, dtype="bfloat16") -> bytes:
    d = {"blocksize": blocksize, "shape": list(shape), "dtype": dtype}
    if nested_blocksize is not None:
        d.update(nested_blocksize=nested_blocksize, nested_offset=nested_offset, nested_dtype="float32")
    return json.dumps(d).encode()


# ---- GGUF Q4_K -------------------------------------------------------------------------------------------
QK_K = 256
Q4K_BYTES = 144


def get_scale_min_k4(j: int, q: np.ndarray):
    """quantized/utils.rs:49-59 — q: the 12 packed scale bytes of one super-block (vectorised over blocks)."""
    if j < 4:
        return q[..., j] & 63, q[..., j + 4] & 63
    d = (q[..., j + 4] & 0xF) | ((q[..., j - 4] >> 6) << 4)
    m = (q[..., j + 4] >> 4) | ((q[..., j] >> 6) << 4)
    return d, m


def dequant_q4k(blocks: np.ndarray) -> np.ndarray:
    """BlockQ4K::to_float (k_quants.rs:1568-1599). blocks: u8 [..., 144] -> f32 [..., 256]."""
    b = blocks.reshape(-1, Q4K_BYTES)
    d = b[:, 0:2].copy().view(np.float16).astype(np.float32)[:, 0]
    dmin = b[:, 2:4].copy().view(np.float16).astype(np.float32)[:, 0]
    sc = b[:, 4:16]
    qs = b[:, 16:144]
    out = np.empty((b.shape[0], QK_K), dtype=np.float32)
    for g in range(4):  # 64 weights per group: low nibbles then high nibbles of 32 bytes
        q = qs[:, g * 32:(g + 1) * 32]
        s1, m1 = get_scale_min_k4(2 * g, sc)
        s2, m2 = get_scale_min_k4(2 * g + 1, sc)
        d1, mm1 = d * s1.astype(np.float32), dmin * m1.astype(np.float32)
        d2, mm2 = d * s2.astype(np.float32), dmin * m2.astype(np.float32)
        out[:, g * 64:g * 64 + 32] = d1[:, None] * (q & 0xF).astype(np.float32) - mm1[:, None]
        out[:, g * 64 + 32:g * 64 + 64] = d2[:, None] * (q >> 4).astype(np.float32) - mm2[:, None]
    return out.reshape(*blocks.shape[:-1], QK_K)


def dequant_q4k_bf16(blocks: np.ndarray) -> np.ndarray:
    """GgufMatMul::dequantize_w (gguf/mod.rs:29-31): dequantize -> f16 -> bf16 (both roundings), returned as f32."""
    f = dequant_q4k(blocks).astype(np.float16).astype(np.float32)
    return bf16_round(f)


def quantize_q4k(w: np.ndarray) -> np.ndarray:
    """Synthetic Q4_K quantiser (simple min/max per 32-weight sub-block; the `gguf` package only ships the
    de-quantiser).  w: f32 [..., K] with K % 256 == 0 -> u8 [..., K/256, 144] in the BlockQ4K byte layout."""
    x = np.ascontiguousarray(w, dtype=np.float32).reshape(-1, 8, 32)  # [blocks, sub-block, 32]
    mn = np.minimum(x.min(2), 0.0)
    mx = x.max(2)
    scale_f = (mx - mn) / 15.0
    min_f = -mn
    d = (scale_f.max(1) / 63.0).astype(np.float16).astype(np.float32)
    dmin = (min_f.max(1) / 63.0).astype(np.float16).astype(np.float32)
    sc = np.where(d[:, None] > 0, np.rint(scale_f / np.where(d == 0, 1, d)[:, None]), 0).clip(0, 63).astype(np.uint8)
    m = np.where(dmin[:, None] > 0, np.rint(min_f / np.where(dmin == 0, 1, dmin)[:, None]), 0).clip(0, 63).astype(np.uint8)
    eff = d[:, None] * sc.astype(np.float32)
    q = np.where(eff[..., None] > 0, np.rint((x + (dmin[:, None] * m)[..., None]) / np.where(eff == 0, 1, eff)[..., None]), 0)
    q = q.clip(0, 15).astype(np.uint8)
    nb = x.shape[0]
    out = np.zeros((nb, Q4K_BYTES), dtype=np.uint8)
    out[:, 0:2] = d.astype(np.float16).view(np.uint8).reshape(nb, 2)
    out[:, 2:4] = dmin.astype(np.float16).view(np.uint8).reshape(nb, 2)
    scb = np.zeros((nb, 12), dtype=np.uint8)
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
    return out.reshape(*w.shape[:-1], w.shape[-1] // QK_K, Q4K_BYTES)


# ---- the reference's own Q4_K quantiser, restated so that its round-trip KAT can pin this file ---------------
def _nearest_int(v) -> int:
    """quantized/utils.rs:3-5 — f32::round: half away from zero."""
    v = float(v)
    import math
    return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)


def make_qkx1_quants(nmax: int, ntry: int, x: np.ndarray):
    """quantized/utils.rs:227-284 (sequential f32 arithmetic)."""
    f = np.float32
    x = x.astype(np.float32)
    n = x.size
    l = [0] * n
    mn, mx = f(x.min()), f(x.max())
    if mx == mn:
        return f(0.0), f(0.0)
    mn = f(min(mn, f(0.0)))
    iscale = f(f(nmax) / f(mx - mn))
    scale = f(f(1.0) / iscale)
    for _ in range(ntry):
        sumlx, suml2, did_change = f(0.0), 0, False
        for i in range(n):
            li = max(0, min(nmax, _nearest_int(f(iscale * f(x[i] - mn)))))
            if li != l[i]:
                l[i] = li
                did_change = True
            sumlx = f(sumlx + f(f(x[i] - mn) * f(li)))
            suml2 += li * li
        scale = f(sumlx / f(suml2))
        s = f(0.0)
        for i in range(n):
            s = f(s + f(x[i] - f(scale * f(l[i]))))
        mn = f(s / f(n))
        if mn > 0:
            mn = f(0.0)
        iscale = f(f(1.0) / scale)
        if not did_change:
            break
    return scale, f(-mn)


def quantize_q4k_reference(xs: np.ndarray) -> np.ndarray:
    """BlockQ4K::from_float (k_quants.rs:1432-1494). xs: f32 [n], n % 256 == 0 -> u8 [n/256, 144]."""
    f = np.float32
    xs = np.ascontiguousarray(xs, dtype=np.float32).reshape(-1, QK_K)
    out = np.zeros((xs.shape[0], Q4K_BYTES), dtype=np.uint8)
    for b, x in enumerate(xs):
        scales, mins = [], []
        for j in range(8):
            s, m = make_qkx1_quants(15, 5, x[32 * j:32 * j + 32])
            scales.append(s)
            mins.append(m)
        max_scale = f(max([f(0.0)] + scales))
        max_min = f(max([f(0.0)] + mins))
        inv_scale = f(f(63.0) / max_scale) if max_scale > 0 else f(0.0)
        inv_min = f(f(63.0) / max_min) if max_min > 0 else f(0.0)
        sc = np.zeros(12, dtype=np.uint8)
        for j in range(8):
            ls = min(_nearest_int(f(inv_scale * scales[j])), 63)
            lm = min(_nearest_int(f(inv_min * mins[j])), 63)
            if j < 4:
                sc[j] = ls
                sc[j + 4] = lm
            else:
                sc[j + 4] = (ls & 0xF) | ((lm & 0xF) << 4)
                sc[j - 4] |= (ls >> 4) << 6
                sc[j] |= (lm >> 4) << 6
        d16 = np.float16(f(max_scale / f(63.0)))
        m16 = np.float16(f(max_min / f(63.0)))
        out[b, 0:2] = np.frombuffer(d16.tobytes(), dtype=np.uint8)
        out[b, 2:4] = np.frombuffer(m16.tobytes(), dtype=np.uint8)
        out[b, 4:16] = sc
        l = np.zeros(QK_K, dtype=np.uint8)
        for j in range(8):
            s6, m6 = get_scale_min_k4(j, sc)
            d = f(f(d16) * f(int(s6)))
            if d != 0:
                dm = f(f(m16) * f(int(m6)))
                for ii in range(32):
                    l[32 * j + ii] = max(0, min(15, _nearest_int(f(f(x[32 * j + ii] + dm) / d))))
        for g in range(4):
            out[b, 16 + g * 32:16 + (g + 1) * 32] = l[g * 64:g * 64 + 32] | (l[g * 64 + 32:g * 64 + 64] << 4)
    return out
