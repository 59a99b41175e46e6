"""Operator-level host mirror of the reference's backend interface (QuantMethod::forward, ops::sdpa, ...).

Thin argument marshalling over the C ABI; tensors are torch CUDA bf16 tensors (torch = device memory plumbing).
"""
from __future__ import annotations

import torch

from . import lib as L

BIAS_NONE, BIAS_FUSED, BIAS_AFTER_ROUND = 0, 1, 2
ACT_NONE, ACT_GELU = 0, 1


def _chk_bf16(*ts):
    for t in ts:
        if t is not None:
            if not t.is_cuda:
                raise L.Fluxb200Error("fluxb200 ops need CUDA tensors (no CPU fallback)")
            if t.dtype != torch.bfloat16:
                raise L.Fluxb200Error(f"expected bf16, got {t.dtype}")
            if not t.is_contiguous():
                raise L.Fluxb200Error("input has to be contiguous")


def linear(x, w, bias=None, *, bias_mode=BIAS_FUSED, act=ACT_NONE, gate=None, rows_per_batch=0, res=None,
           alpha=1.0, out=None):
    """out = epilogue(x @ w.T) — UnquantLinear::forward (unquantized/mod.rs:34-77) plus the fused epilogue."""
    _chk_bf16(x, w, bias, gate, res)
    K = x.shape[-1]
    M = x.numel() // K
    N = w.shape[0]
    if w.shape[1] != K:
        raise L.Fluxb200Error(f"shape mismatch in linear: x[...,{K}] vs w{tuple(w.shape)}")
    if out is None:
        out = torch.empty(*x.shape[:-1], N, device=x.device, dtype=torch.bfloat16)
    gate_bstride = gate.stride(0) if gate is not None and gate.dim() > 1 else 0
    L.check(L.load().fluxb200_linear(L.ptr(x), K, L.ptr(w), K, L.ptr(bias), L.ptr(out), N, M, N, K,
                                     bias_mode if bias is not None else BIAS_NONE, act, L.ptr(gate), gate_bstride,
                                     rows_per_batch, L.ptr(res), float(alpha), L.current_stream()))
    return out


QUANT_KINDS = {"nf4": 1, "fp4": 2, "q4k": 3, "int8": 4}


def linear_quant(x, packed, aux, kind: str, N: int, bias=None, *, blocksize=64, bias_mode=BIAS_AFTER_ROUND,
                 act=ACT_NONE, out=None):
    """out = epilogue(x @ dequant(packed).T) with the weight expanded inside the GEMM — BnbLinear / GgufMatMul."""
    _chk_bf16(x, bias)
    K = x.shape[-1]
    M = x.numel() // K
    if out is None:
        out = torch.empty(*x.shape[:-1], N, device=x.device, dtype=torch.bfloat16)
    L.check(L.load().fluxb200_linear_quant(L.ptr(x), K, L.ptr(packed), L.ptr(aux), QUANT_KINDS[kind], blocksize,
                                           L.ptr(bias), L.ptr(out), N, M, N, K,
                                           bias_mode if bias is not None else BIAS_NONE, act, L.current_stream()))
    return out


def sdpa(q, k, v, scale: float, softcapping: float = 1.0):
    """q,k,v [B,H,L,128] -> [B,L,H*128] — ops::sdpa(q, k, v, scale, softcapping) (ops.rs:247-262) followed by the
    transpose(1,2).flatten_from(2) of its FLUX call site (model.rs:101).  head_dim != 128 or softcapping != 1.0 are
    rejected by the library (a Rust shim falls through to the stock ops::sdpa on that status)."""
    _chk_bf16(q, k, v)
    B, H, Lq, D = q.shape
    out = torch.empty(B, Lq, H * D, device=q.device, dtype=torch.bfloat16)
    L.check(L.load().fluxb200_sdpa(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(out), B, H, Lq, D, float(scale),
                                   float(softcapping), L.current_stream()))
    return out


def layernorm_modulate(x, shift, scale, eps=1e-6):
    """x [B,T,3072], shift/scale [B,3072] (may be strided views of the modulation output)."""
    _chk_bf16(x)
    B, T, D = x.shape
    out = torch.empty_like(x)
    assert shift.stride(-1) == 1 and scale.stride(-1) == 1 and shift.stride(0) == scale.stride(0)
    L.check(L.load().fluxb200_layernorm_modulate(L.ptr(x), L.ptr(shift), L.ptr(scale), shift.stride(0), L.ptr(out), B,
                                                 T, D, float(eps), L.current_stream()))
    return out


def qknorm_rope(qkv, wq, wk, pe_cos, pe_sin, H, L_total, l_off, Q, K, V, eps=1e-6):
    """qkv [B,T,3*H*128] -> writes Q,K,V [B,H,L_total,128] at sequence offset l_off."""
    _chk_bf16(qkv, wq, wk, pe_cos, pe_sin, Q, K, V)
    B, T, ld = qkv.shape
    L.check(L.load().fluxb200_qknorm_rope(L.ptr(qkv), ld, B, T, H, L_total, l_off, L.ptr(wq), L.ptr(wk),
                                          L.ptr(pe_cos), L.ptr(pe_sin), L.ptr(Q), L.ptr(K), L.ptr(V), float(eps),
                                          L.current_stream()))
