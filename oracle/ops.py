"""ORACLE (test infrastructure, not product code) — CPU restatement of the tensor ops on the FLUX hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this.

The reference (EricLBuehler/diffusion-rs @ d977e05) is a Rust workspace and cannot be built in this environment
(no cargo/rustc), so this file restates its arithmetic op by op, each function citing the file:line it follows.
Semantics are those of the reference's CPU backend (`half` crate: every bf16 tensor op = f32 op followed by
round-to-nearest-even to bf16).  Tensors are carried as float32 holding bf16-representable values when
`Mode.round` is True ("ref_bf16" mode); with `Mode.round` False the same graph runs in plain f32 ("f32" mode,
the upper-bound truth used to judge whose rounding error is smaller).

Pinning: the op-level functions are checked against the reference's own known-answer vectors in
tests/test_oracle_golden.py (softmax / layer_norm / rms_norm: diffusion_rs_common/src/nn/tests/ops.rs:9-139,
group_norm: nn/tests/group_norm.rs:33-105, conv2d: core/tests/conv_tests.rs:126-166, Q4_K round trip:
core/tests/quantized_tests.rs:567-612).  At model level (FLUX step, VAE decode, text encoders) the reference ships
no test and no golden tensor; there the oracle's f32 graph is pinned against independent implementations of the same
networks instead (tests/test_flux_oracle_pin.py: the Black Forest Labs FLUX / autoencoder code vendored by torchtitan;
tests/test_text_oracle.py: HuggingFace T5EncoderModel / CLIPTextModel).  The bf16 ROUNDING POINTS (where the reference
rounds between ops) remain a restatement of the Rust source that nothing executable here can confirm: **parity
unpinned** at that level.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass
class Mode:
    round: bool = True  # True: mirror every bf16 rounding point of the reference; False: pure f32


REF = Mode(True)
F32 = Mode(False)


def rb(x: torch.Tensor, mode: Mode = REF) -> torch.Tensor:
    """Round-to-nearest-even to bf16 and back to f32 (what `half::bf16::from_f32` does after every op)."""
    if not mode.round:
        return x
    return x.to(torch.bfloat16).to(torch.float32)


def bf16_scalar(v: float, mode: Mode = REF) -> float:
    if not mode.round:
        return float(v)
    return float(torch.tensor(v, dtype=torch.float32).to(torch.bfloat16).to(torch.float32))


# ---------------------------------------------------------------------------------------------------------
# Linear  (diffusion_rs_backend/src/unquantized/mod.rs:34-77)
# ---------------------------------------------------------------------------------------------------------
def linear(x, w, b=None, *, fused_bias: bool, mode: Mode = REF):
    """fused_bias=True : CUDA + rank-3 + bias -> cuBLASLt with the bias as C (beta=1): one rounding (mod.rs:52-66).
    fused_bias=False: `a.matmul(w.t())` rounded to bf16, then a separate bf16 broadcast_add (mod.rs:67)."""
    y = x @ w.t()
    if b is None:
        return rb(y, mode)
    if fused_bias:
        return rb(y + b, mode)
    return rb(rb(y, mode) + b, mode)


# ---------------------------------------------------------------------------------------------------------
# Unary ops  (diffusion_rs_common/src/core/op.rs:539-578 gelu, 699-706 silu)
# ---------------------------------------------------------------------------------------------------------
SQRT_TWO_OVER_PI = 0.79788456080286535587989211986876373


def gelu(v, mode: Mode = REF):
    """bf16: 0.5*v*(1 + tanh(c*v*(1 + 0.044715*v*v))), left to right, each binary op rounded (op.rs:547-555)."""
    if not mode.round:
        return 0.5 * v * (1.0 + torch.tanh(SQRT_TWO_OVER_PI * v * (1.0 + 0.044715 * v * v)))
    half = bf16_scalar(0.5)
    c = bf16_scalar(SQRT_TWO_OVER_PI)
    k = bf16_scalar(0.044715)
    a = rb(half * v)
    p = rb(1.0 + rb(rb(k * v) * v))
    q = rb(rb(c * v) * p)
    t = rb(torch.tanh(q))
    s = rb(1.0 + t)
    return rb(a * s)


def silu(v, mode: Mode = REF):
    """v / (1 + exp(-v)) with bf16 rounding after exp, add and div (op.rs:703-705)."""
    e = rb(torch.exp(-v), mode)
    d = rb(1.0 + e, mode)
    return rb(v / d, mode)


# ---------------------------------------------------------------------------------------------------------
# Norms
# ---------------------------------------------------------------------------------------------------------
def layer_norm(x, weight=None, bias=None, eps: float = 1e-6, mode: Mode = REF):
    """Fused fast path: f32 sum / sum-of-squares, var = E[x^2] - mean^2, one rounding at the end
    (diffusion_rs_common/src/nn/ops.rs:1021-1043; CUDA reduce.cu:73-131)."""
    d = x.shape[-1]
    mean = x.sum(-1, keepdim=True) / d
    var = (x * x).sum(-1, keepdim=True) / d - mean * mean
    inv_std = 1.0 / torch.sqrt(var + eps)
    y = (x - mean) * inv_std
    if weight is not None:
        y = y * weight
    if bias is not None:
        y = y + bias
    return rb(y, mode)


def rms_norm_slow(x, weight, eps: float = 1e-6, mode: Mode = REF):
    """RmsNorm<RmsNormNonQuantized> -> LayerNorm slow path without mean removal (nn/layer_norm.rs:136-153):
    f32 normalise -> bf16, * weight -> bf16, + bias(0) -> bf16."""
    d = x.shape[-1]
    norm = (x * x).sum(-1, keepdim=True) / d
    y = rb(x / torch.sqrt(norm + eps), mode)
    y = rb(y * weight, mode)
    return rb(y + 0.0, mode)


def group_norm(x, weight, bias, groups: int, eps: float = 1e-6, mode: Mode = REF):
    """nn::GroupNorm::forward (nn/group_norm.rs:39-74): f32 two-pass mean / centred variance, cast to the input
    dtype, then `* weight` and `+ bias` as two bf16 ops.  x is NCHW (or N,C,*)."""
    shape = x.shape
    n, c = shape[0], shape[1]
    xg = x.reshape(n, groups, -1)
    hidden = xg.shape[-1]
    mean = xg.sum(2, keepdim=True) / hidden
    xc = xg - mean
    var = (xc * xc).sum(2, keepdim=True) / hidden
    y = rb(xc / torch.sqrt(var + eps), mode).reshape(shape)
    wshape = [1, c] + [1] * (len(shape) - 2)
    y = rb(y * weight.reshape(wshape), mode)
    return rb(y + bias.reshape(wshape), mode)


# ---------------------------------------------------------------------------------------------------------
# softmax / attention
# ---------------------------------------------------------------------------------------------------------
def softmax_last_dim(x, mode: Mode = REF):
    """max-subtract, exp, sum, divide; every step rounded when the tensor is bf16 (nn/ops.rs:296-329)."""
    m = x.max(-1, keepdim=True).values
    e = rb(torch.exp(rb(x - m, mode)), mode)
    if mode.round:
        # the CPU kernel accumulates the row sum in the element type (bf16) via vec_reduce_sum; the CUDA kernel
        # accumulates in f32 (reduce.cu:576).  We follow the CUDA (ground-truth GPU) behaviour: f32 sum, bf16 result.
        s = rb(e.sum(-1, keepdim=True), mode)
    else:
        s = e.sum(-1, keepdim=True)
    return rb(e / s, mode)


def sdpa_f32(q, k, v, scale: float):
    """backend ops::sdpa on CUDA/CPU (diffusion_rs_backend/src/ops.rs:252-261) with f32 inputs as called from
    model.rs:40-51: everything in f32, result cast back to bf16 by the caller."""
    att = (q @ k.transpose(-1, -2)) * scale
    att = torch.softmax(att, dim=-1)
    return att @ v


# ---------------------------------------------------------------------------------------------------------
# scalar-broadcast affine  (Tensor::affine on bf16: mul and add are cast to bf16 first; x*mul rounded, +add rounded;
# diffusion_rs_common/src/cuda_kernels/affine.cu:33, cpu_backend Affine map)
# ---------------------------------------------------------------------------------------------------------
def affine(x, mul: float, add: float, mode: Mode = REF):
    return rb(rb(x * bf16_scalar(mul, mode), mode) + bf16_scalar(add, mode), mode)
