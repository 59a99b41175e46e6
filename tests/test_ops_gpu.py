"""GPU parity tests of the operator-level C ABI against the oracle (oracle/ops.py) on seeded inputs."""
import math

import pytest
import torch

from oracle import ops as O

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16)


def _rand(shape, seed, scale=1.0, device="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return _bf(torch.randn(*shape, generator=g) * scale).to(device)


def _err(a, b):
    a = a.float()
    b = b.float()
    diff = (a - b).abs()
    return diff.max().item(), (diff.norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 3072), (4608, 3072, 3072), (512, 12288, 3072),
                                   (200, 320, 192), (4096, 64, 3072), (4096, 3072, 64)])
def test_linear_plain(fluxlib, M, N, K):
    from diffusion_rs_b200 import ops
    x = _rand((M, K), 1)
    w = _rand((N, K), 2, 1.0 / math.sqrt(K))
    b = _rand((N,), 3, 0.02)
    y = ops.linear(x, w, b, bias_mode=ops.BIAS_FUSED)
    torch.cuda.synchronize()
    ref = O.linear(x.float(), w.float(), b.float(), fused_bias=True)
    mx, rel = _err(y, ref)
    # same products, different fp32 summation order: results differ by at most one bf16 ulp on a few elements
    assert rel < 2e-3, (mx, rel)
    mism = (y.float() != ref).float().mean().item()
    assert mism < 0.02, mism


@pytest.mark.parametrize("M,N,K,flag", [(4608, 3072, 12288, 1), (4112, 3072, 8192, 1), (4608, 3072, 15360, 1),
                                        (2500, 12288, 3072, 2), (4608, 3072, 12288, 3), (4112, 3328, 8192, 3),
                                        (2500, 12288, 3072, 4)])
def test_linear_big_tiles_bit_identical(fluxlib, M, N, K, flag):
    """flag 1/2: 512x256 items (two sub-tiles stacked along M share W), 3/4: 256x512 items (side by side along N, share A;
    3328 columns = 13 tiles: odd edge).  The 512x256-per-CTA-pair kernel for the long-K GEMMs (hybrid work list: full waves of big tiles + 256x256 halves)
    accumulates every output element over k in the same order as the 256x256 kernel: same bits, including the fused
    gate * x + residual epilogue and ragged / odd M edges (4112 rows = 33 tiles of 128)."""
    from diffusion_rs_b200 import lib as L
    from diffusion_rs_b200 import ops
    x = _rand((M, K), 11)
    w = _rand((N, K), 12, 1.0 / math.sqrt(K))
    b = _rand((N,), 13, 0.02)
    gate = _rand((2, N), 14, 0.5)
    res = _rand((M, N), 15)
    rpb = (M + 1) // 2
    outs = []
    for f in (0, flag):
        L.check(fluxlib.fluxb200_set_flag(b"gemm_big", f))
        if flag in (2, 4):
            outs.append(ops.linear(x, w, b, bias_mode=ops.BIAS_FUSED, act=ops.ACT_GELU))
        else:
            outs.append(ops.linear(x, w, b, bias_mode=ops.BIAS_FUSED, gate=gate, rows_per_batch=rpb, res=res))
    L.check(fluxlib.fluxb200_set_flag(b"gemm_big", 0))
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    if flag in (1, 3) and M == 4112:  # and the result is right, not just equal
        b_idx = (torch.arange(M, device="cuda") // rpb)
        ref = O.rb(res.float() + O.rb(gate.float()[b_idx] * O.linear(x.float(), w.float(), b.float(), fused_bias=True)))
        assert _err(outs[1], ref)[1] < 2e-3


def test_linear_bias_after_round_and_nobias(fluxlib):
    from diffusion_rs_b200 import ops
    M, N, K = 256, 512, 256
    x = _rand((M, K), 4)
    w = _rand((N, K), 5, 1.0 / math.sqrt(K))
    b = _rand((N,), 6, 0.5)
    y = ops.linear(x, w, b, bias_mode=ops.BIAS_AFTER_ROUND)
    ref = O.linear(x.float(), w.float(), b.float(), fused_bias=False)
    assert _err(y, ref)[1] < 2e-3
    y = ops.linear(x, w, None)
    ref = O.linear(x.float(), w.float(), None, fused_bias=False)
    assert _err(y, ref)[1] < 2e-3


def test_linear_gelu(fluxlib):
    from diffusion_rs_b200 import ops
    M, N, K = 512, 1024, 512
    x = _rand((M, K), 7)
    w = _rand((N, K), 8, 2.0 / math.sqrt(K))
    b = _rand((N,), 9, 0.1)
    y = ops.linear(x, w, b, bias_mode=ops.BIAS_FUSED, act=ops.ACT_GELU)
    pre = O.linear(x.float(), w.float(), b.float(), fused_bias=True)
    ref = O.gelu(pre)
    assert _err(y, ref)[1] < 3e-3
    # the GELU epilogue itself is bit-exact given the same pre-activation
    pre_gpu = ops.linear(x, w, b, bias_mode=ops.BIAS_FUSED)
    ref2 = O.gelu(pre_gpu.float())
    mism = (y.float() != ref2).float().mean().item()
    assert mism < 1e-3, mism


def test_linear_gate_residual(fluxlib):
    from diffusion_rs_b200 import ops
    B, T, N, K = 2, 384, 512, 256
    x = _rand((B, T, K), 10)
    w = _rand((N, K), 11, 1.0 / math.sqrt(K))
    b = _rand((N,), 12, 0.1)
    gate = _rand((B, N), 13)
    res = _rand((B, T, N), 14)
    out = res.clone()
    ops.linear(x, w, b, bias_mode=ops.BIAS_FUSED, gate=gate, rows_per_batch=T, res=out, out=out)
    v = O.linear(x.float(), w.float(), b.float(), fused_bias=True)
    ref = O.rb(res.float() + O.rb(gate.float()[:, None, :] * v))
    assert _err(out, ref)[1] < 3e-3
    # gate * x and + residual are two tensor ops in the reference (two roundings, model.rs:217-226): given the same
    # pre-activation the epilogue must be bit-exact, i.e. not contracted into one fused multiply-add
    pre_gpu = ops.linear(x, w, b, bias_mode=ops.BIAS_FUSED).float()
    ref2 = O.rb(res.float() + O.rb(gate.float()[:, None, :] * pre_gpu))
    mism = (out.float() != ref2).float().mean().item()
    assert mism < 1e-3, mism


def test_sdpa_rejects_what_the_flash_kernel_does_not_cover(fluxlib):
    """ops::sdpa(q, k, v, scale, softcapping) (ops.rs:247-262): head_dim != 128 or softcapping != 1.0 must fail loudly
    (status + message, nothing enqueued) so that a shim falls through to the stock path instead of computing garbage."""
    from diffusion_rs_b200 import lib as L
    from diffusion_rs_b200 import ops
    q = torch.randn(1, 2, 64, 128, device="cuda").bfloat16()
    with pytest.raises(L.Fluxb200Error, match="softcapping"):
        ops.sdpa(q, q, q, 0.1, softcapping=30.0)
    q64 = torch.randn(1, 2, 64, 64, device="cuda").bfloat16()
    with pytest.raises(L.Fluxb200Error, match="head_dim"):
        ops.sdpa(q64, q64, q64, 0.1)
    assert ops.sdpa(q, q, q, 0.1).shape == (1, 64, 256)


@pytest.mark.parametrize("B,H,L", [(1, 2, 256), (1, 3, 512), (2, 2, 384), (1, 2, 1000), (1, 24, 4608)])
def test_sdpa(fluxlib, B, H, L):
    from diffusion_rs_b200 import ops
    q = _rand((B, H, L, 128), 20)
    k = _rand((B, H, L, 128), 21)
    v = _rand((B, H, L, 128), 22)
    scale = 1.0 / math.sqrt(128)
    y = ops.sdpa(q, k, v, scale)
    torch.cuda.synchronize()
    ref = O.sdpa_f32(q.float(), k.float(), v.float(), scale)  # oracle semantics, evaluated on the GPU tensors
    ref = O.rb(ref).transpose(1, 2).reshape(B, L, H * 128)
    mx, rel = _err(y, ref)
    # reference error pattern: nn/tests/sdpa.rs (fused vs naive); P is rounded to bf16 before P.V here
    assert rel < 1e-2, (mx, rel)
    assert mx < 0.05, (mx, rel)


def test_sdpa_every_build_of_the_kernel(fluxlib):
    """Every run-time selectable build of the attention kernel ("attn_variant": polynomial shares, hand-off instalments,
    cooperative softmax warps, CTA pairs, ...) against the oracle on a ragged length (masked last kv block, 1000 is not a
    multiple of 128 / 256 / 512) with a late jump of the running max."""
    from diffusion_rs_b200 import lib as L
    from diffusion_rs_b200 import ops
    B, H, Lq = 1, 3, 1000
    q = _rand((B, H, Lq, 128), 26)
    k = _rand((B, H, Lq, 128), 27)
    v = _rand((B, H, Lq, 128), 28)
    k[:, :, 500:] *= 4.0
    scale = 1.0 / math.sqrt(128)
    ref = O.rb(O.sdpa_f32(q.float(), k.float(), v.float(), scale)).transpose(1, 2).reshape(B, Lq, H * 128)
    n = fluxlib.fluxb200_attn_variants()
    assert n >= 5
    try:
        for var in range(n):
            L.check(fluxlib.fluxb200_set_flag(b"attn_variant", var))
            y = ops.sdpa(q, k, v, scale)
            torch.cuda.synchronize()
            mx, rel = _err(y, ref)
            assert rel < 1e-2 and torch.isfinite(y.float()).all(), (var, mx, rel)
    finally:
        L.check(fluxlib.fluxb200_set_flag(b"attn_variant", 0))


@pytest.mark.parametrize("boost", [4.0, 60.0])
def test_sdpa_late_large_scores(fluxlib, boost):
    """Keys of the later kv blocks are much larger than the early ones, so the running max jumps after the first
    blocks: x4 exercises the deferred (one block late) rescale of the lazy-max softmax, x60 its overflow guard
    (scores exceed everything seen before by more than 2^64)."""
    from diffusion_rs_b200 import ops
    B, H, L = 1, 2, 1000
    q = _rand((B, H, L, 128), 23)
    k = _rand((B, H, L, 128), 24)
    v = _rand((B, H, L, 128), 25)
    k[:, :, 300:600] *= boost
    k[:, :, 900:] *= boost * 1.5
    scale = 1.0 / math.sqrt(128)
    y = ops.sdpa(q, k, v, scale)
    torch.cuda.synchronize()
    ref = O.sdpa_f32(q.float(), k.float(), v.float(), scale)
    ref = O.rb(ref).transpose(1, 2).reshape(B, L, H * 128)
    assert torch.isfinite(y.float()).all()
    mx, rel = _err(y, ref)
    assert rel < 1e-2, (mx, rel)


def test_layernorm_modulate(fluxlib):
    from diffusion_rs_b200 import ops
    B, T, D = 2, 300, 3072
    x = _rand((B, T, D), 30, 2.0)
    mod = _rand((B, 6 * D), 31, 0.5)
    shift, scale = mod[:, 0:D], mod[:, D:2 * D]
    y = ops.layernorm_modulate(x, shift, scale)
    n = O.layer_norm(x.float())
    ref = O.rb(O.rb(n * O.rb(scale.float()[:, None, :] + 1.0)) + shift.float()[:, None, :])
    mism = (y.float() != ref).float().mean().item()
    assert mism < 2e-3, mism
    assert _err(y, ref)[1] < 1e-3


def test_qknorm_rope(fluxlib):
    from diffusion_rs_b200 import ops
    B, T, H, Ltot, loff = 2, 200, 24, 328, 128
    D = H * 128
    qkv = _rand((B, T, 3 * D), 40)
    wq = _bf(1.0 + 0.02 * torch.randn(128, generator=torch.Generator().manual_seed(41))).cuda()
    wk = _bf(1.0 + 0.02 * torch.randn(128, generator=torch.Generator().manual_seed(42))).cuda()
    ang = torch.rand(Ltot, 64, generator=torch.Generator().manual_seed(43)) * 6.0
    pe_cos, pe_sin = _bf(torch.cos(ang)).cuda(), _bf(torch.sin(ang)).cuda()
    Q = torch.zeros(B, H, Ltot, 128, device="cuda", dtype=torch.bfloat16)
    K = torch.zeros_like(Q)
    V = torch.zeros_like(Q)
    ops.qknorm_rope(qkv, wq, wk, pe_cos, pe_sin, H, Ltot, loff, Q, K, V)
    f = qkv.float().reshape(B, T, 3, H, 128).permute(2, 0, 3, 1, 4)  # [3,B,H,T,128]

    def rope(x):
        c = pe_cos.float()[loff:loff + T][None, None]
        s = pe_sin.float()[loff:loff + T][None, None]
        x0, x1 = x[..., 0::2], x[..., 1::2]
        o0 = O.rb(O.rb(c * x0) + O.rb(-s * x1))
        o1 = O.rb(O.rb(s * x0) + O.rb(c * x1))
        return torch.stack([o0, o1], -1).reshape(x.shape)

    rq = rope(O.rms_norm_slow(f[0], wq.float()))
    rk = rope(O.rms_norm_slow(f[1], wk.float()))
    for got, ref in ((Q, rq), (K, rk), (V, f[2])):
        g = got[:, :, loff:loff + T].float()
        mism = (g != ref).float().mean().item()
        assert mism < 2e-3, mism
    assert Q[:, :, :loff].abs().max().item() == 0
