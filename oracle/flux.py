"""ORACLE (test infrastructure, not product code) — CPU restatement of the FLUX MMDiT step.

Follows diffusion_rs_core/src/models/flux/model.rs line by line (citations on each function).  Tensors are float32
carrying bf16-representable values in `ops.REF` mode (every reference rounding point mirrored) or plain f32 in
`ops.F32` mode.  The reference ships no FLUX test / golden tensor and cannot be run here (Rust); the op-level pieces
are pinned by the reference's own known-answer vectors (tests/test_oracle_golden.py) and the model-level graph is
pinned, in f32 mode, against an independent implementation of the same network — the Black Forest Labs code as
vendored by torchtitan — under the diffusers -> BFL weight-name map (tests/test_flux_oracle_pin.py, rel. error < 2e-5).

Weight names are the diffusers names the reference's VarBuilder paths produce (model.rs:722-787).
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field

import torch

from . import ops as O
from .ops import Mode, rb

HIDDEN = 3072  # HIDDEN_SIZE model.rs:17
MLP = 4 * HIDDEN  # MLP_RATIO model.rs:16
AXES_DIM = (16, 56, 56)  # model.rs:18
THETA = 10000  # model.rs:19


@dataclass
class FluxConfig:  # model.rs:21-31
    in_channels: int = 64
    pooled_projection_dim: int = 768
    joint_attention_dim: int = 4096
    num_attention_heads: int = 24
    num_layers: int = 19
    num_single_layers: int = 38
    guidance_embeds: bool = True


# ---------------------------------------------------------------------------------------------------------
# synthetic weights: per-tensor generator seeded by crc32(name), N(0, 1/sqrt(fan_in)) -> bf16
# ---------------------------------------------------------------------------------------------------------
def _gen(name: str) -> torch.Generator:
    return torch.Generator(device="cpu").manual_seed(zlib.crc32(name.encode()))


def linear_shapes(cfg: FluxConfig) -> dict[str, tuple[int, int]]:
    """name prefix -> (out, in) of every Linear of the transformer (all have a bias)."""
    s: dict[str, tuple[int, int]] = {}
    s["x_embedder"] = (HIDDEN, cfg.in_channels)
    s["context_embedder"] = (HIDDEN, cfg.joint_attention_dim)
    s["time_text_embed.timestep_embedder.linear_1"] = (HIDDEN, 256)
    s["time_text_embed.timestep_embedder.linear_2"] = (HIDDEN, HIDDEN)
    s["time_text_embed.text_embedder.linear_1"] = (HIDDEN, cfg.pooled_projection_dim)
    s["time_text_embed.text_embedder.linear_2"] = (HIDDEN, HIDDEN)
    if cfg.guidance_embeds:
        s["time_text_embed.guidance_embedder.linear_1"] = (HIDDEN, 256)
        s["time_text_embed.guidance_embedder.linear_2"] = (HIDDEN, HIDDEN)
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        s[p + "norm1.linear"] = (6 * HIDDEN, HIDDEN)
        s[p + "norm1_context.linear"] = (6 * HIDDEN, HIDDEN)
        for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
            s[p + "attn." + n] = (HIDDEN, HIDDEN)
        s[p + "ff.net.0.proj"] = (MLP, HIDDEN)
        s[p + "ff.net.2"] = (HIDDEN, MLP)
        s[p + "ff_context.net.0.proj"] = (MLP, HIDDEN)
        s[p + "ff_context.net.2"] = (HIDDEN, MLP)
    for i in range(cfg.num_single_layers):
        p = f"single_transformer_blocks.{i}."
        s[p + "norm.linear"] = (3 * HIDDEN, HIDDEN)
        for n in ("to_q", "to_k", "to_v"):
            s[p + "attn." + n] = (HIDDEN, HIDDEN)
        s[p + "proj_mlp"] = (MLP, HIDDEN)
        s[p + "proj_out"] = (HIDDEN, HIDDEN + MLP)
    s["norm_out.linear"] = (2 * HIDDEN, HIDDEN)
    s["proj_out"] = (cfg.in_channels, HIDDEN)
    return s


def norm_weight_names(cfg: FluxConfig) -> list[str]:
    names = []
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}.attn."
        names += [p + "norm_q.weight", p + "norm_k.weight", p + "norm_added_q.weight", p + "norm_added_k.weight"]
    for i in range(cfg.num_single_layers):
        p = f"single_transformer_blocks.{i}.attn."
        names += [p + "norm_q.weight", p + "norm_k.weight"]
    return names


def make_tensor(name: str, shape, kind: str) -> torch.Tensor:
    """bf16 tensor for `name`; kind in {'weight','bias','norm'}."""
    g = _gen(name)
    if kind == "weight":
        t = torch.randn(*shape, generator=g) * (1.0 / math.sqrt(shape[-1]))
    elif kind == "bias":
        t = torch.randn(*shape, generator=g) * 0.02
    else:
        t = 1.0 + torch.randn(*shape, generator=g) * 0.02
    return t.to(torch.bfloat16)


def make_weights(cfg: FluxConfig) -> dict[str, torch.Tensor]:
    """All transformer tensors as bf16 CPU tensors (only for small configs; the full model is 23.8 GB)."""
    w = {}
    for name, (o, i) in linear_shapes(cfg).items():
        w[name + ".weight"] = make_tensor(name + ".weight", (o, i), "weight")
        w[name + ".bias"] = make_tensor(name + ".bias", (o,), "bias")
    for name in norm_weight_names(cfg):
        w[name] = make_tensor(name, (128,), "norm")
    return w


# ---------------------------------------------------------------------------------------------------------
# model pieces
# ---------------------------------------------------------------------------------------------------------
def rope_table(pos: torch.Tensor, dim: int, theta: int, mode: Mode):
    """rope() model.rs:65-84.  pos: [L] values in the ids dtype (bf16 in REF mode).  Returns (cos, sin) [L, dim/2].
    inv_freq is computed in f64, cast to f32, then to the ids dtype; pos*inv_freq, cos and sin are dtype ops."""
    idx = torch.arange(0, dim, 2, dtype=torch.float64)
    inv_freq = (1.0 / torch.pow(torch.tensor(float(theta), dtype=torch.float64), idx / dim)).to(torch.float32)
    inv_freq = rb(inv_freq, mode)
    freqs = rb(pos.to(torch.float32)[:, None] * inv_freq[None, :], mode)
    return rb(torch.cos(freqs), mode), rb(torch.sin(freqs), mode)


def embed_nd(ids: torch.Tensor, mode: Mode):
    """EmbedNd::forward model.rs:142-157: per-axis rope tables concatenated on the frequency dim.
    ids: [L, 3] -> (cos, sin) each [L, 64].  The 2x2 matrix is [[cos, -sin], [sin, cos]]."""
    cs, ss = [], []
    for a, d in enumerate(AXES_DIM):
        c, s = rope_table(rb(ids[:, a].to(torch.float32), mode), d, THETA, mode)
        cs.append(c)
        ss.append(s)
    return torch.cat(cs, 1), torch.cat(ss, 1)


def make_ids(h2: int, w2: int, l_txt: int):
    """State::new pipelines/flux/sampling.rs:32-49: txt ids are zeros, img ids are (0, row, col)."""
    img_ids = torch.zeros(h2, w2, 3)
    img_ids[..., 1] = torch.arange(h2)[:, None]
    img_ids[..., 2] = torch.arange(w2)[None, :]
    return torch.cat([torch.zeros(l_txt, 3), img_ids.reshape(h2 * w2, 3)], 0)


def apply_rope(x, cos, sin, mode: Mode):
    """apply_rope model.rs:86-95 on x [B,H,L,128]: interleaved pairs; 2 products and 1 sum, each rounded."""
    x0, x1 = x[..., 0::2], x[..., 1::2]
    c, s = cos[None, None], sin[None, None]
    o0 = rb(rb(c * x0, mode) + rb(-s * x1, mode), mode)
    o1 = rb(rb(s * x0, mode) + rb(c * x1, mode), mode)
    return torch.stack([o0, o1], -1).reshape(x.shape)


def timestep_embedding(t: torch.Tensor, dim: int, mode: Mode):
    """model.rs:104-122 — f32 throughout, cast to the model dtype at the end."""
    half = dim // 2
    t = t.to(torch.float32) * 1000.0
    coef = torch.tensor(-math.log(10000.0) / half, dtype=torch.float32)
    freqs = torch.exp(torch.arange(half, dtype=torch.float32) * coef)
    args = t[:, None] * freqs[None]
    return rb(torch.cat([torch.cos(args), torch.sin(args)], -1), mode)


class _WidenOnAccess:
    def __init__(self, weights, device):
        self._w, self._device = weights, device

    def __getitem__(self, k):
        return self._w[k].to(device=self._device, dtype=torch.float32)

    def __contains__(self, k):
        return k in self._w


class FluxOracle:
    def __init__(self, cfg: FluxConfig, weights: dict[str, torch.Tensor], mode: Mode = O.REF, device="cpu"):
        self.cfg, self.mode, self.device = cfg, mode, device
        # weights stay in their storage dtype (bf16) and are widened on access, so a full-depth model (23.8 GB bf16)
        # does not need a second 48 GB f32 copy
        self.w = _WidenOnAccess(weights, device)

    # Linear on a rank-3 activation with bias: cuBLASLt fused bias (unquantized/mod.rs:52-66)
    def lin3(self, x, name):
        return O.linear(x, self.w[name + ".weight"], self.w[name + ".bias"], fused_bias=True, mode=self.mode)

    # Linear on a rank-2 activation: matmul then separate add (unquantized/mod.rs:67)
    def lin2(self, x, name):
        return O.linear(x, self.w[name + ".weight"], self.w[name + ".bias"], fused_bias=False, mode=self.mode)

    def mlp_embedder(self, x, prefix):  # MlpEmbedder::forward model.rs:178-183
        return self.lin2(O.silu(self.lin2(x, prefix + ".linear_1"), self.mode), prefix + ".linear_2")

    def modulation(self, vec, name, n):  # Modulation1/2::forward model.rs:244-299
        ys = self.lin2(O.silu(vec, self.mode), name)
        return [c[:, None, :] for c in ys.chunk(n, -1)]

    def scale_shift(self, x, shift, scale):  # ModulationOut::scale_shift model.rs:218-221
        m = self.mode
        return rb(rb(x * rb(scale + 1.0, m), m) + shift, m)

    def heads(self, x):  # reshape + transpose model.rs:415-423
        b, l, _ = x.shape
        return x.reshape(b, l, self.cfg.num_attention_heads, -1).transpose(1, 2)

    def attention(self, q, k, v, pe):  # attention() model.rs:97-102 + scaled_dot_product_attention :40-51
        cos, sin = pe
        q = apply_rope(q, cos, sin, self.mode)
        k = apply_rope(k, cos, sin, self.mode)
        o = O.sdpa_f32(q, k, v, 1.0 / math.sqrt(q.shape[-1]))
        o = rb(o, self.mode)
        b, h, l, d = o.shape
        return o.transpose(1, 2).reshape(b, l, h * d)

    def qkv(self, x, p, names, norm_names):  # SelfAttention::qkv model.rs:399-427
        q = self.heads(self.lin3(x, p + names[0]))
        k = self.heads(self.lin3(x, p + names[1]))
        v = self.heads(self.lin3(x, p + names[2]))
        q = O.rms_norm_slow(q, self.w[p + norm_names[0]], 1e-6, self.mode)
        k = O.rms_norm_slow(k, self.w[p + norm_names[1]], 1e-6, self.mode)
        return q, k, v

    def mlp(self, x, p):  # Mlp::forward model.rs:458-464
        return self.lin3(O.gelu(self.lin3(x, p + ".0.proj"), self.mode), p + ".2")

    def double_block(self, i, img, txt, vec, pe):  # DoubleStreamBlock::forward model.rs:523-565
        m = self.mode
        p = f"transformer_blocks.{i}."
        i_sh1, i_sc1, i_g1, i_sh2, i_sc2, i_g2 = self.modulation(vec, p + "norm1.linear", 6)
        t_sh1, t_sc1, t_g1, t_sh2, t_sc2, t_g2 = self.modulation(vec, p + "norm1_context.linear", 6)
        img_mod = self.scale_shift(O.layer_norm(img, mode=m), i_sh1, i_sc1)
        iq, ik, iv = self.qkv(img_mod, p + "attn.", ("to_q", "to_k", "to_v"), ("norm_q.weight", "norm_k.weight"))
        txt_mod = self.scale_shift(O.layer_norm(txt, mode=m), t_sh1, t_sc1)
        tq, tk, tv = self.qkv(txt_mod, p + "attn.", ("add_q_proj", "add_k_proj", "add_v_proj"),
                              ("norm_added_q.weight", "norm_added_k.weight"))
        q, k, v = torch.cat([tq, iq], 2), torch.cat([tk, ik], 2), torch.cat([tv, iv], 2)
        attn = self.attention(q, k, v, pe)
        lt = txt.shape[1]
        txt_attn, img_attn = attn[:, :lt], attn[:, lt:]
        img = rb(img + rb(i_g1 * self.lin3(img_attn, p + "attn.to_out.0"), m), m)
        h = self.mlp(self.scale_shift(O.layer_norm(img, mode=m), i_sh2, i_sc2), p + "ff.net")
        img = rb(img + rb(i_g2 * h, m), m)
        txt = rb(txt + rb(t_g1 * self.lin3(txt_attn, p + "attn.to_add_out"), m), m)
        h = self.mlp(self.scale_shift(O.layer_norm(txt, mode=m), t_sh2, t_sc2), p + "ff_context.net")
        txt = rb(txt + rb(t_g2 * h, m), m)
        return img, txt

    def single_block(self, i, x, vec, pe):  # SingleStreamBlock::forward model.rs:638-662
        m = self.mode
        p = f"single_transformer_blocks.{i}."
        sh, sc, g = self.modulation(vec, p + "norm.linear", 3)
        x_mod = self.scale_shift(O.layer_norm(x, mode=m), sh, sc)
        q = self.heads(self.lin3(x_mod, p + "attn.to_q"))
        k = self.heads(self.lin3(x_mod, p + "attn.to_k"))
        v = self.heads(self.lin3(x_mod, p + "attn.to_v"))
        q = O.rms_norm_slow(q, self.w[p + "attn.norm_q.weight"], 1e-6, m)
        k = O.rms_norm_slow(k, self.w[p + "attn.norm_k.weight"], 1e-6, m)
        mlp = self.lin3(x_mod, p + "proj_mlp")
        attn = self.attention(q, k, v, pe)
        out = self.lin3(torch.cat([attn, O.gelu(mlp, m)], 2), p + "proj_out")
        return rb(x + rb(g * out, m), m)

    def last_layer(self, x, vec):  # LastLayer::forward model.rs:694-705
        m = self.mode
        scale, shift = self.lin2(O.silu(vec, m), "norm_out.linear").chunk(2, 1)
        x = rb(rb(O.layer_norm(x, mode=m) * rb(scale[:, None] + 1.0, m), m) + shift[:, None], m)
        return self.lin3(x, "proj_out")

    def vec(self, timesteps, y, guidance):  # model.rs:813-820
        m = self.mode
        v = self.mlp_embedder(timestep_embedding(timesteps, 256, m), "time_text_embed.timestep_embedder")
        if self.cfg.guidance_embeds and guidance is not None:
            v = rb(v + self.mlp_embedder(timestep_embedding(guidance, 256, m), "time_text_embed.guidance_embedder"), m)
        return rb(v + self.mlp_embedder(y, "time_text_embed.text_embedder"), m)

    def forward(self, img, ids, txt, timesteps, y, guidance, *, taps=None):
        """Flux::forward model.rs:790-833.  img [B,L_img,64], txt [B,L_txt,4096], ids [L,3] (txt first),
        timesteps/guidance f32 [B], y [B,768].  `taps` (dict) collects intermediates for parity tests."""
        pe = embed_nd(ids.to(self.device), self.mode)
        txt = self.lin3(txt, "context_embedder")
        img = self.lin3(img, "x_embedder")
        vec = self.vec(timesteps, y, guidance)
        if taps is not None:
            taps.update(pe_cos=pe[0], pe_sin=pe[1], vec=vec, img_in=img, txt_in=txt)
        for i in range(self.cfg.num_layers):
            img, txt = self.double_block(i, img, txt, vec, pe)
            if taps is not None:
                taps[f"double{i}.img"], taps[f"double{i}.txt"] = img, txt
        x = torch.cat([txt, img], 1)
        for i in range(self.cfg.num_single_layers):
            x = self.single_block(i, x, vec, pe)
            if taps is not None:
                taps[f"single{i}"] = x
        x = x[:, txt.shape[1]:]
        return self.last_layer(x, vec)


# ---------------------------------------------------------------------------------------------------------
# sampler / scheduler  (pipelines/scheduler.rs:22-51, flux/sampling.rs:5-80, sampling.rs:25-48)
# ---------------------------------------------------------------------------------------------------------
def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=4096, base_shift=0.5, max_shift=1.15):
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def get_timesteps(num_steps, mu=None, shift=1.0, dynamic=True):
    sig = [v / num_steps for v in range(num_steps, -1, -1)]
    out = []
    for s in sig:
        if dynamic:
            e = math.exp(mu)
            out.append(e / (e + (1.0 / s - 1.0) ** 1.0) if s > 0 else 0.0)
        else:
            out.append(shift * s / (1.0 + (shift - 1.0) * s))
    return out


def patchify(lat):  # State::new flux/sampling.rs:29-31: [B,C,H,W] -> [B,(H/2)(W/2),C*4]
    b, c, h, w = lat.shape
    return lat.reshape(b, c, h // 2, 2, w // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(b, (h // 2) * (w // 2), c * 4)


def unpack(x, height, width):  # flux/sampling.rs:61-68
    b, _, cpp = x.shape
    h, w = (height + 15) // 16, (width + 15) // 16
    return x.reshape(b, h, w, cpp // 4, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(b, cpp // 4, h * 2, w * 2)


def euler_step(img, pred, t_curr, t_prev, mode: Mode):  # sampling.rs:43: img + pred * (t_prev - t_curr)
    return rb(img + rb(pred * O.bf16_scalar(t_prev - t_curr, mode), mode), mode)
