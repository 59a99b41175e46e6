"""ORACLE (test infrastructure, not product code) — CPU restatement of the VAE decode path.

Follows diffusion_rs_core/src/models/vaes/vae.rs (Decoder::forward :437-455, ResnetBlock :158-171, AttnBlock :96-110,
Upsample :224-228), nn/group_norm.rs:39-74, nn/conv.rs:212-230 and the pipeline tail flux/mod.rs:327-332.
NCHW float32 tensors carrying bf16-representable values (ops.REF) or plain f32 (ops.F32).
Model-level parity is unpinned by the reference (no VAE test / golden tensor); group_norm and conv2d are pinned by
the reference's known-answer vectors in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

from . import ops as O
from .ops import Mode, rb


@dataclass
class VaeConfig:  # autoencoder_kl.rs:15-32 (decoder-relevant fields; FLUX.1 values)
    latent_channels: int = 16
    out_channels: int = 3
    block_out_channels: tuple = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    mid_block_add_attention: bool = True
    scaling_factor: float = 0.3611
    shift_factor: float = 0.1159


def _gen(name):
    return torch.Generator(device="cpu").manual_seed(zlib.crc32(name.encode()))


def weight_specs(cfg: VaeConfig):
    """name -> (shape, kind) for every decoder tensor (vae.rs:371-433)."""
    s = {}
    ch = cfg.block_out_channels

    def conv(p, cin, cout, k):
        s[p + ".weight"] = ((cout, cin, k, k), "conv")
        s[p + ".bias"] = ((cout,), "bias")

    def norm(p, c):
        s[p + ".weight"] = ((c,), "norm")
        s[p + ".bias"] = ((c,), "bias")

    def resnet(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cin, cout, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cin, cout, 1)

    d = "decoder."
    block_in = ch[-1]
    conv(d + "conv_in", cfg.latent_channels, block_in, 3)
    resnet(d + "mid_block.resnets.0", block_in, block_in)
    if cfg.mid_block_add_attention:
        a = d + "mid_block.attentions.0."
        norm(a + "group_norm", block_in)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            s[a + n + ".weight"] = ((block_in, block_in), "linear")
            s[a + n + ".bias"] = ((block_in,), "bias")
    resnet(d + "mid_block.resnets.1", block_in, block_in)
    for lvl, block_out in enumerate(reversed(ch)):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"{d}up_blocks.{lvl}.resnets.{j}", block_in, block_out)
            block_in = block_out
        if lvl != 3:
            conv(f"{d}up_blocks.{lvl}.upsamplers.0.conv", block_in, block_in, 3)
    norm(d + "conv_norm_out", ch[0])
    conv(d + "conv_out", ch[0], cfg.out_channels, 3)
    return s


def make_weights(cfg: VaeConfig):
    w = {}
    for name, (shape, kind) in weight_specs(cfg).items():
        g = _gen(name)
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(*shape, generator=g) / math.sqrt(fan_in)
        elif kind == "linear":
            t = torch.randn(*shape, generator=g) / math.sqrt(shape[1])
        elif kind == "norm":
            t = 1.0 + 0.02 * torch.randn(*shape, generator=g)
        else:
            t = 0.02 * torch.randn(*shape, generator=g)
        w[name] = t.to(torch.bfloat16)
    return w


class VaeOracle:
    def __init__(self, cfg: VaeConfig, weights, mode: Mode = O.REF):
        self.cfg, self.mode = cfg, mode
        self.w = {k: v.to(torch.float32) for k, v in weights.items()}

    def conv(self, x, p, pad):  # Conv2d::forward nn/conv.rs:212-230: conv -> bf16, + bias -> bf16
        w = self.w[p + ".weight"]
        if w.dim() == 2:
            w = w[:, :, None, None]  # Linear used as a 1x1 conv kernel (vae.rs:46-82)
        y = rb(F.conv2d(x, w, None, padding=pad), self.mode)
        return rb(y + self.w[p + ".bias"][None, :, None, None], self.mode)

    def gn(self, x, p):
        return O.group_norm(x, self.w[p + ".weight"], self.w[p + ".bias"], self.cfg.norm_num_groups, 1e-6, self.mode)

    def resnet(self, x, p):  # vae.rs:158-171
        m = self.mode
        h = self.conv(O.silu(self.gn(x, p + ".norm1"), m), p + ".conv1", 1)
        h = self.conv(O.silu(self.gn(h, p + ".norm2"), m), p + ".conv2", 1)
        if (p + ".conv_shortcut.weight") in self.w:
            x = self.conv(x, p + ".conv_shortcut", 0)
        return rb(x + h, m)

    def attn(self, x, p):  # vae.rs:96-110 with scaled_dot_product_attention :28-33 in the model dtype
        m = self.mode
        h = self.gn(x, p + "group_norm")
        q = self.conv(h, p + "to_q", 0)
        k = self.conv(h, p + "to_k", 0)
        v = self.conv(h, p + "to_v", 0)
        b, c, hh, ww = q.shape
        q, k, v = (t.flatten(2).transpose(1, 2) for t in (q, k, v))  # [b, hw, c]
        att = rb(q @ k.transpose(1, 2), m)
        att = O.affine(att, 1.0 / math.sqrt(c), 0.0, m)  # `* scale_factor` on a bf16 tensor
        att = O.softmax_last_dim(att, m)
        o = rb(att @ v, m)
        o = o.transpose(1, 2).reshape(b, c, hh, ww)
        return rb(self.conv(o, p + "to_out.0", 0) + x, m)

    def decode(self, z):  # Decoder::forward vae.rs:437-455
        m = self.mode
        d = "decoder."
        h = self.conv(z, d + "conv_in", 1)
        h = self.resnet(h, d + "mid_block.resnets.0")
        if self.cfg.mid_block_add_attention:
            h = self.attn(h, d + "mid_block.attentions.0.")
        h = self.resnet(h, d + "mid_block.resnets.1")
        for lvl in range(4):
            for j in range(self.cfg.layers_per_block + 1):
                h = self.resnet(h, f"{d}up_blocks.{lvl}.resnets.{j}")
            if lvl != 3:
                h = F.interpolate(h, scale_factor=2, mode="nearest")  # upsample_nearest2d vae.rs:227
                h = self.conv(h, f"{d}up_blocks.{lvl}.upsamplers.0.conv", 1)
        h = O.silu(self.gn(h, d + "conv_norm_out"), m)
        return self.conv(h, d + "conv_out", 1)

    def decode_packed_u8(self, packed, height, width):
        """flux/mod.rs:327-332: unpack, z/scale + shift, decode, clamp, (x+1)*127.5, u8 (truncating cast)."""
        from .flux import unpack
        m = self.mode
        z = unpack(packed, height, width)
        z = O.affine(z, 1.0 / self.cfg.scaling_factor, 0.0, m)
        z = O.affine(z, 1.0, self.cfg.shift_factor, m)
        img = self.decode(z)
        img = img.clamp(-1.0, 1.0)
        img = O.affine(img, 1.0, 1.0, m)
        img = O.affine(img, 127.5, 0.0, m)
        return img.clamp(0, 255).to(torch.uint8)  # `v.to_f32() as u8`: truncation, saturating
