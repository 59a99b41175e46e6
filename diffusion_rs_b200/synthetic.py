"""Synthetic checkpoints (there is no network and no model weights on disk): random-init tensors of the exact
FLUX.1 / AutoencoderKL architecture, generated on the GPU, one deterministic generator per tensor name
(seed = crc32(name)), N(0, 1/sqrt(fan_in)) weights so activations stay O(1) through 57 blocks."""
from __future__ import annotations

import math
import zlib

import torch

HIDDEN, MLP = 3072, 12288


def _randn(name: str, shape, std: float, mean: float = 0.0, device="cuda"):
    g = torch.Generator(device=device).manual_seed(zlib.crc32(name.encode()))
    t = torch.randn(*shape, generator=g, device=device, dtype=torch.float32)
    return (t * std + mean).to(torch.bfloat16)


def flux_linear_shapes(cfg) -> dict[str, tuple[int, int]]:
    """(out, in) of every Linear `Flux::new` builds (models/flux/model.rs:722-787)."""
    s = {}
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


def iter_flux_tensors(cfg, device="cuda"):
    """Yield (name, bf16 tensor) for the whole transformer, one tensor at a time (23.8 GB in total for FLUX.1-dev)."""
    for name, (o, i) in flux_linear_shapes(cfg).items():
        yield name + ".weight", _randn(name + ".weight", (o, i), 1.0 / math.sqrt(i), device=device)
        yield name + ".bias", _randn(name + ".bias", (o,), 0.02, device=device)
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}.attn."
        for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            yield p + n + ".weight", _randn(p + n + ".weight", (128,), 0.02, 1.0, device=device)
    for i in range(cfg.num_single_layers):
        p = f"single_transformer_blocks.{i}.attn."
        for n in ("norm_q", "norm_k"):
            yield p + n + ".weight", _randn(p + n + ".weight", (128,), 0.02, 1.0, device=device)


def iter_vae_tensors(cfg, device="cuda"):
    """Yield (name, bf16 tensor) for the VAE decoder (models/vaes/vae.rs:371-433)."""
    ch = cfg.block_out_channels

    def conv(p, cin, cout, k):
        yield p + ".weight", _randn(p + ".weight", (cout, cin, k, k), 1.0 / math.sqrt(cin * k * k), device=device)
        yield p + ".bias", _randn(p + ".bias", (cout,), 0.02, device=device)

    def norm(p, c):
        yield p + ".weight", _randn(p + ".weight", (c,), 0.02, 1.0, device=device)
        yield p + ".bias", _randn(p + ".bias", (c,), 0.02, device=device)

    def resnet(p, cin, cout):
        yield from norm(p + ".norm1", cin)
        yield from conv(p + ".conv1", cin, cout, 3)
        yield from norm(p + ".norm2", cout)
        yield from conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            yield from conv(p + ".conv_shortcut", cin, cout, 1)

    d = "decoder."
    block_in = ch[-1]
    yield from conv(d + "conv_in", cfg.latent_channels, block_in, 3)
    yield from resnet(d + "mid_block.resnets.0", block_in, block_in)
    if cfg.mid_block_add_attention:
        a = d + "mid_block.attentions.0."
        yield from norm(a + "group_norm", block_in)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            yield a + n + ".weight", _randn(a + n + ".weight", (block_in, block_in), 1.0 / math.sqrt(block_in), device=device)
            yield a + n + ".bias", _randn(a + n + ".bias", (block_in,), 0.02, device=device)
    yield from resnet(d + "mid_block.resnets.1", block_in, block_in)
    for lvl, block_out in enumerate(reversed(ch)):
        for j in range(cfg.layers_per_block + 1):
            yield from resnet(f"{d}up_blocks.{lvl}.resnets.{j}", block_in, block_out)
            block_in = block_out
        if lvl != 3:
            yield from conv(f"{d}up_blocks.{lvl}.upsamplers.0.conv", block_in, block_in, 3)
    yield from norm(d + "conv_norm_out", ch[0])
    yield from conv(d + "conv_out", ch[0], cfg.out_channels, 3)


def iter_t5_tensors(cfg, device="cuda"):
    """Yield (name, bf16 tensor) for the T5 encoder (models/t5/mod.rs:645-657; 4.7e9 parameters for t5-v1_1-xxl)."""
    inner = cfg.num_heads * cfg.d_kv
    yield "shared.weight", _randn("shared.weight", (cfg.vocab_size, cfg.d_model), 1.0, device=device)
    for i in range(cfg.num_layers):
        p = f"encoder.block.{i}.layer."
        for n, shp in (("q", (inner, cfg.d_model)), ("k", (inner, cfg.d_model)), ("v", (inner, cfg.d_model)),
                       ("o", (cfg.d_model, inner))):
            name = p + f"0.SelfAttention.{n}.weight"
            yield name, _randn(name, shp, 1.0 / math.sqrt(shp[1]), device=device)
        if i == 0:
            name = p + "0.SelfAttention.relative_attention_bias.weight"
            yield name, _randn(name, (cfg.relative_attention_num_buckets, cfg.num_heads), 0.5, device=device)
        yield p + "0.layer_norm.weight", _randn(p + "0.layer_norm.weight", (cfg.d_model,), 0.02, 1.0, device=device)
        for n, shp in (("wi_0", (cfg.d_ff, cfg.d_model)), ("wi_1", (cfg.d_ff, cfg.d_model)), ("wo", (cfg.d_model, cfg.d_ff))):
            name = p + f"1.DenseReluDense.{n}.weight"
            yield name, _randn(name, shp, 1.0 / math.sqrt(shp[1]), device=device)
        yield p + "1.layer_norm.weight", _randn(p + "1.layer_norm.weight", (cfg.d_model,), 0.02, 1.0, device=device)
    yield "encoder.final_layer_norm.weight", _randn("encoder.final_layer_norm.weight", (cfg.d_model,), 0.02, 1.0,
                                                    device=device)


def iter_clip_tensors(cfg, device="cuda"):
    """Yield (name, bf16 tensor) for the CLIP text tower, names relative to `text_model.` (models/clip/text.rs:253-265)."""
    D, I = cfg.projection_dim, cfg.intermediate_size
    yield "embeddings.token_embedding.weight", _randn("clip.tok", (cfg.vocab_size, D), 0.5, device=device)
    yield "embeddings.position_embedding.weight", _randn("clip.pos", (cfg.max_position_embeddings, D), 0.1, device=device)

    def lin(name, o, i):
        yield name + ".weight", _randn("clip." + name + ".w", (o, i), 1.0 / math.sqrt(i), device=device)
        yield name + ".bias", _randn("clip." + name + ".b", (o,), 0.02, device=device)

    def ln(name):
        yield name + ".weight", _randn("clip." + name + ".w", (D,), 0.02, 1.0, device=device)
        yield name + ".bias", _randn("clip." + name + ".b", (D,), 0.02, device=device)

    for i in range(cfg.num_hidden_layers):
        p = f"encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            yield from lin(p + "self_attn." + n, D, D)
        yield from ln(p + "layer_norm1")
        yield from ln(p + "layer_norm2")
        yield from lin(p + "mlp.fc1", I, D)
        yield from lin(p + "mlp.fc2", D, I)
    yield from ln("final_layer_norm")
