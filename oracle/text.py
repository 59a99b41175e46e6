"""ORACLE (test infrastructure, not product code) — CPU restatement of the two text encoders FLUX conditions on.

T5 encoder : diffusion_rs_core/src/models/t5/mod.rs (T5LayerNorm :96-122, gated FF :160-198, attention with the
             bidirectional relative-position buckets :283-383, block/stack :512-640, T5EncoderModel :642-661).
CLIP text  : diffusion_rs_core/src/models/clip/text.rs (embeddings :37-72, attention :74-152, quick_gelu MLP :8-20,
             :154-180, encoder layer :182-221, causal mask + EOS pooling :268-317).
They are called once per prompt by FluxPipeline::forward (pipelines/flux/mod.rs:236-262).

Tensors are float32 carrying bf16-representable values in `ops.REF` mode (every reference rounding point mirrored)
or plain f32 in `ops.F32` mode.  Pinning: the reference ships no test / golden tensor for these models (parity
unpinned at model level); tests/test_text_oracle.py pins the F32 graph against the independent HuggingFace
`transformers` implementations (T5EncoderModel, CLIPTextModel) on random small configs, which is the code the
reference itself cites as its source (t5/mod.rs:4).
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass

import torch

from . import ops as O
from .ops import Mode, rb


def _gen(name: str) -> torch.Generator:
    return torch.Generator(device="cpu").manual_seed(zlib.crc32(name.encode()))


def _t(name, shape, std=None, mean=0.0):
    std = (1.0 / math.sqrt(shape[-1])) if std is None else std
    return (mean + torch.randn(*shape, generator=_gen(name)) * std).to(torch.bfloat16)


# ---------------------------------------------------------------------------------------------------------
# T5 encoder
# ---------------------------------------------------------------------------------------------------------
@dataclass
class T5Config:  # t5/mod.rs:75-93 (google/t5-v1_1-xxl values as defaults)
    vocab_size: int = 32128
    d_model: int = 4096
    d_kv: int = 64
    d_ff: int = 10240
    num_layers: int = 24
    num_heads: int = 64
    relative_attention_num_buckets: int = 32
    relative_attention_max_distance: int = 128
    layer_norm_epsilon: float = 1e-6


def t5_make_weights(cfg: T5Config) -> dict[str, torch.Tensor]:
    """bf16 tensors under the names T5EncoderModel::new resolves (mod.rs:645-657)."""
    w = {"shared.weight": _t("shared.weight", (cfg.vocab_size, cfg.d_model), std=1.0)}
    inner = cfg.num_heads * cfg.d_kv
    for i in range(cfg.num_layers):
        p = f"encoder.block.{i}.layer."
        for n, shp in (("q", (inner, cfg.d_model)), ("k", (inner, cfg.d_model)), ("v", (inner, cfg.d_model)),
                       ("o", (cfg.d_model, inner))):
            w[p + f"0.SelfAttention.{n}.weight"] = _t(p + f"0.SelfAttention.{n}.weight", shp)
        if i == 0:
            name = p + "0.SelfAttention.relative_attention_bias.weight"
            w[name] = _t(name, (cfg.relative_attention_num_buckets, cfg.num_heads), std=0.5)
        w[p + "0.layer_norm.weight"] = _t(p + "0.layer_norm.weight", (cfg.d_model,), std=0.02, mean=1.0)
        w[p + "1.DenseReluDense.wi_0.weight"] = _t(p + "1.DenseReluDense.wi_0.weight", (cfg.d_ff, cfg.d_model))
        w[p + "1.DenseReluDense.wi_1.weight"] = _t(p + "1.DenseReluDense.wi_1.weight", (cfg.d_ff, cfg.d_model))
        w[p + "1.DenseReluDense.wo.weight"] = _t(p + "1.DenseReluDense.wo.weight", (cfg.d_model, cfg.d_ff))
        w[p + "1.layer_norm.weight"] = _t(p + "1.layer_norm.weight", (cfg.d_model,), std=0.02, mean=1.0)
    w["encoder.final_layer_norm.weight"] = _t("encoder.final_layer_norm.weight", (cfg.d_model,), std=0.02, mean=1.0)
    return w


def t5_relative_buckets(q_len: int, kv_len: int, num_buckets_total: int, max_distance: int) -> torch.Tensor:
    """The bidirectional bucket function exactly as written at mod.rs:334-372 (f32 log, truncating casts)."""
    num_buckets = num_buckets_total // 2
    max_exact = num_buckets // 2
    out = torch.zeros(q_len, kv_len, dtype=torch.int64)
    f32 = torch.float32
    for i in range(q_len):
        for j in range(kv_len):
            if i < j:
                d = j - i
                if d < max_exact:
                    out[i, j] = d + num_buckets
                else:
                    b = (torch.log(torch.tensor(d / max_exact, dtype=f32)) /
                         torch.log(torch.tensor(max_distance / max_exact, dtype=f32))) * (num_buckets - max_exact)
                    out[i, j] = min(max_exact + num_buckets + int(b.item()), num_buckets_total - 1)
            else:
                d = i - j
                if d < max_exact:
                    out[i, j] = d
                else:
                    b = (torch.log(torch.tensor(d / max_exact, dtype=f32)) /
                         torch.log(torch.tensor(max_distance / max_exact, dtype=f32))) * (num_buckets - max_exact)
                    out[i, j] = min(max_exact + int(b.item()), num_buckets - 1)
    return out


def t5_layer_norm(x, weight, eps, mode: Mode):
    """T5LayerNorm::forward mod.rs:111-121: f32 x / sqrt(mean(x^2) + eps) -> dtype, then * weight (dtype op)."""
    var = (x * x).mean(-1, keepdim=True)
    y = rb(x / torch.sqrt(var + eps), mode)
    return rb(y * weight, mode)


class T5Oracle:
    def __init__(self, cfg: T5Config, weights: dict[str, torch.Tensor], mode: Mode = O.REF):
        self.cfg, self.mode = cfg, mode
        self.w = {k: v.to(torch.float32) for k, v in weights.items()}

    def _lin(self, x, name):  # linear_no_bias -> UnquantLinear: matmul rounded to the activation dtype
        return rb(x @ self.w[name].t(), self.mode)

    def position_bias(self, L: int) -> torch.Tensor:
        """[1, H, L, L] = relative_attention_bias(buckets).permute(2,0,1).unsqueeze(0) (mod.rs:374-379)."""
        c = self.cfg
        b = t5_relative_buckets(L, L, c.relative_attention_num_buckets, c.relative_attention_max_distance)
        emb = self.w["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
        return emb[b].permute(2, 0, 1).unsqueeze(0)

    def attention(self, x, i, pos_bias):  # T5Attention::forward mod.rs:283-395 (self-attention, no mask, no scaling)
        c, m = self.cfg, self.mode
        B, L, _ = x.shape
        p = f"encoder.block.{i}.layer.0.SelfAttention."
        q = self._lin(x, p + "q.weight").reshape(B, L, c.num_heads, c.d_kv).transpose(1, 2)
        k = self._lin(x, p + "k.weight").reshape(B, L, c.num_heads, c.d_kv).transpose(1, 2)
        v = self._lin(x, p + "v.weight").reshape(B, L, c.num_heads, c.d_kv).transpose(1, 2)
        scores = rb(q @ k.transpose(-1, -2), m)          # bf16 matmul
        scores = rb(scores + pos_bias, m)                # broadcast_add
        att = O.softmax_last_dim(scores, m)              # softmax_last_dim on the bf16 tensor
        out = rb(att @ v, m)                             # bf16 matmul
        out = out.transpose(1, 2).reshape(B, L, c.num_heads * c.d_kv)
        return self._lin(out, p + "o.weight")

    def ff(self, x, i):  # T5DenseGatedActDense mod.rs:189-197: NewGelu(wi_0 x) * wi_1 x -> wo
        p = f"encoder.block.{i}.layer.1.DenseReluDense."
        g = O.gelu(self._lin(x, p + "wi_0.weight"), self.mode)
        h = rb(g * self._lin(x, p + "wi_1.weight"), self.mode)
        return self._lin(h, p + "wo.weight")

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        """ids int64 [B, L] -> last hidden state [B, L, d_model] (T5Stack::forward mod.rs:612-627)."""
        c, m = self.cfg, self.mode
        x = self.w["shared.weight"][ids]
        pos_bias = self.position_bias(ids.shape[1])
        for i in range(c.num_layers):
            p = f"encoder.block.{i}.layer."
            n = t5_layer_norm(x, self.w[p + "0.layer_norm.weight"], c.layer_norm_epsilon, m)
            x = rb(x + self.attention(n, i, pos_bias), m)
            n = t5_layer_norm(x, self.w[p + "1.layer_norm.weight"], c.layer_norm_epsilon, m)
            x = rb(x + self.ff(n, i), m)
        return t5_layer_norm(x, self.w["encoder.final_layer_norm.weight"], c.layer_norm_epsilon, m)


# ---------------------------------------------------------------------------------------------------------
# CLIP text transformer
# ---------------------------------------------------------------------------------------------------------
@dataclass
class ClipConfig:  # clip/text.rs:22-31 (openai/clip-vit-large-patch14 text tower as defaults)
    vocab_size: int = 49408
    projection_dim: int = 768  # the reference uses projection_dim as the hidden width
    intermediate_size: int = 3072
    max_position_embeddings: int = 77
    num_hidden_layers: int = 12
    num_attention_heads: int = 12


def clip_make_weights(cfg: ClipConfig) -> dict[str, torch.Tensor]:
    """Names relative to `text_model.` (ClipTextTransformer::new text.rs:253-265)."""
    D, I = cfg.projection_dim, cfg.intermediate_size
    w = {"embeddings.token_embedding.weight": _t("clip.tok", (cfg.vocab_size, D), std=0.5),
         "embeddings.position_embedding.weight": _t("clip.pos", (cfg.max_position_embeddings, D), std=0.1)}
    for i in range(cfg.num_hidden_layers):
        p = f"encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            w[p + f"self_attn.{n}.weight"] = _t("clip." + p + n + ".w", (D, D))
            w[p + f"self_attn.{n}.bias"] = _t("clip." + p + n + ".b", (D,), std=0.02)
        for n in ("layer_norm1", "layer_norm2"):
            w[p + n + ".weight"] = _t("clip." + p + n + ".w", (D,), std=0.02, mean=1.0)
            w[p + n + ".bias"] = _t("clip." + p + n + ".b", (D,), std=0.02)
        w[p + "mlp.fc1.weight"] = _t("clip." + p + "fc1.w", (I, D))
        w[p + "mlp.fc1.bias"] = _t("clip." + p + "fc1.b", (I,), std=0.02)
        w[p + "mlp.fc2.weight"] = _t("clip." + p + "fc2.w", (D, I))
        w[p + "mlp.fc2.bias"] = _t("clip." + p + "fc2.b", (D,), std=0.02)
    w["final_layer_norm.weight"] = _t("clip.final_ln.w", (D,), std=0.02, mean=1.0)
    w["final_layer_norm.bias"] = _t("clip.final_ln.b", (D,), std=0.02)
    return w


def sigmoid(v, mode: Mode):
    """recip(1 + exp(-v)), every step rounded on bf16 tensors (nn/ops.rs Sigmoid bf16; unary.cu:64-66)."""
    e = rb(torch.exp(-v), mode)
    return rb(1.0 / rb(1.0 + e, mode), mode)


def quick_gelu(v, mode: Mode):
    """xs * sigmoid(xs * 1.702) (text.rs:15-19); `xs * 1.702` is Tensor::affine(1.702, 0)."""
    return rb(v * sigmoid(O.affine(v, 1.702, 0.0, mode), mode), mode)


class ClipOracle:
    def __init__(self, cfg: ClipConfig, weights: dict[str, torch.Tensor], mode: Mode = O.REF):
        self.cfg, self.mode = cfg, mode
        self.w = {k: v.to(torch.float32) for k, v in weights.items()}

    def _lin(self, x, name):  # nn::Linear::forward (nn/linear.rs): matmul rounded, then broadcast_add(bias) rounded
        return O.linear(x, self.w[name + ".weight"], self.w[name + ".bias"], fused_bias=False, mode=self.mode)

    def _ln(self, x, name):  # nn::LayerNorm fast path with affine weight/bias, eps 1e-5 (text.rs:193-198)
        return O.layer_norm(x, self.w[name + ".weight"], self.w[name + ".bias"], eps=1e-5, mode=self.mode)

    def attention(self, x, i, mask):  # ClipAttention::forward text.rs:113-151
        c, m = self.cfg, self.mode
        B, L, D = x.shape
        H = c.num_attention_heads
        hd = D // H
        p = f"encoder.layers.{i}.self_attn."
        q = O.affine(self._lin(x, p + "q_proj"), hd ** -0.5, 0.0, m)  # (q_proj(x) * scale) in the model dtype
        q = q.reshape(B, L, H, hd).transpose(1, 2)
        k = self._lin(x, p + "k_proj").reshape(B, L, H, hd).transpose(1, 2)
        v = self._lin(x, p + "v_proj").reshape(B, L, H, hd).transpose(1, 2)
        att = q @ k.transpose(-1, -2) + mask  # f32 from here on (to_dtype(F32) text.rs:121-130)
        att = torch.softmax(att, dim=-1)
        out = rb(att @ v, m)                   # .to_dtype(in_dtype)
        out = out.transpose(1, 2).reshape(B, L, D)
        return self._lin(out, p + "out_proj")

    def hidden(self, ids: torch.Tensor) -> torch.Tensor:
        """forward_with_mask(ids, usize::MAX) text.rs:291-300: [B, L, D] after the final LayerNorm."""
        c, m = self.cfg, self.mode
        B, L = ids.shape
        x = rb(self.w["embeddings.token_embedding.weight"][ids] +
               self.w["embeddings.position_embedding.weight"][:L][None], m)
        mask = torch.zeros(L, L)
        mask[torch.triu(torch.ones(L, L, dtype=torch.bool), diagonal=1)] = torch.finfo(torch.float32).min
        for i in range(c.num_hidden_layers):
            p = f"encoder.layers.{i}."
            h = self.attention(self._ln(x, p + "layer_norm1"), i, mask)
            x = rb(h + x, m)
            h = self._ln(x, p + "layer_norm2")
            h = self._lin(quick_gelu(self._lin(h, p + "mlp.fc1"), m), p + "mlp.fc2")
            x = rb(h + x, m)
        return self._ln(x, "final_layer_norm")

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        """Module::forward text.rs:304-316: the hidden state at argmax(ids) (the EOS token has the largest id)."""
        h = self.hidden(ids)
        idx = ids.argmax(-1)
        return h[torch.arange(ids.shape[0]), idx]
