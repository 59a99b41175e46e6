"""Host-side mirror of the two text encoders (`T5EncoderModel`, diffusion_rs_core/src/models/t5/mod.rs:637-661;
`ClipTextTransformer`, models/clip/text.rs:245-317) over the C ABI.

`X.new(cfg, tensors)` plays the role of `X::new(vb, cfg)`: every tensor the reference fetches through its VarBuilder is
handed to the library under the same checkpoint name.  Token ids come from the caller (tokenizers are out of scope).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import lib as L


@dataclass
class T5Config:  # t5/mod.rs:75-93 (google/t5-v1_1-xxl encoder)
    vocab_size: int = 32128
    d_model: int = 4096
    d_kv: int = 64
    d_ff: int = 10240
    num_layers: int = 24
    num_heads: int = 64
    relative_attention_num_buckets: int = 32
    relative_attention_max_distance: int = 128
    layer_norm_epsilon: float = 1e-6


@dataclass
class ClipTextConfig:  # clip/text.rs:22-31 (openai/clip-vit-large-patch14 text tower)
    vocab_size: int = 49408
    projection_dim: int = 768
    intermediate_size: int = 3072
    max_position_embeddings: int = 77
    num_hidden_layers: int = 12
    num_attention_heads: int = 12


class _Encoder:
    _prefix = ""

    def _create(self, cfg_c):
        self._lib = L.load()
        h = C.c_void_p()
        L.check(getattr(self._lib, f"fluxb200_{self._prefix}_create")(C.byref(cfg_c), C.byref(h)))
        self._h = h
        self._ws = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            getattr(self._lib, f"fluxb200_{self._prefix}_destroy")(h)
            self._h = None

    def load_weight(self, name: str, t: torch.Tensor):
        if t.dtype != torch.bfloat16:
            raise L.Fluxb200Error(f"text encoders take bf16 tensors, got {t.dtype} for {name}")
        t = t.contiguous()
        shape = list(t.shape)
        arr = (C.c_int64 * len(shape))(*shape)
        L.check(getattr(self._lib, f"fluxb200_{self._prefix}_load_weight")(self._h, name.encode(), t.data_ptr(), 0, arr,
                                                                            len(shape), 1 if t.is_cuda else 0,
                                                                            L.current_stream()))
        if not t.is_cuda:
            torch.cuda.current_stream().synchronize()

    def finalize(self):
        L.check(getattr(self._lib, f"fluxb200_{self._prefix}_finalize")(self._h, L.current_stream()))

    @classmethod
    def new(cls, cfg, tensors):
        m = cls(cfg)
        items = tensors.items() if hasattr(tensors, "items") else tensors
        for name, t in items:
            m.load_weight(name, t)
        m.finalize()
        return m

    def _workspace(self, B: int, Lq: int) -> torch.Tensor:
        n = C.c_uint64()
        L.check(getattr(self._lib, f"fluxb200_{self._prefix}_workspace_size")(self._h, B, Lq, C.byref(n)))
        if self._ws is None or self._ws.numel() < n.value:
            self._ws = None
            self._ws = torch.empty(n.value, dtype=torch.uint8, device="cuda")
        return self._ws

    @staticmethod
    def _ids(input_ids: torch.Tensor) -> torch.Tensor:
        if input_ids.dim() != 2:
            raise L.Fluxb200Error(f"input_ids must be [batch, seq], got {tuple(input_ids.shape)}")
        return input_ids.to(device="cuda", dtype=torch.int32).contiguous()


class T5EncoderModel(_Encoder):
    _prefix = "t5"

    def __init__(self, cfg: T5Config):
        self.cfg = cfg
        self._create(L.T5ConfigC(cfg.vocab_size, cfg.d_model, cfg.d_kv, cfg.d_ff, cfg.num_layers, cfg.num_heads,
                                 cfg.relative_attention_num_buckets, cfg.relative_attention_max_distance,
                                 cfg.layer_norm_epsilon))

    def forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        """ids [B, L] -> last hidden state bf16 [B, L, d_model] (T5EncoderModel::forward, t5/mod.rs:659)."""
        ids = self._ids(input_ids)
        B, Lq = ids.shape
        ws = self._workspace(B, Lq)
        out = torch.empty(B, Lq, self.cfg.d_model, device="cuda", dtype=torch.bfloat16)
        L.check(self._lib.fluxb200_t5_forward(self._h, ids.data_ptr(), out.data_ptr(), B, Lq, ws.data_ptr(), ws.numel(),
                                              L.current_stream()))
        return out


class ClipTextTransformer(_Encoder):
    _prefix = "clip"

    def __init__(self, cfg: ClipTextConfig):
        self.cfg = cfg
        self._create(L.ClipConfigC(cfg.vocab_size, cfg.projection_dim, cfg.intermediate_size,
                                   cfg.max_position_embeddings, cfg.num_hidden_layers, cfg.num_attention_heads))

    def forward_with_mask(self, input_ids: torch.Tensor) -> torch.Tensor:
        """ids [B, L] -> hidden states bf16 [B, L, D] after the final LayerNorm (text.rs:291-300, mask_after = MAX)."""
        return self._run(input_ids, want_hidden=True)[0]

    def forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        """ids [B, L] -> pooled bf16 [B, D]: the hidden state at argmax(ids), i.e. the EOS token (text.rs:304-316)."""
        return self._run(input_ids, want_hidden=False)[1]

    def _run(self, input_ids, want_hidden: bool):
        ids = self._ids(input_ids)
        B, Lq = ids.shape
        ws = self._workspace(B, Lq)
        D = self.cfg.projection_dim
        hidden = torch.empty(B, Lq, D, device="cuda", dtype=torch.bfloat16) if want_hidden else None
        pooled = torch.empty(B, D, device="cuda", dtype=torch.bfloat16)
        L.check(self._lib.fluxb200_clip_forward(self._h, ids.data_ptr(), L.ptr(hidden), pooled.data_ptr(), B, Lq,
                                                ws.data_ptr(), ws.numel(), L.current_stream()))
        return hidden, pooled
