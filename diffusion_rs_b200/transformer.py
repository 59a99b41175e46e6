"""Host-side mirror of `Flux` (diffusion_rs_core/src/models/flux/model.rs:709-838) over the C ABI.

`FluxTransformer.new(cfg, tensors)` plays the role of `Flux::new(cfg, vb)`: every tensor the reference would fetch
through its VarBuilder is handed to the library under the same diffusers name; `forward` is `Flux::forward`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import lib as L

DT = {torch.bfloat16: 0, torch.float32: 1, torch.uint8: 2, torch.int8: 3, torch.float16: 4}
DT_Q4K = 10


@dataclass
class FluxConfig:  # model.rs:21-31
    in_channels: int = 64
    pooled_projection_dim: int = 768
    joint_attention_dim: int = 4096
    num_attention_heads: int = 24
    num_layers: int = 19
    num_single_layers: int = 38
    guidance_embeds: bool = True


class FluxTransformer:
    def __init__(self, cfg: FluxConfig):
        self.cfg = cfg
        self._lib = L.load()
        c = L.FluxConfigC(cfg.in_channels, cfg.pooled_projection_dim, cfg.joint_attention_dim,
                          cfg.num_attention_heads, cfg.num_layers, cfg.num_single_layers, int(cfg.guidance_embeds))
        h = C.c_void_p()
        L.check(self._lib.fluxb200_model_create(C.byref(c), C.byref(h)))
        self._h = h
        self._ws = None
        self._finalized = False

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.fluxb200_model_destroy(h)
            self._h = None

    # -- Flux::new ---------------------------------------------------------------------------------------
    def load_weight(self, name: str, t: torch.Tensor, dtype_code: int | None = None, logical_shape=None):
        t = t.contiguous()
        code = DT[t.dtype] if dtype_code is None else dtype_code
        shape = list(t.shape) if logical_shape is None else list(logical_shape)
        arr = (C.c_int64 * len(shape))(*shape)
        L.check(self._lib.fluxb200_model_load_weight(self._h, name.encode(), t.data_ptr(), code, arr, len(shape),
                                                     1 if t.is_cuda else 0, L.current_stream()))
        if not t.is_cuda:
            torch.cuda.current_stream().synchronize()  # pageable host memory: keep `t` alive until copied

    def finalize(self):
        L.check(self._lib.fluxb200_model_finalize(self._h, L.current_stream()))
        self._finalized = True

    @classmethod
    def new(cls, cfg: FluxConfig, tensors) -> "FluxTransformer":
        """tensors: mapping or iterable of (name, tensor)."""
        m = cls(cfg)
        items = tensors.items() if hasattr(tensors, "items") else tensors
        for name, t in items:
            m.load_weight(name, t)
        m.finalize()
        return m

    # -- workspace (caller-owned, as every activation buffer) --------------------------------------------
    def workspace(self, B: int, l_img: int, l_txt: int, n_timesteps: int | None = None) -> torch.Tensor:
        n = C.c_uint64()
        if n_timesteps is None:
            L.check(self._lib.fluxb200_model_workspace_size(self._h, B, l_img, l_txt, C.byref(n)))
        else:  # the denoising loop keeps the per-step tables of the whole image in the workspace
            L.check(self._lib.fluxb200_model_denoise_workspace_size(self._h, B, l_img, l_txt, n_timesteps,
                                                                    C.byref(n)))
        if self._ws is None or self._ws.numel() < n.value:
            self._ws = None
            self._ws = torch.empty(n.value, dtype=torch.uint8, device="cuda")
        return self._ws

    # -- Flux::forward (model.rs:790-833) -----------------------------------------------------------------
    def forward(self, img, img_ids, txt, txt_ids, timesteps, y, guidance=None):
        B, l_img, _ = img.shape
        l_txt = txt.shape[1]
        for t in (img, img_ids, txt, txt_ids, y):
            if t.dtype != torch.bfloat16 or not t.is_cuda or not t.is_contiguous():
                raise L.Fluxb200Error("forward expects contiguous CUDA bf16 tensors")
        if img.dim() != 3 or txt.dim() != 3:
            raise L.Fluxb200Error(f"unexpected shape for img {tuple(img.shape)} / txt {tuple(txt.shape)}")
        timesteps = timesteps.to(device="cuda", dtype=torch.float32).contiguous()
        g = None if guidance is None else guidance.to(device="cuda", dtype=torch.float32).contiguous()
        ws = self.workspace(B, l_img, l_txt)
        out = torch.empty(B, l_img, self.cfg.in_channels, device="cuda", dtype=torch.bfloat16)
        L.check(self._lib.fluxb200_model_forward(self._h, img.data_ptr(), img_ids.data_ptr(), txt.data_ptr(),
                                                 txt_ids.data_ptr(), timesteps.data_ptr(), y.data_ptr(), L.ptr(g),
                                                 out.data_ptr(), B, l_img, l_txt, ws.data_ptr(), ws.numel(),
                                                 L.current_stream()))
        return out

    # -- Sampler::sample (pipelines/sampling.rs:25-48) ----------------------------------------------------
    def denoise(self, img, img_ids, txt, txt_ids, y, guidance_scale: float, timesteps: list[float]):
        """In-place Euler loop over `timesteps` on img [B,l_img,64]."""
        B, l_img, _ = img.shape
        l_txt = txt.shape[1]
        ws = self.workspace(B, l_img, l_txt, len(timesteps))
        ts = (C.c_double * len(timesteps))(*timesteps)
        L.check(self._lib.fluxb200_model_denoise(self._h, img.data_ptr(), img_ids.data_ptr(), txt.data_ptr(),
                                                 txt_ids.data_ptr(), y.data_ptr(), float(guidance_scale), ts,
                                                 len(timesteps), B, l_img, l_txt, ws.data_ptr(), ws.numel(),
                                                 L.current_stream()))
        return img

    def denoise_info(self) -> tuple[bool, str]:
        """(steps of the last denoise were CUDA-graph replays, note explaining why not)."""
        used, note = C.c_int32(0), C.c_char_p()
        L.check(self._lib.fluxb200_model_denoise_info(self._h, C.byref(used), C.byref(note)))
        return bool(used.value), (note.value or b"").decode()

    def tap(self, which: int, shape) -> torch.Tensor:
        out = torch.empty(*shape, device="cuda", dtype=torch.bfloat16)
        L.check(self._lib.fluxb200_model_tap(self._h, which, out.data_ptr(), out.numel() * 2, L.current_stream()))
        return out
