"""Host-side mirror of `AutoEncoderKl` / `VAEModel` (diffusion_rs_core/src/models/vaes/{mod,autoencoder_kl,vae}.rs)
over the C ABI — decode only (the pipeline never encodes)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import lib as L


@dataclass
class VaeConfig:  # autoencoder_kl.rs:15-32, FLUX.1 values
    latent_channels: int = 16
    out_channels: int = 3
    block_out_channels: tuple = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    mid_block_add_attention: bool = True
    scaling_factor: float = 0.3611
    shift_factor: float = 0.1159


class AutoEncoderKl:
    def __init__(self, cfg: VaeConfig):
        self.cfg = cfg
        self._lib = L.load()
        c = L.VaeConfigC(cfg.latent_channels, cfg.out_channels, (C.c_int32 * 4)(*cfg.block_out_channels),
                         cfg.layers_per_block, cfg.norm_num_groups, int(cfg.mid_block_add_attention),
                         cfg.scaling_factor, cfg.shift_factor)
        h = C.c_void_p()
        L.check(self._lib.fluxb200_vae_create(C.byref(c), C.byref(h)))
        self._h = h
        self._ws = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.fluxb200_vae_destroy(h)
            self._h = None

    @classmethod
    def new(cls, cfg: VaeConfig, tensors) -> "AutoEncoderKl":
        v = cls(cfg)
        items = tensors.items() if hasattr(tensors, "items") else tensors
        for name, t in items:
            v.load_weight(name, t)
        v.finalize()
        return v

    def load_weight(self, name: str, t: torch.Tensor):
        t = t.contiguous()
        if t.dtype != torch.bfloat16:
            raise L.Fluxb200Error("VAE weights must be bf16")
        shape = list(t.shape)
        arr = (C.c_int64 * len(shape))(*shape)
        L.check(self._lib.fluxb200_vae_load_weight(self._h, name.encode(), t.data_ptr(), 0, arr, len(shape),
                                                   1 if t.is_cuda else 0, L.current_stream()))
        if not t.is_cuda:
            torch.cuda.current_stream().synchronize()

    def finalize(self):
        L.check(self._lib.fluxb200_vae_finalize(self._h, L.current_stream()))

    def workspace(self, B, h, w):
        n = C.c_uint64()
        L.check(self._lib.fluxb200_vae_workspace_size(self._h, B, h, w, C.byref(n)))
        if self._ws is None or self._ws.numel() < n.value:
            self._ws = None
            self._ws = torch.empty(n.value, dtype=torch.uint8, device="cuda")
        return self._ws

    # VAEModel trait (vaes/mod.rs:15-28)
    def scale_factor(self) -> float:
        return self.cfg.scaling_factor

    def shift_factor(self) -> float:
        return self.cfg.shift_factor

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """z bf16 NCHW [B,16,h,w] -> bf16 NCHW [B,3,8h,8w]."""
        if z.dtype != torch.bfloat16 or not z.is_cuda or not z.is_contiguous() or z.dim() != 4:
            raise L.Fluxb200Error("decode expects a contiguous CUDA bf16 NCHW tensor")
        B, c, h, w = z.shape
        ws = self.workspace(B, h, w)
        out = torch.empty(B, self.cfg.out_channels, 8 * h, 8 * w, device="cuda", dtype=torch.bfloat16)
        L.check(self._lib.fluxb200_vae_decode(self._h, z.data_ptr(), out.data_ptr(), B, h, w, ws.data_ptr(),
                                              ws.numel(), L.current_stream()))
        return out

    def decode_packed_u8(self, packed: torch.Tensor, h2: int, w2: int, nchw: bool = False, out=None) -> torch.Tensor:
        """Tail of FluxPipeline::forward (flux/mod.rs:327-332) on packed latents [B, h2*w2, 64]."""
        B = packed.shape[0]
        ws = self.workspace(B, 2 * h2, 2 * w2)
        shape = (B, 3, 16 * h2, 16 * w2) if nchw else (B, 16 * h2, 16 * w2, 3)
        if out is None:
            out = torch.empty(*shape, device="cuda", dtype=torch.uint8)
        L.check(self._lib.fluxb200_vae_decode_packed_u8(self._h, packed.data_ptr(), out.data_ptr(), B, h2, w2,
                                                        int(nchw), ws.data_ptr(), ws.numel(), L.current_stream()))
        return out
