"""ctypes binding of libfluxb200.so — the same C ABI a Rust `extern "C"` block would bind.

PyTorch is used by callers only to own device memory and streams; every compute call goes through the C ABI.
There is no CPU / eager fallback: if the library is missing or a call fails, a `Fluxb200Error` is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_LIB_PATH = _PKG / "libfluxb200.so"
_lib = None


class Fluxb200Error(RuntimeError):
    pass


c_void_p, c_int, c_int64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol declared in include/fluxb200.h
SIGNATURES = {
    "fluxb200_last_error": (C.c_char_p, []),
    "fluxb200_version": (c_int, []),
    "fluxb200_linear": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                                c_int, c_int, c_void_p, c_int64, c_int, c_void_p, c_float, c_void_p]),
    "fluxb200_linear_quant": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64,
                                      c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "fluxb200_sdpa": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float,
                              c_void_p]),
    "fluxb200_attn_variants": (c_int, []),
    "fluxb200_debug_sdpa_trace": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p,
                                          c_void_p]),
    "fluxb200_debug_gemm_trace": (c_int, [c_void_p]),
    "fluxb200_layernorm_modulate": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_int,
                                            c_float, c_void_p]),
    "fluxb200_qknorm_rope": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p]),
    # the reference's own FFI surface (bitsandbytes/ffi.rs:5-114)
    **{f"dequantize_blockwise_{t}_{q}": (None, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p])
       for t in ("f32", "f16", "bf16") for q in ("int8", "fp4", "nf4")},
    **{f"dequantize_8bit_kernel_{t}": (None, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int])
       for t in ("f32", "f16", "bf16")},
    "fluxb200_dequantize_q4k_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    # model level
    "fluxb200_model_create": (c_int, [c_void_p, C.POINTER(c_void_p)]),
    "fluxb200_model_destroy": (None, [c_void_p]),
    "fluxb200_model_load_weight": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, C.POINTER(c_int64), c_int, c_int,
                                           c_void_p]),
    "fluxb200_model_finalize": (c_int, [c_void_p, c_void_p]),
    "fluxb200_model_workspace_size": (c_int, [c_void_p, c_int, c_int, c_int, C.POINTER(C.c_uint64)]),
    "fluxb200_model_denoise_workspace_size": (c_int, [c_void_p, c_int, c_int, c_int, c_int, C.POINTER(C.c_uint64)]),
    "fluxb200_model_denoise_info": (c_int, [c_void_p, C.POINTER(c_int), C.POINTER(C.c_char_p)]),
    "fluxb200_model_forward": (c_int, [c_void_p] + [c_void_p] * 8 + [c_int, c_int, c_int, c_void_p, C.c_uint64,
                                                                      c_void_p]),
    "fluxb200_model_denoise": (c_int, [c_void_p] + [c_void_p] * 5 + [c_float, C.POINTER(C.c_double), c_int, c_int,
                                                                      c_int, c_int, c_void_p, C.c_uint64, c_void_p]),
    "fluxb200_model_tap": (c_int, [c_void_p, c_int, c_void_p, C.c_uint64, c_void_p]),
    # VAE
    "fluxb200_vae_create": (c_int, [c_void_p, C.POINTER(c_void_p)]),
    "fluxb200_vae_destroy": (None, [c_void_p]),
    "fluxb200_vae_load_weight": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, C.POINTER(c_int64), c_int, c_int,
                                         c_void_p]),
    "fluxb200_vae_finalize": (c_int, [c_void_p, c_void_p]),
    "fluxb200_vae_workspace_size": (c_int, [c_void_p, c_int, c_int, c_int, C.POINTER(C.c_uint64)]),
    "fluxb200_vae_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, C.c_uint64,
                                    c_void_p]),
    "fluxb200_vae_decode_packed_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                              C.c_uint64, c_void_p]),
    # text encoders
    "fluxb200_t5_create": (c_int, [c_void_p, C.POINTER(c_void_p)]),
    "fluxb200_t5_destroy": (None, [c_void_p]),
    "fluxb200_t5_load_weight": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, C.POINTER(c_int64), c_int, c_int,
                                        c_void_p]),
    "fluxb200_t5_finalize": (c_int, [c_void_p, c_void_p]),
    "fluxb200_t5_workspace_size": (c_int, [c_void_p, c_int, c_int, C.POINTER(C.c_uint64)]),
    "fluxb200_t5_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, C.c_uint64, c_void_p]),
    "fluxb200_clip_create": (c_int, [c_void_p, C.POINTER(c_void_p)]),
    "fluxb200_clip_destroy": (None, [c_void_p]),
    "fluxb200_clip_load_weight": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, C.POINTER(c_int64), c_int, c_int,
                                          c_void_p]),
    "fluxb200_clip_finalize": (c_int, [c_void_p, c_void_p]),
    "fluxb200_clip_workspace_size": (c_int, [c_void_p, c_int, c_int, C.POINTER(C.c_uint64)]),
    "fluxb200_clip_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, C.c_uint64,
                                      c_void_p]),
    "fluxb200_conv2d_nhwc": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_void_p]),
    "fluxb200_repack_conv_weight": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "fluxb200_set_flag": (c_int, [C.c_char_p, c_int]),
    "fluxb200_profile_enable": (None, [c_int]),
    "fluxb200_profile_kinds": (c_int, []),
    "fluxb200_profile_kind_name": (C.c_char_p, [c_int]),
    "fluxb200_profile_collect": (c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                         C.POINTER(C.c_uint64)]),
    "fluxb200_launch_count": (C.c_uint64, [c_int]),
    "fluxb200_groupnorm_nhwc": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_float, c_int, c_void_p, c_void_p]),
}


class VaeConfigC(C.Structure):
    _fields_ = [("latent_channels", c_int), ("out_channels", c_int), ("block_out_channels", c_int * 4),
                ("layers_per_block", c_int), ("norm_num_groups", c_int), ("mid_block_add_attention", c_int),
                ("scaling_factor", c_float), ("shift_factor", c_float)]


class T5ConfigC(C.Structure):
    _fields_ = [(n, c_int) for n in ("vocab_size", "d_model", "d_kv", "d_ff", "num_layers", "num_heads",
                                     "relative_attention_num_buckets", "relative_attention_max_distance")] + [
        ("layer_norm_epsilon", c_float)]


class ClipConfigC(C.Structure):
    _fields_ = [(n, c_int) for n in ("vocab_size", "projection_dim", "intermediate_size", "max_position_embeddings",
                                     "num_hidden_layers", "num_attention_heads")]


class FluxConfigC(C.Structure):
    _fields_ = [(n, c_int) for n in ("in_channels", "pooled_projection_dim", "joint_attention_dim",
                                     "num_attention_heads", "num_layers", "num_single_layers", "guidance_embeds")]


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (building is `diffusion_rs_b200.build.build()`); fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise Fluxb200Error(
            f"{_LIB_PATH} not found: build it with `python -m diffusion_rs_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().fluxb200_last_error()
        raise Fluxb200Error(msg.decode() if msg else f"fluxb200 call failed with status {rc}")


def ptr(t) -> int | None:
    """Device (or host) pointer of a torch tensor, None -> NULL."""
    if t is None:
        return None
    return t.data_ptr()


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
