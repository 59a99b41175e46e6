"""Weight ingest: local diffusers directories and DDUF archives -> tensors under the names the
reference's VarBuilder serves.

Mirrors diffusion_rs_common/src/model_source.rs:97-145 (FileLoader over a hub snapshot directory or a zip-in-mmap
`.dduf`) and pipelines/mod.rs:136-209 (model_index.json must say FluxPipeline; per-component config + safetensors).
There is no network here, so `ModelSource::ModelId` accepts a local snapshot directory only.
Dtype policy = varbuilder.rs:312-323: every floating tensor is cast to the model dtype (bf16) at `vb.get`, except the
tensors fetched with `get_unchecked_dtype` (bitsandbytes side tensors, bitsandbytes/mod.rs:112-222), which keep theirs.
"""
from __future__ import annotations

import json
import zipfile
from pathlib import Path

import torch

from . import lib as L

# tensors BnbLinear reads with get_unchecked_dtype (bitsandbytes/mod.rs:118-190)
_KEEP_DTYPE_SUFFIXES = (".absmax", ".quant_map", ".nested_absmax", ".nested_quant_map", ".SCB",
                        ".quant_state.bitsandbytes__nf4", ".quant_state.bitsandbytes__fp4")


class DirLoader:
    def __init__(self, root: str | Path):
        self.root = Path(root)
        if not self.root.is_dir():
            raise L.Fluxb200Error(f"{root} is not a local model directory (hub download needs network access)")

    def exists(self, rel: str) -> bool:
        return (self.root / rel).exists()

    def read(self, rel: str) -> bytes:
        return (self.root / rel).read_bytes()

    def list(self, prefix: str) -> list[str]:
        d = self.root / prefix
        return sorted(str(p.relative_to(self.root)) for p in d.iterdir()) if d.is_dir() else []


class DdufLoader:
    """A DDUF file is an (uncompressed) zip of the diffusers directory layout (model_source.rs:62-95)."""

    def __init__(self, path: str | Path):
        self.zf = zipfile.ZipFile(path)
        self.names = set(self.zf.namelist())

    def exists(self, rel: str) -> bool:
        return rel in self.names

    def read(self, rel: str) -> bytes:
        return self.zf.read(rel)

    def list(self, prefix: str) -> list[str]:
        p = prefix.rstrip("/") + "/"
        return sorted(n for n in self.names if n.startswith(p) and "/" not in n[len(p):] and n != p)


def open_source(kind: str, location: str):
    if kind == "dduf":
        return DdufLoader(location)
    return DirLoader(location)


def _load_safetensors(loader, component: str) -> dict[str, torch.Tensor]:
    from safetensors.torch import load as st_load
    files = [f for f in loader.list(component) if f.endswith(".safetensors")]
    if not files:
        raise L.Fluxb200Error(f"no .safetensors files under {component}/")
    out: dict[str, torch.Tensor] = {}
    for f in files:  # the reference uses one loader thread per shard (varbuilder_loading.rs:72-86)
        out.update(st_load(loader.read(f)))
    return out


def cast_policy(name: str, t: torch.Tensor) -> torch.Tensor:
    if name.endswith(_KEEP_DTYPE_SUFFIXES):
        return t
    if name.endswith(".weight") and t.dtype in (torch.uint8, torch.int8):
        return t  # packed 4-bit / int8 weights
    if t.is_floating_point():
        return t.to(torch.bfloat16)
    return t


def load_flux_components(loader):
    """-> (flux_cfg dict, transformer tensors, vae_cfg dict, vae tensors, scheduler cfg dict)"""
    if not loader.exists("model_index.json"):
        raise L.Fluxb200Error("model_index.json not found")
    index = json.loads(loader.read("model_index.json"))
    if index.get("_class_name") != "FluxPipeline":  # pipelines/mod.rs:140-149
        raise L.Fluxb200Error(f"Unexpected loader type `{index.get('_class_name')}`, only FluxPipeline is supported")
    tcfg = json.loads(loader.read("transformer/config.json"))
    vcfg = json.loads(loader.read("vae/config.json"))
    scfg = json.loads(loader.read("scheduler/scheduler_config.json"))
    tr = {k: cast_policy(k, v) for k, v in _load_safetensors(loader, "transformer").items()}
    va = {k: cast_policy(k, v) for k, v in _load_safetensors(loader, "vae").items() if k.startswith("decoder.")}
    return tcfg, tr, vcfg, va, scfg


def flux_config_from_json(j: dict):
    from .transformer import FluxConfig
    return FluxConfig(in_channels=j["in_channels"], pooled_projection_dim=j["pooled_projection_dim"],
                      joint_attention_dim=j["joint_attention_dim"], num_attention_heads=j["num_attention_heads"],
                      num_layers=j["num_layers"], num_single_layers=j["num_single_layers"],
                      guidance_embeds=bool(j.get("guidance_embeds", False)))


def vae_config_from_json(j: dict):
    from .vae import VaeConfig
    # AutoEncoderKl::decode applies post_quant_conv first when the config asks for it (autoencoder_kl.rs:78-96,
    # 112-119); FLUX's VAE does not use it and the decoder here has no 1x1 conv for it: refuse rather than decode wrong
    if j.get("use_post_quant_conv", False) or j.get("use_quant_conv", False):
        raise L.Fluxb200Error("VAE configs with use_quant_conv / use_post_quant_conv are not supported "
                              "(FLUX.1's autoencoder sets both to false)")
    return VaeConfig(latent_channels=j["latent_channels"], out_channels=j["out_channels"],
                     block_out_channels=tuple(j["block_out_channels"]), layers_per_block=j["layers_per_block"],
                     norm_num_groups=j["norm_num_groups"],
                     mid_block_add_attention=bool(j.get("mid_block_add_attention", True)),
                     scaling_factor=float(j["scaling_factor"]), shift_factor=float(j.get("shift_factor", 0.0)))


def scheduler_config_from_json(j: dict):
    from .pipeline import SchedulerConfig
    if j.get("_class_name", "FlowMatchEulerDiscreteScheduler") != "FlowMatchEulerDiscreteScheduler":
        raise L.Fluxb200Error(f"unsupported scheduler {j.get('_class_name')}")
    return SchedulerConfig(base_image_seq_len=j.get("base_image_seq_len", 256), base_shift=j.get("base_shift", 0.5),
                           max_image_seq_len=j.get("max_image_seq_len", 4096), max_shift=j.get("max_shift", 1.15),
                           shift=j.get("shift", 1.0), use_dynamic_shifting=bool(j.get("use_dynamic_shifting", False)))


def load_text_components(loader):
    """Optional: the text encoders + tokenizers of a FLUX snapshot (pipelines/flux/mod.rs:60-140): `text_encoder`
    (CLIP, tensors under `text_model.`), `text_encoder_2` (T5), `tokenizer/{vocab.json,merges.txt}`,
    `tokenizer_2/tokenizer.json`.  Returns None when the snapshot does not carry them."""
    need = ("text_encoder/config.json", "text_encoder_2/config.json")
    if not all(loader.exists(n) for n in need):
        return None
    ccfg = json.loads(loader.read("text_encoder/config.json"))
    tcfg = json.loads(loader.read("text_encoder_2/config.json"))
    clip = {k[len("text_model."):]: cast_policy(k, v) for k, v in _load_safetensors(loader, "text_encoder").items()
            if k.startswith("text_model.") and not k.endswith("position_ids")}
    t5 = {k: cast_policy(k, v) for k, v in _load_safetensors(loader, "text_encoder_2").items()
          if k.startswith(("shared.", "encoder."))}
    toks = None
    if loader.exists("tokenizer_2/tokenizer.json") and loader.exists("tokenizer/vocab.json") and \
            loader.exists("tokenizer/merges.txt"):
        toks = {"t5": loader.read("tokenizer_2/tokenizer.json"), "clip_vocab": loader.read("tokenizer/vocab.json"),
                "clip_merges": loader.read("tokenizer/merges.txt")}
    return ccfg, clip, tcfg, t5, toks


def clip_config_from_json(j: dict):
    from .text_encoders import ClipTextConfig
    if j.get("hidden_act", "quick_gelu") != "quick_gelu":
        raise L.Fluxb200Error(f"unsupported CLIP activation {j.get('hidden_act')} (clip/text.rs:8-11 knows quick_gelu)")
    return ClipTextConfig(vocab_size=j["vocab_size"], projection_dim=j["projection_dim"],
                          intermediate_size=j["intermediate_size"], max_position_embeddings=j["max_position_embeddings"],
                          num_hidden_layers=j["num_hidden_layers"], num_attention_heads=j["num_attention_heads"])


def t5_config_from_json(j: dict):
    from .text_encoders import T5Config
    if j.get("feed_forward_proj", "gated-gelu") != "gated-gelu":
        raise L.Fluxb200Error(f"unsupported T5 feed_forward_proj {j.get('feed_forward_proj')}")
    return T5Config(vocab_size=j["vocab_size"], d_model=j["d_model"], d_kv=j["d_kv"], d_ff=j["d_ff"],
                    num_layers=j["num_layers"], num_heads=j["num_heads"],
                    relative_attention_num_buckets=j["relative_attention_num_buckets"],
                    relative_attention_max_distance=j.get("relative_attention_max_distance", 128),
                    layer_norm_epsilon=float(j["layer_norm_epsilon"]))
