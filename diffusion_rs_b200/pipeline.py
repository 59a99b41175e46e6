"""Host-side mirror of the reference's public API: `Pipeline::load` / `Pipeline::forward`,
`DiffusionGenerationParams`, `ModelSource`, `TokenSource`, `ModelDType`, `Offloading`
(diffusion_rs_core/src/lib.rs:43-45, pipelines/mod.rs:23-33, 110-270; Python surface diffusion_rs_py/src/lib.rs).

What is B200-native here is the denoising hot path (`FluxPipeline::forward`, pipelines/flux/mod.rs:225-335, from the
noise latents to the u8 image) and, as the first widening step, the two text encoders in front
of it (T5 / CLIP, text_encoders.py).  A prompt is one of: `PromptEmbeds` (encoder outputs computed elsewhere),
`PromptTokens` (token ids -> the encoders run on the GPU), or a string (tokenised when the snapshot ships its
tokenizers; for `ModelSource.synthetic` pipelines only - random-init weights - mapped to deterministic synthetic
embeddings so the API stays runnable without weights or network; anything else raises).
Multi-GPU: one process per GPU; prompts are sharded across ranks (independent trajectories — no per-step collective);
the only NCCL traffic is the weight broadcast at load.
"""
from __future__ import annotations

import enum
import math
import zlib
from dataclasses import dataclass

import torch

from . import lib as L
from .text_encoders import ClipTextConfig, ClipTextTransformer, T5Config, T5EncoderModel
from .transformer import FluxConfig, FluxTransformer
from .vae import AutoEncoderKl, VaeConfig


# ------------------------------------------------------------------------------------------------------------------
# API enums / params (names and meaning as in the reference)
# ------------------------------------------------------------------------------------------------------------------
class ModelDType(enum.Enum):  # util/auto_dtype.rs:9-20
    Auto = "auto"
    BF16 = "bf16"
    F16 = "f16"
    F32 = "f32"


class Offloading(enum.Enum):  # pipelines/mod.rs:27-33
    Full = "full"


class TokenSource(enum.Enum):  # diffusion_rs_common tokens.rs (only the variants that make sense offline)
    CacheToken = "cache"
    Nothing = "none"


@dataclass
class ModelSource:  # diffusion_rs_common/src/model_source.rs:17-60
    kind: str                 # "model_id" | "dduf" | "synthetic" | "tensors"
    model_id: str = ""
    quant: str | None = None  # synthetic only: None | "nf4" | "q4k"
    transformer: dict | None = None
    vae: dict | None = None
    num_layers: int | None = None         # synthetic only: reduced depth for tests
    num_single_layers: int | None = None
    text_encoders: bool = False           # synthetic: also build random-init T5-XXL + CLIP-L
    t5: dict | None = None                # tensors: optional text encoder checkpoints + their configs
    clip: dict | None = None
    t5_config: T5Config | None = None
    clip_config: ClipTextConfig | None = None

    @staticmethod
    def from_model_id(model_id: str) -> "ModelSource":
        return ModelSource("model_id", model_id)

    @staticmethod
    def dduf(path: str) -> "ModelSource":
        return ModelSource("dduf", path)

    @staticmethod
    def synthetic(model_id: str = "black-forest-labs/FLUX.1-dev", quant=None, num_layers=None,
                  num_single_layers=None, text_encoders: bool = False) -> "ModelSource":
        return ModelSource("synthetic", model_id, quant, None, None, num_layers, num_single_layers, text_encoders)

    @staticmethod
    def tensors(model_id: str, transformer: dict, vae: dict) -> "ModelSource":
        """Already-loaded checkpoint tensors under their diffusers names (what the reference's VarBuilder serves)."""
        return ModelSource("tensors", model_id, None, transformer, vae)


@dataclass
class DiffusionGenerationParams:  # pipelines/mod.rs:103-108
    height: int
    width: int
    num_steps: int
    guidance_scale: float


@dataclass
class PromptEmbeds:
    """Output of the (out-of-scope) text encoders for one prompt: T5 `txt` [l_txt, 4096], CLIP pooled `vec` [768]."""
    txt: torch.Tensor
    vec: torch.Tensor


@dataclass
class PromptTokens:
    """Token ids of one prompt: what the reference's tokenizers produce (flux/mod.rs:203-222).  `t5_ids` / `clip_ids`
    are 1-D integer tensors; batches are padded with 0 to the longest prompt exactly like tokenize_and_pad."""
    t5_ids: torch.Tensor
    clip_ids: torch.Tensor


@dataclass
class SchedulerConfig:  # pipelines/scheduler.rs:4-15 (FlowMatchEulerDiscreteScheduler)
    base_image_seq_len: int = 256
    base_shift: float = 0.5
    max_image_seq_len: int = 4096
    max_shift: float = 1.15
    shift: float = 3.0
    use_dynamic_shifting: bool = True

    def get_timesteps(self, num_steps: int, mu: float | None) -> list[float]:  # scheduler.rs:28-51
        sigmas = [v / num_steps for v in range(num_steps, -1, -1)]
        if self.use_dynamic_shifting:
            if mu is None:
                raise L.Fluxb200Error("`mu` is required for dynamic shifting")
            e = math.exp(mu)
            return [e / (e + (1.0 / s - 1.0)) if s > 0.0 else 0.0 for s in sigmas]
        return [self.shift * s / (1.0 + (self.shift - 1.0) * s) for s in sigmas]


def calculate_shift(image_seq_len, base_seq_len, max_seq_len, base_shift, max_shift) -> float:  # flux/sampling.rs:70-80
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def shift_seq_len(noise_shape, mode: str = "reference") -> int:
    """The `image_seq_len` FluxPipeline::forward feeds to calculate_shift.  The reference passes `img.dims()[1]` of the
    UNPACKED noise [bs, 16, h, w] (pipelines/flux/mod.rs:279, `img` is only packed inside State::new), i.e. the latent
    channel count 16 whatever the resolution -> mu = 0.4594 and one sigma schedule for every image size.  BFL / diffusers
    pass the packed sequence length (h/2)*(w/2) (mu = 1.15 at 1024^2).  A drop-in mirrors the reference ("reference",
    the default); "bfl" selects the upstream-FLUX behaviour."""
    if mode == "reference":
        return int(noise_shape[1])
    if mode == "bfl":
        return (int(noise_shape[2]) // 2) * (int(noise_shape[3]) // 2)
    raise L.Fluxb200Error(f"unknown shift mode {mode!r} (reference | bfl)")


def latent_hw(height: int, width: int) -> tuple[int, int]:  # get_noise flux/sampling.rs:11-12
    return (height + 15) // 16 * 2, (width + 15) // 16 * 2


def patchify(lat: torch.Tensor) -> torch.Tensor:  # State::new flux/sampling.rs:29-31
    b, c, h, w = lat.shape
    return lat.reshape(b, c, h // 2, 2, w // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(b, (h // 2) * (w // 2), c * 4)


def make_ids(h2: int, w2: int, l_txt: int, dtype=torch.bfloat16):  # State::new flux/sampling.rs:32-49
    img_ids = torch.zeros(h2, w2, 3)
    img_ids[..., 1] = torch.arange(h2)[:, None]
    img_ids[..., 2] = torch.arange(w2)[None, :]
    return img_ids.reshape(h2 * w2, 3).to(dtype), torch.zeros(l_txt, 3, dtype=dtype)


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice of the prompt batch owned by `rank` (prompts [0..N) -> N/world per rank)."""
    per = (n + world - 1) // world
    return min(n, rank * per), min(n, (rank + 1) * per)


def dist_info():
    d = torch.distributed
    if d.is_available() and d.is_initialized():
        return d, d.get_rank(), d.get_world_size()
    return None, 0, 1


def broadcast_weight(t: torch.Tensor) -> torch.Tensor:
    """The only collective on the path: rank 0's weights are broadcast once at load (NCCL on GPUs, gloo in tests)."""
    d, _, world = dist_info()
    if world > 1:
        d.broadcast(t, src=0)
    return t


NOISE_SEED = 299792458  # the reference seeds cuRAND with this constant (cuda_backend/device.rs:206)


class Pipeline:
    """`diffusion_rs_core::Pipeline` (pipelines/mod.rs:110-270)."""

    def __init__(self, transformer: FluxTransformer, vae: AutoEncoderKl, scheduler: SchedulerConfig, is_dev: bool,
                 t5: T5EncoderModel | None = None, clip: ClipTextTransformer | None = None, tokenizers=None,
                 synthetic: bool = False):
        self.transformer, self.vae, self.scheduler, self.is_dev = transformer, vae, scheduler, is_dev
        self.t5, self.clip, self.tokenizers = t5, clip, tokenizers
        self.synthetic = synthetic      # random-init weights: string prompts map to deterministic synthetic embeddings
        self.shift_mode = "reference"   # see shift_seq_len
        self.max_batch = 4  # images per denoise call on one GPU
        self._pinned = {}

    # -- Pipeline::load (pipelines/mod.rs:120-236) -----------------------------------------------------------------
    @classmethod
    def load(cls, source: ModelSource, silent: bool = False, token: TokenSource = TokenSource.CacheToken,
             revision: str | None = None, offloading: Offloading | None = None,
             dtype: ModelDType = ModelDType.Auto) -> "Pipeline":
        if not torch.cuda.is_available():
            raise L.Fluxb200Error("fluxb200 needs a B200 (sm_100a) GPU; there is no CPU fallback")
        if dtype not in (ModelDType.Auto, ModelDType.BF16):
            raise L.Fluxb200Error("the B200 hot path computes in bf16 (ModelDType::Auto resolves to BF16 on sm_100)")
        if offloading is not None:
            raise L.Fluxb200Error("Offloading::Full is pointless with 180 GB of HBM and is not implemented")
        bcast = broadcast_weight  # NCCL broadcast of weights at load only (north_star / SURVEY §8(e))
        if source.kind in ("model_id", "dduf"):
            # FileLoader::from_model_source (model_source.rs:97-145): a local snapshot directory or a .dduf archive
            from . import ingest
            loader = ingest.open_source(source.kind, source.model_id)
            tj, tr_tensors, vj, va_tensors, sj = ingest.load_flux_components(loader)
            fcfg, vcfg, sched = (ingest.flux_config_from_json(tj), ingest.vae_config_from_json(vj),
                                 ingest.scheduler_config_from_json(sj))
            tr, va = FluxTransformer(fcfg), AutoEncoderKl(vcfg)
            for name, t in tr_tensors.items():
                tr.load_weight(name, bcast(t.cuda()))
            for name, t in va_tensors.items():
                va.load_weight(name, bcast(t.cuda()))
            tr.finalize()
            va.finalize()
            t5 = clip = toks = None
            text = ingest.load_text_components(loader)
            if text is not None:
                ccfg, clip_t, t5cfg, t5_t, tok_bytes = text
                clip = ClipTextTransformer(ingest.clip_config_from_json(ccfg))
                for name, t in clip_t.items():
                    clip.load_weight(name, bcast(t.cuda()))
                clip.finalize()
                t5 = T5EncoderModel(ingest.t5_config_from_json(t5cfg))
                for name, t in t5_t.items():
                    t5.load_weight(name, bcast(t.cuda()))
                t5.finalize()
                toks = cls._build_tokenizers(tok_bytes)
            torch.cuda.synchronize()
            return cls(tr, va, sched, fcfg.guidance_embeds, t5, clip, toks)
        is_dev = "schnell" not in source.model_id.lower()
        fcfg = FluxConfig(guidance_embeds=is_dev)
        if source.num_layers is not None:
            fcfg.num_layers = source.num_layers
        if source.num_single_layers is not None:
            fcfg.num_single_layers = source.num_single_layers
        vcfg = VaeConfig()
        sched = SchedulerConfig(use_dynamic_shifting=is_dev, shift=3.0 if is_dev else 1.0)

        tr = FluxTransformer(fcfg)
        va = AutoEncoderKl(vcfg)
        if source.kind == "synthetic":
            from . import synthetic as S
            if source.quant is None:
                for name, t in S.iter_flux_tensors(fcfg):
                    tr.load_weight(name, bcast(t))
            else:
                from . import quantize as QZ
                for name, t in S.iter_flux_tensors(fcfg):
                    for qname, qt, code, shape in QZ.quantize_tensor(name, t, source.quant):
                        tr.load_weight(qname, bcast(qt), code, shape)
            for name, t in S.iter_vae_tensors(vcfg):
                va.load_weight(name, bcast(t))
        else:
            for name, t in source.transformer.items():
                tr.load_weight(name, bcast(t.cuda()))
            for name, t in source.vae.items():
                va.load_weight(name, bcast(t.cuda().to(torch.bfloat16)))
        tr.finalize()
        va.finalize()
        t5 = clip = None
        if source.kind == "synthetic" and source.text_encoders:
            from . import synthetic as S
            t5, clip = T5EncoderModel(T5Config()), ClipTextTransformer(ClipTextConfig())
            for name, t in S.iter_t5_tensors(t5.cfg):
                t5.load_weight(name, bcast(t))
            for name, t in S.iter_clip_tensors(clip.cfg):
                clip.load_weight(name, bcast(t))
            t5.finalize()
            clip.finalize()
        elif source.kind == "tensors" and source.t5 is not None and source.clip is not None:
            t5, clip = T5EncoderModel(source.t5_config), ClipTextTransformer(source.clip_config)
            for name, t in source.t5.items():
                t5.load_weight(name, bcast(t.cuda()))
            for name, t in source.clip.items():
                clip.load_weight(name, bcast(t.cuda()))
            t5.finalize()
            clip.finalize()
        torch.cuda.synchronize()
        return cls(tr, va, sched, is_dev, t5, clip, synthetic=(source.kind == "synthetic"))

    # -- helpers ------------------------------------------------------------------------------------------------------
    def text_len(self) -> int:
        return 512 if self.is_dev else 256  # dev benchmarked at 512, schnell pads to 256

    def synthetic_embeds(self, prompt: str) -> PromptEmbeds:
        g = torch.Generator().manual_seed(zlib.crc32(prompt.encode()))
        txt = torch.randn(self.text_len(), self.transformer.cfg.joint_attention_dim, generator=g).to(torch.bfloat16)
        vec = torch.randn(self.transformer.cfg.pooled_projection_dim, generator=g).to(torch.bfloat16)
        return PromptEmbeds(txt, vec)

    @staticmethod
    def _build_tokenizers(tok_bytes):
        """tokenizer_2/tokenizer.json (T5: `Tokenizer::from_bytes`, flux/mod.rs:82-88) and tokenizer/{vocab.json,
        merges.txt} (CLIP: `load_bpe_tokenizer`, diffusion_rs_common/src/tokenizer.rs:7-23 — a BARE BPE model: no
        normaliser, no pre-tokeniser, no BOS/EOS, merges read from line 2 on; mirrored as is)."""
        if tok_bytes is None:
            return None
        import json as _json
        try:
            from tokenizers import Tokenizer
            from tokenizers.models import BPE
        except ImportError as e:  # the snapshot ships tokenizer files: not being able to read them is an error
            raise L.Fluxb200Error("the snapshot ships tokenizers but the `tokenizers` package is not importable; "
                                  "install it or pass PromptTokens / PromptEmbeds") from e
        t5_tok = Tokenizer.from_str(tok_bytes["t5"].decode())
        vocab = _json.loads(tok_bytes["clip_vocab"].decode())
        merges = [tuple(x.split(" ")) for x in tok_bytes["clip_merges"].decode().split("\n")[1:]]
        merges = [m for m in merges if len(m) == 2]
        clip_tok = Tokenizer(BPE(vocab=vocab, merges=merges))
        return {"t5": t5_tok, "clip": clip_tok}

    def tokenize(self, prompt: str) -> PromptTokens:
        """`tokenizer.encode_batch(prompts, true)` -> ids (flux/mod.rs:203-222)."""
        if self.tokenizers is None:
            raise L.Fluxb200Error("this pipeline was loaded without tokenizers; pass PromptTokens or PromptEmbeds")
        t5_ids = self.tokenizers["t5"].encode(prompt, add_special_tokens=True).ids
        clip_ids = self.tokenizers["clip"].encode(prompt, add_special_tokens=True).ids
        return PromptTokens(torch.tensor(t5_ids, dtype=torch.int64), torch.tensor(clip_ids, dtype=torch.int64))

    def encode_prompts(self, prompts: list[PromptTokens]) -> list[PromptEmbeds]:
        """tokenize_and_pad + t5_model.forward + clip_model.forward (flux/mod.rs:236-268): device-resident embeddings."""
        if self.t5 is None or self.clip is None:
            raise L.Fluxb200Error("this pipeline was loaded without text encoders; pass PromptEmbeds")

        def pad(seqs, min_len=0):
            n = max(max(len(s) for s in seqs), min_len)
            out = torch.zeros(len(seqs), n, dtype=torch.int32)
            for i, s in enumerate(seqs):
                out[i, :len(s)] = s.to(torch.int32)
            return out

        t5_ids = pad([p.t5_ids for p in prompts])
        if not self.is_dev:  # schnell: pad to exactly 256, longer prompts are an error (flux/mod.rs:243-253)
            if t5_ids.shape[1] > 256:
                raise L.Fluxb200Error("T5 embedding length greater than 256, please shrink the prompt or use the -dev "
                                      "(with guidance distillation) version.")
            t5_ids = pad([p.t5_ids for p in prompts], 256)
        clip_ids = pad([p.clip_ids for p in prompts])
        txt = self.t5.forward(t5_ids)      # [N, L, 4096]
        vec = self.clip.forward(clip_ids)  # [N, 768]
        return [PromptEmbeds(txt[i], vec[i]) for i in range(len(prompts))]

    def _pin(self, key, shape, dtype):
        t = self._pinned.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(*shape, dtype=dtype, pin_memory=True)
            self._pinned[key] = t
        return t

    # -- Pipeline::forward (pipelines/mod.rs:241-270 -> FluxPipeline::forward flux/mod.rs:225-335) -----------------
    def forward(self, prompts, params: DiffusionGenerationParams, noise: torch.Tensor | None = None,
                seed: int = NOISE_SEED) -> list[torch.Tensor]:
        """N prompts -> N images (HWC uint8 host tensors).  `prompts` are strings or PromptEmbeds."""
        prompts = list(prompts)
        if not prompts:
            return []
        if self.tokenizers is not None and self.t5 is not None:
            prompts = [self.tokenize(p) if isinstance(p, str) else p for p in prompts]
        tok_idx = [i for i, p in enumerate(prompts) if isinstance(p, PromptTokens)]
        if tok_idx:
            for i, e in zip(tok_idx, self.encode_prompts([prompts[i] for i in tok_idx])):
                prompts[i] = e
        if any(isinstance(p, str) for p in prompts) and not self.synthetic:
            # a real checkpoint must never answer a text prompt with an image of random embeddings
            raise L.Fluxb200Error("this pipeline cannot encode text prompts (no tokenizers / text encoders in the "
                                  "snapshot); pass PromptTokens or PromptEmbeds")
        embeds = [p if isinstance(p, PromptEmbeds) else self.synthetic_embeds(p) for p in prompts]
        n = len(embeds)
        h, w = latent_hw(params.height, params.width)
        if noise is None:  # get_noise flux/sampling.rs:5-14 -> .to_dtype(bf16) flux/mod.rs:276
            noise = torch.randn(n, 16, h, w, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)
        noise = noise.to(torch.bfloat16)
        # shard the prompt batch over ranks (contiguous slices); every rank returns only its own images
        _, rank, world = dist_info()
        lo, hi = shard_range(n, rank, world)
        images = []
        for s in range(lo, hi, self.max_batch):
            e = min(hi, s + self.max_batch)
            images += self._generate(embeds[s:e], noise[s:e], params)
        return images

    def forward_png(self, prompts, params: DiffusionGenerationParams, **kw) -> list[bytes]:
        """What the reference's Python binding returns (`diffuse_rs.pyi`: `Pipeline.forward(...) -> list[bytes]`,
        diffusion_rs_py/src/lib.rs:140-154): one PNG-encoded image per prompt of this rank's shard."""
        from .image import encode_png
        return [encode_png(im) for im in self.forward(prompts, params, **kw)]

    def _generate(self, embeds, noise, params) -> list[torch.Tensor]:
        B = len(embeds)
        l_txt = embeds[0].txt.shape[0]
        if any(e.txt.shape[0] != l_txt for e in embeds):
            raise L.Fluxb200Error("prompts in one batch must have the same T5 length (the reference pads to the max)")
        _, _, h, w = noise.shape
        h2, w2 = h // 2, w // 2
        l_img = h2 * w2
        # host staging in pinned memory, then H2D on the compute stream
        txt_h = self._pin("txt", (B, l_txt, embeds[0].txt.shape[1]), torch.bfloat16)
        vec_h = self._pin("vec", (B, embeds[0].vec.shape[0]), torch.bfloat16)
        img_h = self._pin("img", (B, l_img, 64), torch.bfloat16)
        if all(e.txt.is_cuda and e.vec.is_cuda for e in embeds):  # outputs of our own encoders: stay on the device
            txt = torch.stack([e.txt for e in embeds]).contiguous()
            vec = torch.stack([e.vec for e in embeds]).contiguous()
        else:
            for i, e in enumerate(embeds):
                txt_h[i].copy_(e.txt)
                vec_h[i].copy_(e.vec)
            txt = txt_h.to("cuda", non_blocking=True)
            vec = vec_h.to("cuda", non_blocking=True)
        img_h.copy_(patchify(noise))
        img = img_h.to("cuda", non_blocking=True)
        img_ids1, txt_ids1 = make_ids(h2, w2, l_txt)
        img_ids = img_ids1[None].repeat(B, 1, 1).contiguous().cuda()
        txt_ids = txt_ids1[None].repeat(B, 1, 1).contiguous().cuda()
        sc = self.scheduler
        mu = calculate_shift(shift_seq_len(noise.shape, self.shift_mode), sc.base_image_seq_len, sc.max_image_seq_len,
                             sc.base_shift, sc.max_shift)
        timesteps = sc.get_timesteps(params.num_steps, mu)
        self.transformer.denoise(img, img_ids, txt, txt_ids, vec, params.guidance_scale, timesteps)
        out_d = self.vae.decode_packed_u8(img, h2, w2)  # [B, 16*h2, 16*w2, 3] u8
        out_h = self._pin("out", tuple(out_d.shape), torch.uint8)
        out_h.copy_(out_d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        H, W = params.height, params.width
        return [out_h[i, :H, :W].clone() for i in range(B)]

    # bytes moved per generated batch, for bench.py's e2e accounting
    def io_bytes(self, B: int, params: DiffusionGenerationParams) -> tuple[int, int]:
        h, w = latent_hw(params.height, params.width)
        l_img = (h // 2) * (w // 2)
        cfg = self.transformer.cfg
        h2d = B * (self.text_len() * cfg.joint_attention_dim + cfg.pooled_projection_dim + l_img * 64) * 2
        d2h = B * (8 * h) * (8 * w) * 3
        return h2d, d2h
