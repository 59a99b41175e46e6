"""GPU parity of the text encoders (csrc/text_encoders.cu) against the oracle restatement (oracle/text.py, itself
pinned against HuggingFace transformers in tests/test_text_oracle.py)."""
import pytest
import torch

from oracle import ops as O
from oracle import text as T

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float().cpu() - b).norm() / b.norm()).item()


@pytest.mark.parametrize("B,L", [(1, 40), (2, 77), (1, 512)])
def test_t5_encoder_matches_oracle(fluxlib, B, L):
    from diffusion_rs_b200.text_encoders import T5Config, T5EncoderModel
    cfg = T.T5Config(vocab_size=500, d_model=256, d_kv=64, d_ff=512, num_layers=3, num_heads=4)
    w = T.t5_make_weights(cfg)
    model = T5EncoderModel.new(T5Config(**cfg.__dict__), {k: v.cuda() for k, v in w.items()})
    ids = torch.randint(0, cfg.vocab_size, (B, L), generator=torch.Generator().manual_seed(3))
    ids[:, L - L // 4:] = 0  # zero padding as in FluxPipeline::tokenize_and_pad
    y = model.forward(ids)
    torch.cuda.synchronize()
    ref = T.T5Oracle(cfg, w, O.REF).forward(ids)
    f32 = T.T5Oracle(cfg, w, O.F32).forward(ids)
    assert torch.isfinite(y.float()).all()
    # relative L2 against the bf16-semantics oracle; and no further from the f32 truth than that oracle is (x1.5)
    assert _rel(y, ref) < 1.5e-2, _rel(y, ref)
    assert _rel(y, f32) < 1.5 * _rel(ref, f32) + 1e-3, (_rel(y, f32), _rel(ref, f32))


def test_t5_full_width_layer(fluxlib):
    """One layer at the real T5-XXL widths (d_model 4096, 64 heads, d_ff 10240), L = 128."""
    from diffusion_rs_b200.text_encoders import T5Config, T5EncoderModel
    cfg = T.T5Config(vocab_size=1000, num_layers=1)
    w = T.t5_make_weights(cfg)
    model = T5EncoderModel.new(T5Config(**cfg.__dict__), {k: v.cuda() for k, v in w.items()})
    ids = torch.randint(0, cfg.vocab_size, (1, 128), generator=torch.Generator().manual_seed(4))
    y = model.forward(ids)
    torch.cuda.synchronize()
    ref = T.T5Oracle(cfg, w, O.REF).forward(ids)
    assert _rel(y, ref) < 1.5e-2, _rel(y, ref)


@pytest.mark.parametrize("B,L", [(1, 12), (3, 77)])
def test_clip_text_matches_oracle(fluxlib, B, L):
    from diffusion_rs_b200.text_encoders import ClipTextConfig, ClipTextTransformer
    cfg = T.ClipConfig(vocab_size=600, projection_dim=256, intermediate_size=512, max_position_embeddings=77,
                       num_hidden_layers=3, num_attention_heads=4)
    w = T.clip_make_weights(cfg)
    model = ClipTextTransformer.new(ClipTextConfig(**cfg.__dict__), {k: v.cuda() for k, v in w.items()})
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(3, cfg.vocab_size - 1, (B, L), generator=g)
    for b in range(B):
        ids[b, (5 + 3 * b) % L] = cfg.vocab_size - 1  # EOS = largest id
    hidden = model.forward_with_mask(ids)
    pooled = model.forward(ids)
    torch.cuda.synchronize()
    orc = T.ClipOracle(cfg, w, O.REF)
    ref_h, ref_p = orc.hidden(ids), orc.forward(ids)
    f32_p = T.ClipOracle(cfg, w, O.F32).forward(ids)
    assert _rel(hidden, ref_h) < 1.5e-2, _rel(hidden, ref_h)
    assert _rel(pooled, ref_p) < 1.5e-2, _rel(pooled, ref_p)
    assert _rel(pooled, f32_p) < 1.5 * _rel(ref_p, f32_p) + 1e-3
    # pooling picks the EOS row exactly
    idx = ids.argmax(-1)
    assert torch.equal(pooled.cpu(), hidden.cpu()[torch.arange(B), idx])


def test_clip_full_width(fluxlib):
    """The real CLIP-L text tower widths (768 / 3072, 12 heads), 2 layers, L = 77."""
    from diffusion_rs_b200.text_encoders import ClipTextConfig, ClipTextTransformer
    cfg = T.ClipConfig(vocab_size=2000, num_hidden_layers=2)
    w = T.clip_make_weights(cfg)
    model = ClipTextTransformer.new(ClipTextConfig(**cfg.__dict__), {k: v.cuda() for k, v in w.items()})
    ids = torch.randint(3, cfg.vocab_size - 1, (2, 77), generator=torch.Generator().manual_seed(6))
    ids[:, 20] = cfg.vocab_size - 1
    pooled = model.forward(ids)
    torch.cuda.synchronize()
    ref = T.ClipOracle(cfg, w, O.REF).forward(ids)
    assert _rel(pooled, ref) < 1.5e-2, _rel(pooled, ref)


def test_text_encoder_errors(fluxlib):
    from diffusion_rs_b200 import lib as L
    from diffusion_rs_b200.text_encoders import T5Config, T5EncoderModel
    with pytest.raises(L.Fluxb200Error):
        T5EncoderModel(T5Config(d_kv=32))  # unsupported head dim is reported, not silently mis-computed
    cfg = T.T5Config(vocab_size=50, d_model=128, d_kv=64, d_ff=256, num_layers=1, num_heads=2)
    w = T.t5_make_weights(cfg)
    m = T5EncoderModel(T5Config(**cfg.__dict__))
    for k, v in w.items():
        if "wo.weight" not in k:
            m.load_weight(k, v.cuda())
    with pytest.raises(L.Fluxb200Error, match="missing tensor"):
        m.finalize()


def test_pipeline_prompt_tokens_end_to_end(fluxlib):
    """Pipeline.forward with PromptTokens (ids -> T5 + CLIP on the GPU -> denoise -> VAE) gives the same image as feeding
    the encoders' outputs as PromptEmbeds, and the T5 padding rule of tokenize_and_pad (pad with 0 to the longest)."""
    from diffusion_rs_b200.pipeline import (DiffusionGenerationParams, ModelSource, Pipeline, PromptEmbeds,
                                            PromptTokens)
    from diffusion_rs_b200.text_encoders import ClipTextConfig, T5Config
    from oracle import flux as OF
    from oracle import vae as OV
    fcfg = OF.FluxConfig(num_layers=1, num_single_layers=1, guidance_embeds=True)
    tcfg = T.T5Config(vocab_size=300, d_model=4096, d_kv=64, d_ff=256, num_layers=1, num_heads=2)
    ccfg = T.ClipConfig(vocab_size=400, projection_dim=768, intermediate_size=256, max_position_embeddings=77,
                        num_hidden_layers=1, num_attention_heads=12)
    src = ModelSource.tensors("black-forest-labs/FLUX.1-dev", OF.make_weights(fcfg), OV.make_weights(OV.VaeConfig()))
    src.num_layers, src.num_single_layers = 1, 1
    src.t5, src.clip = T.t5_make_weights(tcfg), T.clip_make_weights(ccfg)
    src.t5_config, src.clip_config = T5Config(**tcfg.__dict__), ClipTextConfig(**ccfg.__dict__)
    pipe = Pipeline.load(src)
    g = torch.Generator().manual_seed(9)
    toks = [PromptTokens(torch.randint(1, 300, (20,), generator=g), torch.randint(1, 399, (9,), generator=g)),
            PromptTokens(torch.randint(1, 300, (32,), generator=g), torch.randint(1, 399, (14,), generator=g))]
    params = DiffusionGenerationParams(height=64, width=64, num_steps=2, guidance_scale=3.5)
    a = pipe.forward(toks, params)
    emb = pipe.encode_prompts(toks)
    assert emb[0].txt.shape == (32, 4096) and emb[0].vec.shape == (768,)  # padded to the longest prompt of the batch
    b = pipe.forward([PromptEmbeds(e.txt.cpu(), e.vec.cpu()) for e in emb], params)
    assert len(a) == 2 and a[0].shape == (64, 64, 3)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # the T5 output itself matches the oracle on the padded ids
    ids = torch.zeros(2, 32, dtype=torch.int64)
    ids[0, :20], ids[1] = toks[0].t5_ids, toks[1].t5_ids
    ref = T.T5Oracle(tcfg, src.t5, O.REF).forward(ids)
    assert _rel(torch.stack([e.txt for e in emb]), ref) < 1.5e-2
