"""CPU: pin the T5 / CLIP oracle restatement (oracle/text.py) against the independent HuggingFace implementations
the reference cites as its source (t5/mod.rs:4), on random small configs with the oracle's own weights."""
import pytest
import torch

from oracle import ops as O
from oracle import text as T

transformers = pytest.importorskip("transformers")


def test_t5_bucket_function_matches_hf():
    from transformers.models.t5.modeling_t5 import T5Attention
    L = 70
    ctx = torch.arange(L)[:, None]
    mem = torch.arange(L)[None, :]
    hf = T5Attention._relative_position_bucket(mem - ctx, bidirectional=True, num_buckets=32, max_distance=128)
    ours = T.t5_relative_buckets(L, L, 32, 128)
    assert torch.equal(hf, ours)


def test_t5_oracle_f32_matches_hf():
    cfg = T.T5Config(vocab_size=100, d_model=64, d_kv=16, d_ff=96, num_layers=2, num_heads=4)
    w = T.t5_make_weights(cfg)
    hf_cfg = transformers.T5Config(vocab_size=cfg.vocab_size, d_model=cfg.d_model, d_kv=cfg.d_kv, d_ff=cfg.d_ff,
                                   num_layers=cfg.num_layers, num_heads=cfg.num_heads,
                                   relative_attention_num_buckets=32, relative_attention_max_distance=128,
                                   layer_norm_epsilon=cfg.layer_norm_epsilon, feed_forward_proj="gated-gelu",
                                   dropout_rate=0.0)
    model = transformers.T5EncoderModel(hf_cfg).eval()
    sd = {k: v.float() for k, v in w.items()}
    sd["encoder.embed_tokens.weight"] = sd["shared.weight"]
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("embed_tokens" in m or "shared" in m for m in missing), missing
    ids = torch.randint(0, cfg.vocab_size, (2, 40), generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ref = model(input_ids=ids).last_hidden_state
    got = T.T5Oracle(cfg, w, O.F32).forward(ids)
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4), (got - ref).abs().max()
    # the bf16-rounding mode stays close to the f32 graph
    got_ref = T.T5Oracle(cfg, w, O.REF).forward(ids)
    assert ((got_ref - ref).norm() / ref.norm()).item() < 2e-2


def test_clip_oracle_f32_matches_hf():
    cfg = T.ClipConfig(vocab_size=120, projection_dim=64, intermediate_size=128, max_position_embeddings=20,
                       num_hidden_layers=2, num_attention_heads=4)
    w = T.clip_make_weights(cfg)
    hf_cfg = transformers.CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.projection_dim,
                                         intermediate_size=cfg.intermediate_size, projection_dim=cfg.projection_dim,
                                         num_hidden_layers=cfg.num_hidden_layers,
                                         num_attention_heads=cfg.num_attention_heads,
                                         max_position_embeddings=cfg.max_position_embeddings, hidden_act="quick_gelu",
                                         layer_norm_eps=1e-5, eos_token_id=2, attention_dropout=0.0)
    model = transformers.CLIPTextModel(hf_cfg).eval()
    sd = {"text_model." + k: v.float() for k, v in w.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in m for m in missing), missing
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(3, cfg.vocab_size - 1, (2, 12), generator=g)
    ids[0, 7] = cfg.vocab_size - 1  # EOS = largest id, as in the CLIP vocabulary
    ids[1, 11] = cfg.vocab_size - 1
    with torch.no_grad():
        out = model(input_ids=ids)
    orc = T.ClipOracle(cfg, w, O.F32)
    assert torch.allclose(orc.hidden(ids), out.last_hidden_state, rtol=1e-4, atol=1e-4)
    assert torch.allclose(orc.forward(ids), out.pooler_output, rtol=1e-4, atol=1e-4)
    ref_mode = T.ClipOracle(cfg, w, O.REF).forward(ids)
    assert ((ref_mode - out.pooler_output).norm() / out.pooler_output.norm()).item() < 2e-2
