"""Weight ingest: local diffusers directory / DDUF archive -> tensors (CPU part), and a GPU round trip."""
import json
import zipfile

import pytest
import torch
from safetensors.torch import save_file

from diffusion_rs_b200 import ingest
from diffusion_rs_b200 import lib as L

TCFG = {"in_channels": 64, "pooled_projection_dim": 768, "joint_attention_dim": 4096, "num_attention_heads": 24,
        "num_layers": 1, "num_single_layers": 1, "guidance_embeds": True}
VCFG = {"latent_channels": 16, "out_channels": 3, "block_out_channels": [128, 256, 512, 512], "layers_per_block": 2,
        "norm_num_groups": 32, "scaling_factor": 0.3611, "shift_factor": 0.1159, "mid_block_add_attention": True,
        "in_channels": 3, "down_block_types": ["DownEncoderBlock2D"] * 4, "up_block_types": ["UpDecoderBlock2D"] * 4}
SCFG = {"_class_name": "FlowMatchEulerDiscreteScheduler", "base_image_seq_len": 256, "base_shift": 0.5,
        "max_image_seq_len": 4096, "max_shift": 1.15, "shift": 3.0, "use_dynamic_shifting": True}


def _write_model_dir(root, transformer, vae, class_name="FluxPipeline"):
    (root / "transformer").mkdir(parents=True)
    (root / "vae").mkdir()
    (root / "scheduler").mkdir()
    (root / "model_index.json").write_text(json.dumps({"_class_name": class_name}))
    (root / "transformer" / "config.json").write_text(json.dumps(TCFG))
    (root / "vae" / "config.json").write_text(json.dumps(VCFG))
    (root / "scheduler" / "scheduler_config.json").write_text(json.dumps(SCFG))
    names = sorted(transformer)
    half = len(names) // 2  # two shards, like a real checkpoint
    save_file({k: transformer[k] for k in names[:half]}, str(root / "transformer" / "model-00001-of-00002.safetensors"))
    save_file({k: transformer[k] for k in names[half:]}, str(root / "transformer" / "model-00002-of-00002.safetensors"))
    save_file(vae, str(root / "vae" / "diffusion_pytorch_model.safetensors"))


def _tiny_tensors():
    tr = {"x_embedder.weight": torch.randn(8, 4), "x_embedder.bias": torch.randn(8).half(),
          "transformer_blocks.0.attn.to_q.weight": torch.randint(0, 255, (16, 1), dtype=torch.uint8),
          "transformer_blocks.0.attn.to_q.weight.absmax": torch.randint(0, 255, (4,), dtype=torch.uint8),
          "transformer_blocks.0.attn.to_q.weight.quant_map": torch.randn(16),
          "transformer_blocks.0.attn.to_q.weight.nested_absmax": torch.randn(1),
          "transformer_blocks.0.attn.to_q.weight.quant_state.bitsandbytes__nf4": torch.tensor(list(b"{}"), dtype=torch.uint8),
          "transformer_blocks.0.attn.to_k.SCB": torch.randn(8)}
    va = {"decoder.conv_in.weight": torch.randn(4, 2, 3, 3), "encoder.conv_in.weight": torch.randn(4, 2, 3, 3)}
    return tr, va


@pytest.mark.parametrize("kind", ["dir", "dduf"])
def test_ingest_layout_and_dtype_policy(tmp_path, kind):
    tr, va = _tiny_tensors()
    root = tmp_path / "model"
    _write_model_dir(root, tr, va)
    if kind == "dduf":
        z = tmp_path / "model.dduf"
        with zipfile.ZipFile(z, "w", zipfile.ZIP_STORED) as zf:
            for p in root.rglob("*"):
                if p.is_file():
                    zf.write(p, str(p.relative_to(root)))
        loader = ingest.open_source("dduf", str(z))
    else:
        loader = ingest.open_source("model_id", str(root))
    tj, t_tr, vj, t_va, sj = ingest.load_flux_components(loader)
    assert ingest.flux_config_from_json(tj).num_layers == 1 and ingest.flux_config_from_json(tj).guidance_embeds
    assert ingest.vae_config_from_json(vj).block_out_channels == (128, 256, 512, 512)
    assert ingest.scheduler_config_from_json(sj).use_dynamic_shifting
    assert set(t_tr) == set(tr)
    assert t_tr["x_embedder.weight"].dtype == torch.bfloat16 and t_tr["x_embedder.bias"].dtype == torch.bfloat16
    assert t_tr["transformer_blocks.0.attn.to_q.weight"].dtype == torch.uint8
    assert t_tr["transformer_blocks.0.attn.to_q.weight.absmax"].dtype == torch.uint8
    assert t_tr["transformer_blocks.0.attn.to_q.weight.quant_map"].dtype == torch.float32
    assert t_tr["transformer_blocks.0.attn.to_q.weight.nested_absmax"].dtype == torch.float32
    assert t_tr["transformer_blocks.0.attn.to_k.SCB"].dtype == torch.float32
    assert set(t_va) == {"decoder.conv_in.weight"}  # the pipeline never encodes


def test_ingest_text_components(tmp_path):
    """text_encoder (CLIP, `text_model.` prefix stripped), text_encoder_2 (T5) and the tokenizer files of a snapshot."""
    tr, va = _tiny_tensors()
    root = tmp_path / "model"
    _write_model_dir(root, tr, va)
    loader = ingest.open_source("model_id", str(root))
    assert ingest.load_text_components(loader) is None  # a snapshot without text encoders is fine
    for d in ("text_encoder", "text_encoder_2", "tokenizer", "tokenizer_2"):
        (root / d).mkdir()
    ccfg = {"vocab_size": 100, "projection_dim": 64, "intermediate_size": 128, "max_position_embeddings": 77,
            "num_hidden_layers": 1, "num_attention_heads": 1, "hidden_act": "quick_gelu"}
    t5cfg = {"vocab_size": 100, "d_model": 64, "d_kv": 64, "d_ff": 256, "num_layers": 1, "num_heads": 1,
             "relative_attention_num_buckets": 32, "layer_norm_epsilon": 1e-6, "feed_forward_proj": "gated-gelu"}
    (root / "text_encoder" / "config.json").write_text(json.dumps(ccfg))
    (root / "text_encoder_2" / "config.json").write_text(json.dumps(t5cfg))
    save_file({"text_model.final_layer_norm.weight": torch.randn(64), "text_model.embeddings.position_ids":
               torch.arange(77)[None].float(), "logit_scale": torch.randn(1)},
              str(root / "text_encoder" / "model.safetensors"))
    save_file({"shared.weight": torch.randn(100, 64), "encoder.final_layer_norm.weight": torch.randn(64).half(),
               "decoder.junk": torch.randn(2)}, str(root / "text_encoder_2" / "model.safetensors"))
    c, clip, t, t5, toks = ingest.load_text_components(ingest.open_source("model_id", str(root)))
    assert set(clip) == {"final_layer_norm.weight"} and clip["final_layer_norm.weight"].dtype == torch.bfloat16
    assert set(t5) == {"shared.weight", "encoder.final_layer_norm.weight"}
    assert toks is None  # no tokenizer files yet
    assert ingest.clip_config_from_json(c).projection_dim == 64
    assert ingest.t5_config_from_json(t).relative_attention_max_distance == 128
    with pytest.raises(L.Fluxb200Error):
        ingest.t5_config_from_json({**t5cfg, "feed_forward_proj": "relu"})


def test_ingest_rejects_other_pipelines(tmp_path):
    tr, va = _tiny_tensors()
    root = tmp_path / "sd"
    _write_model_dir(root, tr, va, class_name="StableDiffusionPipeline")
    with pytest.raises(L.Fluxb200Error):
        ingest.load_flux_components(ingest.open_source("model_id", str(root)))
    with pytest.raises(L.Fluxb200Error):
        ingest.open_source("model_id", str(tmp_path / "does-not-exist"))


@pytest.mark.gpu
def test_pipeline_load_from_local_directory(tmp_path, fluxlib):
    """Pipeline.load(ModelSource.from_model_id(dir)) == loading the same tensors directly."""
    from diffusion_rs_b200.pipeline import DiffusionGenerationParams, ModelSource, Pipeline
    from oracle import flux as OF
    from oracle import vae as OV
    cfg = OF.FluxConfig(num_layers=1, num_single_layers=1, guidance_embeds=True)
    tr = OF.make_weights(cfg)
    va = OV.make_weights(OV.VaeConfig())
    root = tmp_path / "flux"
    _write_model_dir(root, tr, va)
    p1 = Pipeline.load(ModelSource.from_model_id(str(root)))
    src = ModelSource.tensors("black-forest-labs/FLUX.1-dev", tr, va)
    src.num_layers, src.num_single_layers = 1, 1
    p2 = Pipeline.load(src)
    params = DiffusionGenerationParams(height=64, width=96, num_steps=2, guidance_scale=3.5)
    emb = [p1.synthetic_embeds("a cat")]
    a = p1.forward(emb, params)
    b = p2.forward(emb, params)
    assert a[0].shape == (64, 96, 3) and a[0].dtype == torch.uint8
    assert torch.equal(a[0], b[0])
    # a real snapshot without tokenizers / text encoders must refuse a text prompt instead of inventing embeddings
    with pytest.raises(L.Fluxb200Error):
        p1.forward(["a cat"], params)


@pytest.mark.gpu
def test_pipeline_from_snapshot_with_text_encoders_and_tokenizers(tmp_path, fluxlib):
    """A local snapshot that ships text_encoder / text_encoder_2 / tokenizer / tokenizer_2: string prompts are tokenised
    like the reference does, encoded by T5 + CLIP on the GPU and turned into images; the result equals feeding the same
    token ids explicitly."""
    pytest.importorskip("tokenizers")
    from tokenizers import Tokenizer, models, pre_tokenizers, processors
    from diffusion_rs_b200.pipeline import DiffusionGenerationParams, ModelSource, Pipeline, PromptTokens
    from oracle import flux as OF
    from oracle import text as OT
    from oracle import vae as OV
    cfg = OF.FluxConfig(num_layers=1, num_single_layers=1, guidance_embeds=True)
    root = tmp_path / "flux"
    _write_model_dir(root, OF.make_weights(cfg), OV.make_weights(OV.VaeConfig()))
    for d in ("text_encoder", "text_encoder_2", "tokenizer", "tokenizer_2"):
        (root / d).mkdir()
    ccfg = OT.ClipConfig(vocab_size=16, projection_dim=768, intermediate_size=256, max_position_embeddings=77,
                         num_hidden_layers=1, num_attention_heads=12)
    tcfg = OT.T5Config(vocab_size=16, d_model=4096, d_kv=64, d_ff=256, num_layers=1, num_heads=2)
    (root / "text_encoder" / "config.json").write_text(json.dumps({**ccfg.__dict__, "hidden_act": "quick_gelu"}))
    (root / "text_encoder_2" / "config.json").write_text(json.dumps({**tcfg.__dict__, "feed_forward_proj": "gated-gelu"}))
    save_file({"text_model." + k: v for k, v in OT.clip_make_weights(ccfg).items()},
              str(root / "text_encoder" / "model.safetensors"))
    save_file(OT.t5_make_weights(tcfg), str(root / "text_encoder_2" / "model.safetensors"))
    t5 = Tokenizer(models.WordLevel({"<pad>": 0, "</s>": 1, "<unk>": 2, "a": 3, "cat": 4, "photo": 5, "of": 6},
                                    unk_token="<unk>"))
    t5.pre_tokenizer = pre_tokenizers.Whitespace()
    t5.post_processor = processors.TemplateProcessing(single="$A </s>", special_tokens=[("</s>", 1)])
    (root / "tokenizer_2" / "tokenizer.json").write_text(t5.to_str())
    vocab = {"c": 0, "a": 1, "t": 2, " ": 3, "ca": 4, "cat": 5, "p": 6, "h": 7, "o": 8, "f": 9}
    (root / "tokenizer" / "vocab.json").write_text(json.dumps(vocab))
    (root / "tokenizer" / "merges.txt").write_text("#version: 0.2\nc a\nca t\n")
    pipe = Pipeline.load(ModelSource.from_model_id(str(root)))
    assert pipe.t5 is not None and pipe.clip is not None and pipe.tokenizers is not None
    params = DiffusionGenerationParams(height=64, width=64, num_steps=2, guidance_scale=3.5)
    a = pipe.forward(["a photo of a cat"], params)
    toks = pipe.tokenize("a photo of a cat")
    assert toks.t5_ids.tolist() == [3, 5, 6, 3, 4, 1]
    b = pipe.forward([PromptTokens(toks.t5_ids, toks.clip_ids)], params)
    assert a[0].shape == (64, 64, 3) and torch.equal(a[0], b[0])
    png = pipe.forward_png(["a photo of a cat"], params)
    assert png[0][:8] == b"\x89PNG\r\n\x1a\n"
