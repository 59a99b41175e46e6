"""CPU: pin the FLUX DiT oracle (oracle/flux.py, a restatement of diffusion-rs' models/flux/model.rs) against an
INDEPENDENT implementation of the same network: the Black Forest Labs reference code as vendored by torchtitan
(`torchtitan.experiments.flux.model`, "imported from black-forest-labs/FLUX").  The reference itself (Rust) cannot be
run here; its model file is a port of exactly that network onto diffusers weight names, so agreement of the oracle's
f32 graph with the BFL graph under the diffusers -> BFL name map pins the structure (modulation chunk order, q|k|v
layout, interleaved RoPE, txt-first joint attention, the (scale, shift) order of the last layer, ...)."""
import pytest
import torch

from oracle import flux as OF
from oracle import ops as O

tt_model = pytest.importorskip("torchtitan.experiments.flux.model.model")
tt_args = pytest.importorskip("torchtitan.experiments.flux.model.args")

D = OF.HIDDEN


def bfl_state_dict(cfg: OF.FluxConfig, w: dict) -> dict:
    """diffusers names (what the reference's VarBuilder resolves) -> BFL module names."""
    f = {k: v.float() for k, v in w.items()}

    def lin(dst, src, out):
        out[dst + ".weight"], out[dst + ".bias"] = f[src + ".weight"], f[src + ".bias"]

    def cat(dst, srcs, out):
        out[dst + ".weight"] = torch.cat([f[s + ".weight"] for s in srcs], 0)
        out[dst + ".bias"] = torch.cat([f[s + ".bias"] for s in srcs], 0)

    sd = {}
    lin("img_in", "x_embedder", sd)
    lin("txt_in", "context_embedder", sd)
    lin("time_in.in_layer", "time_text_embed.timestep_embedder.linear_1", sd)
    lin("time_in.out_layer", "time_text_embed.timestep_embedder.linear_2", sd)
    lin("vector_in.in_layer", "time_text_embed.text_embedder.linear_1", sd)
    lin("vector_in.out_layer", "time_text_embed.text_embedder.linear_2", sd)
    for i in range(cfg.num_layers):
        s, d = f"transformer_blocks.{i}.", f"double_blocks.{i}."
        lin(d + "img_mod.lin", s + "norm1.linear", sd)
        lin(d + "txt_mod.lin", s + "norm1_context.linear", sd)
        cat(d + "img_attn.qkv", [s + "attn.to_q", s + "attn.to_k", s + "attn.to_v"], sd)
        cat(d + "txt_attn.qkv", [s + "attn.add_q_proj", s + "attn.add_k_proj", s + "attn.add_v_proj"], sd)
        sd[d + "img_attn.norm.query_norm.weight"] = f[s + "attn.norm_q.weight"]
        sd[d + "img_attn.norm.key_norm.weight"] = f[s + "attn.norm_k.weight"]
        sd[d + "txt_attn.norm.query_norm.weight"] = f[s + "attn.norm_added_q.weight"]
        sd[d + "txt_attn.norm.key_norm.weight"] = f[s + "attn.norm_added_k.weight"]
        lin(d + "img_attn.proj", s + "attn.to_out.0", sd)
        lin(d + "txt_attn.proj", s + "attn.to_add_out", sd)
        lin(d + "img_mlp.0", s + "ff.net.0.proj", sd)
        lin(d + "img_mlp.2", s + "ff.net.2", sd)
        lin(d + "txt_mlp.0", s + "ff_context.net.0.proj", sd)
        lin(d + "txt_mlp.2", s + "ff_context.net.2", sd)
    for i in range(cfg.num_single_layers):
        s, d = f"single_transformer_blocks.{i}.", f"single_blocks.{i}."
        lin(d + "modulation.lin", s + "norm.linear", sd)
        cat(d + "linear1", [s + "attn.to_q", s + "attn.to_k", s + "attn.to_v", s + "proj_mlp"], sd)
        lin(d + "linear2", s + "proj_out", sd)
        sd[d + "norm.query_norm.weight"] = f[s + "attn.norm_q.weight"]
        sd[d + "norm.key_norm.weight"] = f[s + "attn.norm_k.weight"]
    # diffusers' AdaLayerNormContinuous stores (scale, shift); BFL's LastLayer chunks (shift, scale)
    wt, bs = f["norm_out.linear.weight"], f["norm_out.linear.bias"]
    sd["final_layer.adaLN_modulation.1.weight"] = torch.cat([wt[D:], wt[:D]], 0)
    sd["final_layer.adaLN_modulation.1.bias"] = torch.cat([bs[D:], bs[:D]], 0)
    lin("final_layer.linear", "proj_out", sd)
    return sd


def test_flux_oracle_f32_matches_bfl_reference_implementation():
    torch.manual_seed(0)
    cfg = OF.FluxConfig(num_layers=1, num_single_layers=1, guidance_embeds=False)  # the BFL class has no guidance embedder
    w = OF.make_weights(cfg)
    args = tt_args.FluxModelArgs(context_in_dim=cfg.joint_attention_dim, depth=1, depth_single_blocks=1)
    with torch.device("meta"):
        model = tt_model.FluxModel(args)
    model = model.to_empty(device="cpu").float().eval()
    missing, unexpected = model.load_state_dict(bfl_state_dict(cfg, w), strict=False)
    assert not unexpected, unexpected
    assert not missing, missing
    for m in model.modules():  # the reference's QkNorm uses eps 1e-6 (model.rs:186-209); nn.RMSNorm defaults to finfo.eps
        if isinstance(m, torch.nn.RMSNorm):
            m.eps = 1e-6
    h2, w2, l_txt, B = 4, 6, 8, 2
    g = torch.Generator().manual_seed(11)
    img = torch.randn(B, h2 * w2, 64, generator=g)
    txt = torch.randn(B, l_txt, cfg.joint_attention_dim, generator=g)
    y = torch.randn(B, 768, generator=g)
    t = torch.tensor([0.75, 0.3])
    ids = OF.make_ids(h2, w2, l_txt)  # [L, 3], txt rows first
    with torch.no_grad():
        ref = model(img, ids[l_txt:][None].expand(B, -1, -1), txt, ids[:l_txt][None].expand(B, -1, -1), t, y)
    got = OF.FluxOracle(cfg, w, O.F32).forward(img, ids, txt, t, y, None)
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 2e-5, rel
    # and the bf16-rounding mode of the same oracle stays within bf16 noise of it
    got_ref = OF.FluxOracle(cfg, w, O.REF).forward(O.rb(img), ids, O.rb(txt), t, O.rb(y), None)
    assert ((got_ref - ref).norm() / ref.norm()).item() < 3e-2


def test_vae_decoder_oracle_f32_matches_bfl_reference_implementation():
    """Same idea for AutoEncoderKl::decode (models/vaes/vae.rs): diffusers decoder names -> BFL `Decoder` names
    (up_blocks are listed in application order in diffusers and in reverse in BFL; the mid-block attention's Linear
    weights become 1x1 convolutions; conv_shortcut == nin_shortcut)."""
    tt_ae = pytest.importorskip("torchtitan.experiments.flux.model.autoencoder")
    from oracle import vae as OV
    cfg = OV.VaeConfig()
    w = OV.make_weights(cfg)
    f = {k: v.float() for k, v in w.items()}
    p = tt_ae.AutoEncoderParams()
    dec = tt_ae.Decoder(resolution=p.resolution, in_channels=p.in_channels, ch=p.ch, out_ch=p.out_ch, ch_mult=p.ch_mult,
                        num_res_blocks=p.num_res_blocks, z_channels=p.z_channels).float().eval()
    sd = {}

    def put(dst, src):
        for suf in (".weight", ".bias"):
            sd[dst + suf] = f["decoder." + src + suf]

    def resnet(dst, src):
        for n in ("norm1", "conv1", "norm2", "conv2"):
            put(f"{dst}.{n}", f"{src}.{n}")
        if f"decoder.{src}.conv_shortcut.weight" in f:
            put(f"{dst}.nin_shortcut", f"{src}.conv_shortcut")

    put("conv_in", "conv_in")
    resnet("mid.block_1", "mid_block.resnets.0")
    resnet("mid.block_2", "mid_block.resnets.1")
    put("mid.attn_1.norm", "mid_block.attentions.0.group_norm")
    for dst, src in (("q", "to_q"), ("k", "to_k"), ("v", "to_v"), ("proj_out", "to_out.0")):
        sd[f"mid.attn_1.{dst}.weight"] = f[f"decoder.mid_block.attentions.0.{src}.weight"][:, :, None, None]
        sd[f"mid.attn_1.{dst}.bias"] = f[f"decoder.mid_block.attentions.0.{src}.bias"]
    for lvl in range(4):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"up.{3 - lvl}.block.{j}", f"up_blocks.{lvl}.resnets.{j}")
        if lvl != 3:
            put(f"up.{3 - lvl}.upsample.conv", f"up_blocks.{lvl}.upsamplers.0.conv")
    put("norm_out", "conv_norm_out")
    put("conv_out", "conv_out")
    missing, unexpected = dec.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    z = torch.randn(1, 16, 8, 12, generator=torch.Generator().manual_seed(21))
    with torch.no_grad():
        ref = dec(z)
    got = OV.VaeOracle(cfg, w, O.F32).decode(z)
    assert got.shape == ref.shape == (1, 3, 64, 96)
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 2e-5, rel


def test_sampler_helpers_match_bfl_reference_implementation():
    """Scheduler (`calculate_shift` + `get_timesteps`, pipelines/scheduler.rs:22-51, flux/sampling.rs:70-80), latent
    packing / unpacking (flux/sampling.rs:29-31, 61-68) and position ids (flux/sampling.rs:32-49) — both the oracle's and
    the product's host-side mirrors — against the BFL helpers vendored by torchtitan."""
    tt_s = pytest.importorskip("torchtitan.experiments.flux.sampling")
    tt_u = pytest.importorskip("torchtitan.experiments.flux.utils")
    from diffusion_rs_b200 import pipeline as PL
    for l_img, steps in ((256, 4), (4096, 50), (3600, 28)):
        ref = tt_s.get_schedule(steps, l_img, shift=True)
        mu = OF.calculate_shift(l_img)
        assert mu == PL.calculate_shift(l_img, 256, 4096, 0.5, 1.15)
        for ts in (OF.get_timesteps(steps, mu), PL.SchedulerConfig().get_timesteps(steps, mu)):
            assert len(ts) == steps + 1 and ts[0] == 1.0 and ts[-1] == 0.0
            assert max(abs(a - b) for a, b in zip(ts, ref)) < 1e-6  # BFL evaluates the shift in f32
    # schnell: no shift -> plain linspace(1, 0)
    ref = tt_s.get_schedule(4, 256, shift=False)
    assert max(abs(a - b) for a, b in zip(PL.SchedulerConfig(use_dynamic_shifting=False, shift=1.0).get_timesteps(4, None), ref)) < 1e-7
    lat = torch.randn(2, 16, 8, 12, generator=torch.Generator().manual_seed(5))
    packed = tt_u.pack_latents(lat)
    assert torch.equal(OF.patchify(lat), packed) and torch.equal(PL.patchify(lat), packed)
    assert torch.equal(OF.unpack(packed, 8 * 8, 12 * 8), tt_u.unpack_latents(packed, 8, 12))
    ids_ref = tt_u.create_position_encoding_for_latents(1, 8, 12)[0]
    assert torch.equal(OF.make_ids(4, 6, 5)[5:], ids_ref) and torch.equal(OF.make_ids(4, 6, 5)[:5], torch.zeros(5, 3))
    img_ids, txt_ids = PL.make_ids(4, 6, 5, dtype=torch.float32)
    assert torch.equal(img_ids, ids_ref) and torch.equal(txt_ids, torch.zeros(5, 3))
