"""CPU-side tests: the C-ABI library loads and exports every symbol include/fluxb200.h declares; host logic of the
pipeline mirror; world_size-2 gloo test of the prompt sharding + load-time weight broadcast.  No GPU compute here."""
import os
import re
import socket
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    hdr = (ROOT / "include" / "fluxb200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b((?:fluxb200|dequantize)_[a-z0-9_]+)\s*\(", hdr)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(fluxlib):
    from diffusion_rs_b200 import lib as L
    syms = _declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(fluxlib, s), f"{s} declared in include/fluxb200.h but not exported"
        assert s in L.SIGNATURES, f"{s} has no ctypes signature"
    for s in L.SIGNATURES:
        assert s in syms, f"{s} bound in lib.py but not declared in the header"
    # the reference's own 12 FFI symbols (bitsandbytes/ffi.rs:5-114) must be among them
    ffi = [f"dequantize_blockwise_{t}_{q}" for t in ("f32", "f16", "bf16") for q in ("int8", "fp4", "nf4")]
    ffi += [f"dequantize_8bit_kernel_{t}" for t in ("f32", "f16", "bf16")]
    assert all(s in syms for s in ffi)


def test_header_is_plain_c_and_links_from_c(tmp_path, fluxlib):
    """The boundary is a C ABI: include/fluxb200.h must compile as C99 (no C++ or torch types in the signatures) and a
    C program must link against libfluxb200.so and get an error string back (no GPU needed for that)."""
    import shutil
    import subprocess
    from diffusion_rs_b200 import lib as L
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "fluxb200.h"\n'
                   'int main(void) {\n'
                   '  int rc = fluxb200_set_flag("no_such_flag", 1);\n'
                   '  printf("%d|%s|%d\\n", rc, fluxb200_last_error(), fluxb200_version());\n'
                   '  return rc == 0;\n}\n')
    exe = tmp_path / "t"
    libdir = L.lib_path().parent
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(src), "-o", str(exe),
                    f"-L{libdir}", "-lfluxb200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip()
    rc, msg, ver = out.split("|")
    assert int(rc) != 0 and "unknown flag" in msg and int(ver) >= 100


def test_no_cpu_fallback_and_error_reporting(fluxlib):
    """Without a GPU every compute entry point must fail loudly with a message, never silently compute on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ctypes as C
    from diffusion_rs_b200 import lib as L
    cfg = L.FluxConfigC(64, 768, 4096, 24, 1, 1, 1)
    h = C.c_void_p()
    rc = fluxlib.fluxb200_model_create(C.byref(cfg), C.byref(h))
    assert rc != 0
    assert len(fluxlib.fluxb200_last_error()) > 0
    from diffusion_rs_b200.pipeline import ModelSource, Pipeline
    with pytest.raises(L.Fluxb200Error):
        Pipeline.load(ModelSource.synthetic())


def test_product_code_never_imports_the_oracle():
    for p in list((ROOT / "diffusion_rs_b200").rglob("*.py")):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{p} imports the oracle"


def test_scheduler_matches_oracle():
    from diffusion_rs_b200 import pipeline as P
    from oracle import flux as OF
    for l_img in (256, 1024, 3600, 4096):
        sc = P.SchedulerConfig()
        mu = P.calculate_shift(l_img, sc.base_image_seq_len, sc.max_image_seq_len, sc.base_shift, sc.max_shift)
        assert mu == OF.calculate_shift(l_img)
        for n in (1, 4, 50):
            assert sc.get_timesteps(n, mu) == OF.get_timesteps(n, mu)
    sch = P.SchedulerConfig(use_dynamic_shifting=False, shift=1.0)
    assert sch.get_timesteps(4, None) == OF.get_timesteps(4, None, shift=1.0, dynamic=False) == [1.0, 0.75, 0.5, 0.25, 0.0]
    with pytest.raises(Exception):
        P.SchedulerConfig().get_timesteps(4, None)


def test_shift_call_site_mirrors_the_reference():
    """FluxPipeline::forward passes `img.dims()[1]` of the UNPACKED noise [bs,16,h,w] to calculate_shift
    (pipelines/flux/mod.rs:276-285): image_seq_len = 16 at every resolution.  The default mirrors that; "bfl" is the
    packed sequence length upstream FLUX uses."""
    from diffusion_rs_b200 import lib as L
    from diffusion_rs_b200 import pipeline as P
    sc = P.SchedulerConfig()
    for hw in ((128, 128), (90, 160), (32, 32)):
        shape = (3, 16) + hw
        assert P.shift_seq_len(shape) == P.shift_seq_len(shape, "reference") == 16
        assert P.shift_seq_len(shape, "bfl") == (hw[0] // 2) * (hw[1] // 2)
    mu = P.calculate_shift(16, sc.base_image_seq_len, sc.max_image_seq_len, sc.base_shift, sc.max_shift)
    assert abs(mu - (0.5 + (16 - 256) * (1.15 - 0.5) / (4096 - 256))) < 1e-12 and abs(mu - 0.459375) < 1e-9
    ts = sc.get_timesteps(50, mu)
    assert ts[0] == 1.0 and ts[-1] == 0.0 and all(a > b for a, b in zip(ts, ts[1:]))
    with pytest.raises(L.Fluxb200Error):
        P.shift_seq_len((1, 16, 8, 8), "nope")


def test_vae_config_rejects_quant_convs():
    from diffusion_rs_b200 import ingest
    from diffusion_rs_b200 import lib as L
    base = dict(latent_channels=16, out_channels=3, block_out_channels=[128, 256, 512, 512], layers_per_block=2,
                norm_num_groups=32, scaling_factor=0.3611, shift_factor=0.1159)
    assert ingest.vae_config_from_json(dict(base, use_post_quant_conv=False)).latent_channels == 16
    with pytest.raises(L.Fluxb200Error):
        ingest.vae_config_from_json(dict(base, use_post_quant_conv=True))


def test_patchify_ids_latent_geometry():
    from diffusion_rs_b200 import pipeline as P
    from oracle import flux as OF
    assert P.latent_hw(1024, 1024) == (128, 128) and P.latent_hw(720, 1280) == (90, 160) and P.latent_hw(250, 250) == (32, 32)
    lat = torch.randn(2, 16, 6, 8)
    assert torch.equal(P.patchify(lat), OF.patchify(lat))
    img_ids, txt_ids = P.make_ids(5, 7, 11, torch.float32)
    assert torch.equal(torch.cat([txt_ids, img_ids]), OF.make_ids(5, 7, 11))


def test_synthetic_checkpoint_names_match_oracle_and_quantisers_agree():
    from diffusion_rs_b200 import quantize as QZ
    from diffusion_rs_b200 import synthetic as S
    from diffusion_rs_b200.transformer import FluxConfig
    from diffusion_rs_b200.vae import VaeConfig
    from oracle import flux as OF
    from oracle import quant as Q
    from oracle import vae as OV
    cfg = FluxConfig(num_layers=1, num_single_layers=1)
    ocfg = OF.FluxConfig(num_layers=1, num_single_layers=1)
    assert S.flux_linear_shapes(cfg) == OF.linear_shapes(ocfg)
    names = {n: tuple(t.shape) for n, t in S.iter_vae_tensors(VaeConfig(), device="cpu")}
    assert names == {n: tuple(s) for n, (s, _) in OV.weight_specs(OV.VaeConfig()).items()}
    w = torch.randn(64, 512).bfloat16()
    assert np.array_equal(QZ.quantize_q4k(w).numpy(), Q.quantize_q4k(w.float().numpy()).reshape(-1))
    packed, a8, code, nmax, off, lut = QZ.quantize_nf4(w)
    p2, am2 = Q.quantize_4bit(w.float().numpy(), 64, "nf4")
    assert np.array_equal(packed.numpy(), p2)
    a8b, codeb, nmaxb, offb = Q.quantize_absmax_nested(am2, 256)
    assert np.array_equal(a8.numpy(), a8b) and np.allclose(nmax.numpy(), nmaxb) and off == offb
    out = dict((n, t) for n, t, _, _ in QZ.quantize_tensor("transformer_blocks.0.attn.to_q.weight", w, "nf4"))
    assert set(k.split("weight")[-1] for k in out) == {"", ".absmax", ".quant_map", ".nested_absmax", ".nested_quant_map",
                                                       ".quant_state.bitsandbytes__nf4"}
    assert [n for n, *_ in QZ.quantize_tensor("x_embedder.weight", w, "nf4")] == ["x_embedder.weight"]


def test_shard_range_covers_batch():
    from diffusion_rs_b200.pipeline import shard_range
    for n in (0, 1, 7, 8, 32, 33):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
    assert shard_range(8, 3, 8) == (3, 4) and shard_range(32, 7, 8) == (28, 32)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffusion_rs_b200.pipeline import broadcast_weight, dist_info, shard_range
    _, r, w = dist_info()
    t = torch.full((1000,), float(rank + 1))
    broadcast_weight(t)  # load-time weight broadcast: every rank ends with rank 0's tensor
    lo, hi = shard_range(5, r, w)
    got = [None] * world
    dist.all_gather_object(got, (lo, hi, float(t.sum())))
    if rank == 0:
        q.put(got)
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_weight_broadcast():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [(0, 3, 1000.0), (3, 5, 1000.0)]


def test_tokenizers_mirror_the_reference_construction():
    """T5: Tokenizer::from_bytes(tokenizer_2/tokenizer.json); CLIP: a BARE BPE model from vocab.json + merges.txt
    (diffusion_rs_common/src/tokenizer.rs:7-23) - no lower-casing, no BOS/EOS, first merges line skipped."""
    tokenizers = pytest.importorskip("tokenizers")
    import json
    from tokenizers import Tokenizer, models, pre_tokenizers, processors
    from diffusion_rs_b200.pipeline import Pipeline, PromptTokens
    t5 = Tokenizer(models.WordLevel({"<pad>": 0, "</s>": 1, "<unk>": 2, "a": 3, "cat": 4}, unk_token="<unk>"))
    t5.pre_tokenizer = pre_tokenizers.Whitespace()
    t5.post_processor = processors.TemplateProcessing(single="$A </s>", special_tokens=[("</s>", 1)])
    vocab = {"c": 0, "a": 1, "t": 2, " ": 3, "ca": 4, "cat": 5}
    merges = "#version: 0.2\nc a\nca t\n"
    toks = Pipeline._build_tokenizers({"t5": t5.to_str().encode(), "clip_vocab": json.dumps(vocab).encode(),
                                       "clip_merges": merges.encode()})
    pipe = Pipeline.__new__(Pipeline)  # host-side logic only: no GPU objects needed
    pipe.tokenizers = toks
    out = pipe.tokenize("a cat")
    assert isinstance(out, PromptTokens)
    assert out.t5_ids.tolist() == [3, 4, 1]          # words + the </s> the T5 post-processor appends
    assert out.clip_ids.tolist() == [1, 3, 5]        # bare BPE: 'a', ' ', 'cat' - no BOS/EOS, no </w> handling
    ref = Tokenizer(models.BPE(vocab=vocab, merges=[("c", "a"), ("ca", "t")]))
    assert out.clip_ids.tolist() == ref.encode("a cat", add_special_tokens=True).ids
    pipe.tokenizers = None
    from diffusion_rs_b200 import lib as L
    with pytest.raises(L.Fluxb200Error):
        pipe.tokenize("a cat")


@pytest.mark.gpu
def test_two_devices_driven_from_two_threads(fluxlib):
    """SURVEY §8(b) threading row: handles are re-entrant per (device, stream); per-device kernel attributes (opt-in
    shared memory) are set on every device the process uses, not once per process.  Two models on two GPUs, driven
    concurrently from two host threads, each equal to its own single-threaded result."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs in one process")
    import threading
    from diffusion_rs_b200.transformer import FluxConfig, FluxTransformer
    from oracle import flux as OF
    cfg = OF.FluxConfig(num_layers=1, num_single_layers=1, guidance_embeds=True)
    weights = OF.make_weights(cfg)
    B, h2, w2, l_txt = 1, 12, 12, 64
    g = torch.Generator().manual_seed(5)
    img = torch.randn(B, h2 * w2, 64, generator=g).bfloat16()
    txt = torch.randn(B, l_txt, 4096, generator=g).bfloat16()
    y = torch.randn(B, 768, generator=g).bfloat16()
    ids = OF.make_ids(h2, w2, l_txt).bfloat16()
    ts = OF.get_timesteps(3, OF.calculate_shift(16))

    def run(dev, out, reps):
        torch.cuda.set_device(dev)
        d = f"cuda:{dev}"
        m = FluxTransformer.new(FluxConfig(num_layers=1, num_single_layers=1, guidance_embeds=True),
                                {k: v.to(d) for k, v in weights.items()})
        with torch.cuda.stream(torch.cuda.Stream(device=d)):
            for _ in range(reps):
                x = img.to(d).clone()
                m.denoise(x, ids[l_txt:][None].contiguous().to(d), txt.to(d), ids[:l_txt][None].contiguous().to(d),
                          y.to(d), 3.5, ts)
            torch.cuda.current_stream().synchronize()
        out[dev] = x.cpu()

    solo, both = {}, {}
    for dev in (0, 1):
        run(dev, solo, 1)
    threads = [threading.Thread(target=run, args=(dev, both, 3)) for dev in (0, 1)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.set_device(0)
    assert set(both) == {0, 1}, "a worker thread failed"
    for dev in (0, 1):
        assert torch.equal(solo[dev], both[dev])
    assert torch.equal(solo[0], solo[1])


def test_reference_arm_line_has_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line, exactly K timed samples,
    the same metric / unit / config builder as our arm, `cpu_baseline` and a zero-copy `e2e` object."""
    import json
    import subprocess
    import sys
    sys.path.insert(0, str(ROOT))
    import bench
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--height", "256", "--width", "256", "--num-steps", "4"],
                         capture_output=True, text=True, timeout=600, check=True).stdout.strip().splitlines()[-1]
    j = json.loads(out)
    assert j["impl"] == "reference" and j["metric"] == bench.METRIC and j["unit"] == bench.UNIT
    assert j["steps"] == 1 and j["higher_is_better"] is True and j["value"] > 0 and j["ms_per_step"] > 0
    ours = bench.workload_config(256, 256, 4, 1, 1, None)
    assert {k: j["config"][k] for k in ours} == ours  # same config object as our arm, plus the reference-arm notes
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "sample" in cb
    assert j["e2e"] == {"value": j["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the reported per-step time is the measured sample, not an extrapolated image (it must fit the driver's clock)
    assert j["ms_per_step"] / 1e3 < j["wall_s"]
