#!/usr/bin/env python
"""bench.py — images/sec of the FLUX.1-dev 1024x1024 50-step bf16 hot path (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

A "step" is ONE IMAGE per GPU: 50 DiT denoising steps + the VAE decode (configs[1] of BASELINE.json).  For N > 1 the
driver launches this file under torchrun; prompts are sharded one per rank (weak scaling, no per-step collective).
Timing: W untimed images, then exactly K images bracketed by barrier + synchronize, CUDA events, max over ranks.
  value : device-resident inputs (latents/embeddings already in HBM)            -> images/s, whole job
  e2e   : Pipeline.forward with HOST buffers (pinned H2D of embeddings + noise, D2H of the u8 image inside the
          timed region)                                                            -> images/s, whole job
Weights and inputs are synthetic (random-init FLUX.1-dev architecture; no network / checkpoints in this environment).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "images/sec FLUX.1-dev 1024x1024 50-step bf16"
UNIT = "images/s"
HEIGHT = WIDTH = 1024
NUM_STEPS = 50
GUIDANCE = 3.5
L_TXT = 512
# algorithmic work: GEMM-only FLOPs incl. QK^T and PV
D = 3072


def dit_step_flops(l_img, l_txt):
    L = l_img + l_txt
    return 57 * (24 * L * D * D + 4 * L * L * D)


F_VAE_1024 = 1.0472e13


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return dict(tflops_sustained=j["bf16_tflops_sustained"], tflops_burst=j["bf16_tflops"], hbm=j["hbm_gbs"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(tflops_sustained=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# ----------------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            try:
                power.append(float(f[3]))
            except ValueError:
                pass
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # "under load": drop the idle tail by taking the median of the upper half
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        power.sort()
        # the loop is power-capped (sw_power_cap): joules per image = power x seconds per image is what to optimise
        pw = power[len(power) // 2:] if power else []
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w": (pw[len(pw) // 2] if pw else None)}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port (CPU restatement of the reference) timed on host cores
# ----------------------------------------------------------------------------------------------------------------------
def workload_config(height, width, num_steps, batch, world, quant, double_layers=19, single_layers=38) -> dict:
    """The `config` object of the JSON line - shared by both arms so that the driver compares like with like."""
    l_img = ((height + 15) // 16) * ((width + 15) // 16)
    return {"workload": f"FLUX.1-dev {height}x{width} {num_steps}-step bf16 batch={batch} per GPU "
                        f"(1 step = 1 image = {num_steps} DiT steps + VAE decode)",
            "l_img": l_img, "l_txt": L_TXT, "guidance": GUIDANCE, "weights": "random-init " + (quant or "bf16"),
            "parallelism": f"dp{world} (prompt sharding, NCCL weight broadcast at load only)",
            "l2": "inputs+weights per step (24 GB) >> 126 MB L2; no explicit flush needed",
            "double_layers": double_layers, "single_layers": single_layers}


class CpuReference:
    """The reference's CPU path for this workload, restated (oracle port: torch CPU = oneDNN/AVX-512, f32 math on
    bf16-rounded tensors - the reference's own CPU backend cannot multiply bf16), on all host cores.
    One SAMPLE = one DoubleStreamBlock + one SingleStreamBlock at the full width of the workload (L = l_img + 512):
    2 of the 57 x num_steps blocks of an image, ~3 s on 16 cores.  images/s = sample's share of an image's FLOPs /
    sample time (every block of a kind costs the same; the VAE share is extrapolated by FLOPs)."""

    def __init__(self, height=HEIGHT, width=WIDTH, num_steps=NUM_STEPS):
        import torch
        from oracle import flux as OF
        from oracle import ops as O
        self.torch = torch
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        cfg = OF.FluxConfig(num_layers=1, num_single_layers=1, guidance_embeds=True)
        self.orc = OF.FluxOracle(cfg, OF.make_weights(cfg), O.REF)
        h2, w2 = (height + 15) // 16, (width + 15) // 16
        self.l_img, self.l_txt, self.num_steps = h2 * w2, L_TXT, num_steps
        g = torch.Generator().manual_seed(1234)
        self.img = O.rb(torch.randn(1, self.l_img, D, generator=g))
        self.txt = O.rb(torch.randn(1, self.l_txt, D, generator=g))
        self.vec = O.rb(torch.randn(1, D, generator=g))
        self.pe = OF.embed_nd(OF.make_ids(h2, w2, self.l_txt), O.REF)
        L = self.l_img + self.l_txt
        f_block = 24 * L * D * D + 4 * L * L * D
        f_image = num_steps * 57 * f_block + F_VAE_1024 * (self.l_img / 4096.0)
        self.fraction = 2 * f_block / f_image

    def sample(self) -> tuple[float, float, float]:
        """-> (seconds, double-block seconds, single-block seconds)"""
        t0 = time.perf_counter()
        i2, t2 = self.orc.double_block(0, self.img, self.txt, self.vec, self.pe)
        t1 = time.perf_counter()
        self.orc.single_block(0, self.torch.cat([t2, i2], 1), self.vec, self.pe)
        t2_ = time.perf_counter()
        return t2_ - t0, t1 - t0, t2_ - t1

    def describe(self, n, secs, d, s) -> dict:
        value = n * self.fraction / secs
        return dict(value=value, unit=UNIT, cores=self.cores, kind="port",
                    sample=(f"oracle port (torch CPU f32 on bf16-rounded tensors, {self.cores} threads): {n} sample(s) of 1 double "
                            f"block ({d:.2f}s) + 1 single block ({s:.2f}s) at L={self.l_img + self.l_txt} = "
                            f"{self.fraction:.3e} of an image's FLOPs each; images/s = share / time"),
                    seconds_per_image=1.0 / value)


def cpu_reference_sample(height=HEIGHT, width=WIDTH, num_steps=NUM_STEPS) -> dict:
    """cpu_baseline leg of our arm: one sample (~3 s) plus ~10 s of weight generation."""
    ref = CpuReference(height, width, num_steps)
    ref.sample()  # warm-up: oneDNN primitive creation
    secs, d, s = ref.sample()
    return ref.describe(1, secs, d, s)


def run_reference(args):
    """`--impl reference`: W untimed samples, then exactly K timed samples (one "step" = one bounded sample of the
    workload, see CpuReference); rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t0 = time.perf_counter()
    ref = CpuReference(args.height, args.width, args.num_steps)
    for _ in range(max(1, args.warmup)):
        ref.sample()
    secs = d = s = 0.0
    for _ in range(args.steps):
        a, b, c = ref.sample()
        secs, d, s = secs + a, d + b, s + c
    desc = ref.describe(args.steps, secs, d / args.steps, s / args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": desc["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 math on bf16-rounded tensors",
        "data": "synthetic",
        "config": dict(workload_config(args.height, args.width, args.num_steps, args.batch, args.gpus, args.quant),
                       reference_arm="CPU path restated (oracle port); the Rust reference cannot be built here (no "
                                     "cargo/rustc).  One step = one bounded sample, value = sample share of an image / time",
                       sample_fraction_of_image=ref.fraction),
        "cpu_baseline": {k: desc[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": desc["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "seconds_per_image_extrapolated": desc["seconds_per_image"],
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU fallback for the product path "
                         "(use --impl reference for the CPU oracle port)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from diffusion_rs_b200 import build, lib as L
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    from diffusion_rs_b200.pipeline import (DiffusionGenerationParams, ModelSource, Pipeline, PromptEmbeds,
                                            calculate_shift, latent_hw, make_ids, patchify, shift_seq_len)
    lib = L.load()
    if args.attn_variant is not None:  # A/B switch for kernel experiments (scripts/); the default build is variant 0
        L.check(lib.fluxb200_set_flag(b"attn_variant", args.attn_variant))
    quant = args.quant
    t_load = time.perf_counter()
    pipe = Pipeline.load(ModelSource.synthetic("black-forest-labs/FLUX.1-dev", quant=quant, num_layers=args.layers,
                                               num_single_layers=args.single_layers,
                                               text_encoders=args.text_encoders))
    torch.cuda.synchronize()
    t_load = time.perf_counter() - t_load
    params = DiffusionGenerationParams(height=args.height, width=args.width, num_steps=args.num_steps,
                                       guidance_scale=GUIDANCE)
    h, w = latent_hw(params.height, params.width)
    h2, w2 = h // 2, w // 2
    l_img, l_txt = h2 * w2, L_TXT
    B = args.batch

    # ---- synthetic prompt embeddings / noise: B images per rank (weak scaling); every rank builds the same global
    #      prompt list, Pipeline.forward shards it by rank ----
    all_prompts = [f"synthetic prompt {r}-{i}" for r in range(world) for i in range(B)]
    all_embeds = [pipe.synthetic_embeds(p) for p in all_prompts]
    all_noise = torch.randn(world * B, 16, h, w, generator=torch.Generator().manual_seed(1234)).to(torch.bfloat16)
    embeds = all_embeds[rank * B:(rank + 1) * B]
    noise = all_noise[rank * B:(rank + 1) * B]

    # device-resident inputs for `value`
    txt_d = torch.stack([e.txt for e in embeds]).cuda()
    vec_d = torch.stack([e.vec for e in embeds]).cuda()
    lat0 = patchify(noise).contiguous().cuda()
    img_ids1, txt_ids1 = make_ids(h2, w2, l_txt)
    img_ids = img_ids1[None].repeat(B, 1, 1).contiguous().cuda()
    txt_ids = txt_ids1[None].repeat(B, 1, 1).contiguous().cuda()
    sc = pipe.scheduler
    # the reference's call site passes the latent channel count (flux/mod.rs:279), mirrored by Pipeline.forward too
    mu = calculate_shift(shift_seq_len(all_noise.shape, pipe.shift_mode), sc.base_image_seq_len, sc.max_image_seq_len,
                         sc.base_shift, sc.max_shift)
    timesteps = sc.get_timesteps(params.num_steps, mu)
    out_u8 = torch.empty(B, 16 * h2, 16 * w2, 3, dtype=torch.uint8, device="cuda")

    def image_resident():
        img = lat0.clone()
        pipe.transformer.denoise(img, img_ids, txt_d, txt_ids, vec_d, GUIDANCE, timesteps)
        pipe.vae.decode_packed_u8(img, h2, w2, out=out_u8)

    def image_e2e():  # the public API call a user makes: host embeddings + host noise in, host u8 images out
        return pipe.forward([PromptEmbeds(e.txt, e.vec) for e in all_embeds], params, noise=all_noise)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            fn()
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            every = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(every, t)
            rank_ms[fn.__name__] = [round(x.item() / k, 2) for x in every]  # per-rank ms per step: the spread is hardware
            ms = max(x.item() for x in every)
        return ms

    rank_ms = {}

    # ---- warm-up ----
    for _ in range(args.warmup):
        image_resident()
    torch.cuda.synchronize()

    # ---- timed: value (device-resident), with per-kernel-class event timing + launch counting + clocks ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = lib.fluxb200_launch_count(-1)
    ms_total = timed(image_resident, args.steps)
    launches = lib.fluxb200_launch_count(-1) - n0
    # one more image with per-kernel-class CUDA events (two event records per launch) for the roofline numbers
    nk = lib.fluxb200_profile_kinds()
    kms, kfl, kby = (C.c_double * nk)(), (C.c_double * nk)(), (C.c_double * nk)()
    kct = (C.c_uint64 * nk)()
    kinds, ms_prof = {}, None
    if args.kernel_timing:
        lib.fluxb200_profile_enable(1)
        ms_prof = timed(image_resident, 1)
        lib.fluxb200_profile_enable(0)
        L.check(lib.fluxb200_profile_collect(kms, kfl, kby, kct))
        kinds = {lib.fluxb200_profile_kind_name(i).decode(): dict(ms=kms[i], flops=kfl[i], bytes=kby[i],
                                                                    launches=int(kct[i])) for i in range(nk)}

    # ---- timed: e2e through Pipeline.forward with host buffers ----
    image_e2e()  # warm the pinned staging buffers
    ms_e2e = timed(image_e2e, args.steps)
    # ---- optional: the text encoders in front of the hot path, full-size random-init T5-XXL +
    #      CLIP-L: their own time per prompt batch, and the end-to-end number from TOKEN IDS instead of embeddings ----
    text = None
    if args.text_encoders:
        from diffusion_rs_b200.pipeline import PromptTokens
        gt = torch.Generator().manual_seed(77)
        toks = [PromptTokens(torch.randint(1, 32000, (l_txt,), generator=gt), torch.randint(1, 49000, (77,), generator=gt))
                for _ in range(world * B)]

        def encoders_only():
            pipe.encode_prompts(toks[rank * B:(rank + 1) * B])

        def image_e2e_tokens():
            return pipe.forward(toks, params, noise=all_noise)

        encoders_only()
        ms_enc = timed(encoders_only, 5) / 5
        image_e2e_tokens()
        ms_tok = timed(image_e2e_tokens, args.steps)
        text = {"t5": "t5-v1_1-xxl encoder (24 layers, d_model 4096), L=%d" % l_txt, "clip": "CLIP-L text (12 layers), L=77",
                "encoders_ms_per_prompt_batch": ms_enc,
                "e2e_from_token_ids": {"value": args.steps * B * world / (ms_tok / 1e3), "unit": UNIT,
                                       "ms_per_step": ms_tok / args.steps}}
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    images = args.steps * B * world
    value = images / (ms_total / 1e3)
    e2e_value = images / (ms_e2e / 1e3)
    pk = peaks()
    f_img = params.num_steps * dit_step_flops(l_img, l_txt) * (args.layers or 19) / 19 + F_VAE_1024
    gemm = kinds.get("gemm_tcgen05", {})
    roofline = None
    if gemm.get("launches"):
        ach = gemm["flops"] / (gemm["ms"] / 1e3) / 1e12
        traffic = None
        for name in ("r2_gemm_traffic.json", "r1b_gemm_traffic.json"):  # ncu --set full capture of this launch shape
            tj = ROOT / "profiles" / name
            if tj.exists():
                traffic = json.loads(tj.read_text()).get("dram_bytes_per_launch")
                break
        roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": ach,
                    "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tflops_sustained"],
                    "traffic": traffic,
                    "traffic_source": ("ncu --set full capture of the 4608x21504x3072 launch committed under profiles/ "
                                       "(dram bytes read + written per launch), not measured in this run"),
                    "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long step)",
                    "launches": gemm["launches"], "avg_launch_ms": gemm["ms"] / gemm["launches"],
                    "share_of_step": gemm["ms"] / ms_prof}
    # the other kernel classes against their own roofline (attention: tensor; the rest: HBM, algorithmic bytes)
    rooflines = {}
    for kind, v in kinds.items():
        if not v.get("launches") or kind == "gemm_tcgen05":
            continue
        if v["flops"] > 0:
            a = v["flops"] / (v["ms"] / 1e3) / 1e12
            rooflines[kind] = {"bound": "tensor", "achieved": a, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                               "frac": a / pk["tflops_sustained"], "share_of_step": v["ms"] / ms_prof}
        elif v["bytes"] > 0:
            a = v["bytes"] / (v["ms"] / 1e3) / 1e9
            rooflines[kind] = {"bound": "hbm", "achieved": a, "peak": pk["hbm"], "unit": "GB/s", "frac": a / pk["hbm"],
                               "share_of_step": v["ms"] / ms_prof}
    cpu = (cpu_reference_sample(args.height, args.width, args.num_steps)
           if (world == 1 and not args.no_cpu_baseline) else None)
    h2d, d2h = pipe.io_bytes(B, params)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if quant is None else f"bf16 (weights {quant})", "data": "synthetic",
        "config": workload_config(params.height, params.width, params.num_steps, B, world, quant,
                                  pipe.transformer.cfg.num_layers, pipe.transformer.cfg.num_single_layers),
        "dit_ms_per_step": None, "images_per_s_per_gpu": value / world,
        "frac_of_dense_gemm_roofline": (value / world) * f_img / (pk["tflops_sustained"] * 1e12),
        "roofline": roofline, "rooflines_other": rooflines, "kernels": kinds, "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
                                                                 if cpu else None),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "load_s": t_load, "text_encoders": text,
    }
    line["dit_ms_per_step"] = (ms_total / args.steps) / params.num_steps  # upper bound: includes the VAE share
    line["profiled_image_ms"] = ms_prof
    if rank_ms:  # value = max over ranks (slowest GPU of the box); the per-rank figures show how far the others are ahead
        line["ms_per_step_by_rank"] = rank_ms
    if clocks and clocks.get("power_w"):  # the loop is power-capped: energy per image is what kernel variants trade
        line["joules_per_image_per_gpu"] = clocks["power_w"] * (ms_total / args.steps / 1e3) / B
    graph_used, graph_note = pipe.transformer.denoise_info()
    line["step_graph"] = {"used": graph_used, "note": graph_note}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--quant", default=None, choices=[None, "nf4", "q4k"])
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--num-steps", type=int, default=NUM_STEPS, help="denoising steps per image")
    ap.add_argument("--batch", type=int, default=1, help="images per GPU per step")
    ap.add_argument("--layers", type=int, default=None, help="debug: reduced number of double blocks")
    ap.add_argument("--single-layers", type=int, default=None, help="debug: reduced number of single blocks")
    ap.add_argument("--no-kernel-timing", dest="kernel_timing", action="store_false")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--text-encoders", action="store_true",
                    help="also load full-size random-init T5-XXL + CLIP-L and report their time and the e2e number from token ids")
    ap.add_argument("--attn-variant", type=int, default=None, help="debug: attention kernel build (see attention.cu)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
