#!/usr/bin/env python
"""bench.py -- guided images/s on BASELINE.json config 2 (SD-1.4-shaped random-init UNet, fp16, 64x64 latent, 50 DDIM
steps, CFG 7.5, guidance active on the early steps), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [...]                       # the reference's algorithm on the host cores (oracle port)

A "step" is one guided image (50 denoising steps of the per-step guidance path + the UNet passes that drive it).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

META_PROMPT = 'a [robot:.6,.3,.4,.55] and a [blue vase:.2,.3,.4,.55]'
# "guidance active on the early denoising steps": the preset at reference utils/shared_state.py:20
HYPER = {"strict": False, "inside_loss_scale": .2, "outside_loss_scale": .2, "shrink_factor": .15,
         "thresholds": {0: .4, 2: .8, 4: .9, 8: .9}, "use_optimizer": False, "recurse_until": 14, "recurse_steps": 3}
BASE_SEED = 28
EMBED_SEED = 1234


def peaks():
    """(HBM GB/s, dense bf16 TFLOP/s, provenance): the measured numbers of this pool's B200s when the driver left them."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json; burst figures: kernel timed alone)"
    return 6650.0, 1650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ clock sampling
class ClockSampler:
    """SM clock + throttle reasons sampled every 200 ms while the timed regions run (pynvml, else nvidia-smi)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None

    def _loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self._stop.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, n in names.items():
                    if mask & bit:
                        self.reasons.add(n)
                self._stop.wait(0.2)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def start(self):
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ workload set-up
def setup_config(n_steps):
    from guided_attention_b200 import run as R, shared_state as S
    from guided_attention_b200.config import RunConfig
    import tempfile
    cfg = RunConfig(meta_prompt=META_PROMPT, output_path=tempfile.mkdtemp(prefix="ga_bench_"), half_precision=True,
                    n_inference_steps=n_steps)
    S.config = cfg
    S.curHyperParams = dict(HYPER)
    return cfg


def analytic_schedule(thresholds, n_steps, recurse_steps, recurse_until, max_refine=10):
    """UNet-pass counts of one image when no threshold is ever met (the case for a random-init UNet; the CUDA arm
    reports its measured counts next to this).  reference pipeline_guided_attention.py:925-1053, :501-581."""
    c = {"grad_fwd": 0, "bwd": 0, "cfg_fwd": 0, "loss_eval": 0}
    for i in range(n_steps):
        reps = recurse_steps if (i in thresholds and i <= recurse_until) else 1
        for _ in range(reps):
            c["grad_fwd"] += 1; c["loss_eval"] += 1
            if i in thresholds:
                c["grad_fwd"] += max_refine + 1; c["loss_eval"] += max_refine + 1; c["bwd"] += max_refine + 1
            c["cfg_fwd"] += 1
    return c


# ------------------------------------------------------------------------------------------------- reference arm
def cpu_component_times(unet_kind, threads):
    """One bounded sample of the reference's algorithm on the host cores (fp32, oracle port): a grad-enabled text-cond
    UNet forward with explicit-softmax attention hooks, one loss evaluation, one backward to the latents, one CFG
    forward (B=2)."""
    from oracle import oracle as O
    from tests.gpu_harness import setup_prompt, oracle_tokens, oracle_hyper
    from guided_attention_b200.substrate import DDIMScheduler, UNetConfig, build_unet
    from guided_attention_b200.run import synthetic_prompt_embeds
    torch.set_num_threads(threads)
    cfg = setup_prompt(META_PROMPT, HYPER)
    ucfg = UNetConfig.sd14() if unet_kind == "sd14" else UNetConfig.tiny()
    unet = build_unet(ucfg, seed=0)
    embeds = synthetic_prompt_embeds(cfg.prompt, ucfg.cross_attention_dim, seed=EMBED_SEED)
    lat = torch.randn(1, 4, 64, 64, generator=torch.Generator("cpu").manual_seed(BASE_SEED))
    pipe = O.OraclePipeline(unet, DDIMScheduler(), oracle_tokens(cfg), oracle_hyper(cfg))
    t = {}
    with torch.enable_grad():
        x = lat.clone().requires_grad_(True)
        t0 = time.perf_counter()
        unet(x, 981, encoder_hidden_states=embeds[1:2])
        t["grad_fwd"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        r = pipe._loss(attention_res=16, smooth_attentions=True, sigma=0.5, kernel_size=3, last_idx=-1)
        t["loss_eval"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        torch.autograd.grad(r.loss, x)
        t["bwd"] = time.perf_counter() - t0
    with torch.no_grad():
        t0 = time.perf_counter()
        unet(torch.cat([lat] * 2), 981, encoder_hidden_states=embeds)
        t["cfg_fwd"] = time.perf_counter() - t0
    return t, float(r.loss)


def cpu_images_per_s(times, schedule):
    return 1.0 / sum(times[k] * schedule[k] for k in schedule)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sched = analytic_schedule(HYPER["thresholds"], args.denoise_steps, HYPER["recurse_steps"], HYPER["recurse_until"])
    vals = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        times, _ = cpu_component_times(args.unet, threads)
        if i >= args.warmup:
            vals.append((cpu_images_per_s(times, sched), time.perf_counter() - t0, times))
    v = float(np.mean([x[0] for x in vals]))
    sample = ("per step: 1 grad-enabled B=1 UNet forward with explicit-softmax hooks + 1 loss evaluation + 1 backward to "
              "the latents + 1 CFG forward (B=2), fp32 on the host cores; extrapolated to one image with the "
              f"analytic UNet-pass schedule {sched} (no threshold met)")
    line = {"impl": "reference", "metric": "guided_images_per_s", "value": v, "unit": "img/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, "f32"),
            "cpu_baseline": {"value": v, "unit": "img/s", "cores": threads, "kind": "port", "sample": sample,
                             "component_s": {k: float(np.mean([x[2][k] for x in vals])) for k in vals[0][2]}},
            "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_config(args, dtype):
    return {"workload": "BASELINE config 2: SD-1.4-shaped UNet (random init, seed 0), 64x64 latent (512^2 image), "
                        f"meta_prompt '{META_PROMPT}', seeds {BASE_SEED}+, {args.denoise_steps} DDIM steps, CFG 7.5 "
                        f"(batch 2), {dtype}, thresholds {HYPER['thresholds']} (guidance on early steps), "
                        "recurse_steps 3, attention_res 16" + ("" if args.unet == "sd14" else " [TINY UNET: NOT THE BASELINE CONFIG]"),
            "unet": args.unet, "denoise_steps": args.denoise_steps, "images_per_step_per_gpu": 1,
            "cuda_graphs": not getattr(args, "no_graphs", False),
            "l2_policy": "inputs larger than L2: every UNet pass streams 1.7 GB of fp16 weights (L2 is 126 MB)",
            "parallelism": f"seed-sharded, {args.gpus} process(es), no hot-path collective"}


# ------------------------------------------------------------------------------------------------------ CUDA arm
def run_ours(args):
    from guided_attention_b200 import build as B
    B.build()
    from guided_attention_b200 import ops, run as R, shared_state as S, sweep
    from guided_attention_b200.pipeline_guided_attention import GuidedAttention
    from guided_attention_b200.ptp_utils import AttentionStore, register_attention_control
    from guided_attention_b200.substrate import DDIMScheduler, UNetConfig, WhitespaceTokenizer, build_unet
    import torch.distributed as dist

    rank, world, local_rank = sweep.init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (CUDA arm) needs a GPU: the guidance path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    cfg = setup_config(args.denoise_steps)
    ucfg = UNetConfig.sd14() if args.unet == "sd14" else UNetConfig.tiny()
    unet = build_unet(ucfg, seed=0, dtype=torch.float16, device=dev)
    pipe = GuidedAttention(unet=unet, scheduler=DDIMScheduler(), tokenizer=WhitespaceTokenizer())
    cfg.stable = pipe
    pipe.use_cuda_graphs = not args.no_graphs
    R.register_custom_loss("toLeftOf", R.ToLeftOf())
    R.overrideConfig(cfg)
    R.parseMetaPrompt(cfg)
    store = AttentionStore()
    register_attention_control(pipe, store)

    embeds_host = R.synthetic_prompt_embeds(cfg.prompt, ucfg.cross_attention_dim, seed=EMBED_SEED).pin_memory()
    unet_calls = {"n": 0}
    orig_forward = unet.forward

    def counted(*a, **k):
        unet_calls["n"] += 1
        return orig_forward(*a, **k)
    unet.forward = counted

    def host_latents(seed):
        g = torch.Generator("cpu").manual_seed(seed)
        return torch.randn(1, 4, 64, 64, generator=g).pin_memory()

    def image(seed, embeds, latents):
        S.cur_seed = seed
        gen = torch.Generator("cpu").manual_seed(seed)
        out = pipe(prompt=cfg.prompt, attention_store=store, attention_res=16, guidance_scale=7.5, generator=gen,
                   latents=latents, prompt_embeds=embeds[1:2], negative_prompt_embeds=embeds[0:1],
                   num_inference_steps=args.denoise_steps, thresholds=cfg.thresholds, scale_factor=20,
                   scale_range=(1., .5), smooth_attentions=True, sigma=0.5, kernel_size=3, sd_2_1=False,
                   output_type="latent")
        return out.images

    def seed_of(phase, step):   # every rank and every timed image gets its own seed (weak scaling)
        return BASE_SEED + phase * 1000 + step * world + rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        res = [fn(i) for i in range(n)]
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms), res

    embeds_dev = embeds_host.to(dev, non_blocking=True)
    for w in range(args.warmup):
        image(seed_of(0, w), embeds_dev, host_latents(seed_of(0, w)).to(dev))

    sampler = ClockSampler(local_rank)
    sampler.start()
    # (1) device-resident inputs: `value`
    dev_lat = [host_latents(seed_of(1, i)).to(dev) for i in range(args.steps)]
    ops.reset_launch_counts()
    pipe._count_pass("cfg", 0)
    cfg_before = pipe.pass_counts["cfg"]
    ms_value, _ = timed(lambda i: image(seed_of(1, i), embeds_dev, dev_lat[i]), args.steps)
    launches = ops.total_launches()
    counts = dict(ops.launch_counts)
    cfg_passes = pipe.pass_counts["cfg"] - cfg_before
    unet_passes = counts.get("guidance_tail_fwd", 0) + cfg_passes
    # (2) end to end: host buffers in, host buffer out, inside the timed region; final NCCL gather of the latents
    host_lat = [host_latents(seed_of(1, i)) for i in range(args.steps)]
    out_host = torch.empty(args.steps, 4, 64, 64, dtype=torch.float16).pin_memory()

    def e2e_step(i):
        lat = image(seed_of(1, i), embeds_host.to(dev, non_blocking=True), host_lat[i].to(dev, non_blocking=True))
        out_host[i].copy_(lat[0], non_blocking=True)
        return lat

    def e2e_all(_):
        res = [e2e_step(i) for i in range(args.steps)]
        if world > 1:   # the only collective of the workload: gather the final latents (32 KB / image)
            sweep.gather_results(torch.cat(res), args.steps * world, rank, world)
        torch.cuda.synchronize()
        return res
    ms_e2e, e2e_res = timed(e2e_all, 1)
    clocks = sampler.stop()

    # (3) roofline: one more image with per-launch CUDA events on the launching stream
    # (CUDA events cannot be recorded inside a graph replay: this pass runs the eager loop, same kernels)
    pipe.use_cuda_graphs = False
    ops.profiler = ops.LaunchProfiler()
    image(seed_of(1, 0), embeds_dev, dev_lat[0])
    prof = ops.profiler.summary()
    ops.profiler = None
    pipe.use_cuda_graphs = not args.no_graphs
    roofline, kernel_table = roofline_from(prof)

    # (4) extension: S seeds per UNet pass (`generate_batch`), same per-seed semantics; reported next to the headline
    batched = None
    if args.seeds_per_batch > 1:
        Sb = args.seeds_per_batch
        pipe.use_cuda_graphs = not args.no_graphs
        emb16 = embeds_dev.to(torch.float16)

        def batch_step(i):
            seeds = [seed_of(2, i * Sb + j) for j in range(Sb)]
            return pipe.generate_batch(cfg.prompt, store, seeds, emb16[1:2], emb16[0:1], attention_res=16,
                                       num_inference_steps=args.denoise_steps, guidance_scale=7.5,
                                       thresholds=cfg.thresholds)
        batch_step(0)                                   # warm-up: captures the batch-S graphs
        ms_b, _ = timed(batch_step, 1)
        batched = {"seeds_per_batch": Sb, "value": Sb * world / (ms_b / 1e3), "unit": "img/s",
                   "ms_per_batch": ms_b, "note": "extension (SURVEY 8e): per-seed results equal the one-seed path "
                   "(tests/test_gpu_parity.py::test_seed_batching_equals_separate_calls)"}

    # (5) the HBM-bound kernels at GPU-filling batch sizes (rank 0; last, because the tail timer re-installs the prompt)
    saturated = None
    if rank == 0 and not args.no_saturated:
        keep = (S.config, S.curHyperParams)
        torch.cuda.empty_cache()
        saturated = saturated_rooflines(dev)
        S.config, S.curHyperParams = keep
        torch.cuda.empty_cache()

    n_img = args.steps * world
    value = n_img / (ms_value / 1e3)
    e2e = n_img / (ms_e2e / 1e3)
    line = {"metric": "guided_images_per_s", "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": workload_config(args, "fp16"),
            "e2e": {"value": e2e, "unit": "img/s",
                    "h2d_bytes_per_step": int(embeds_host.numel() * 4 + 4 * 64 * 64 * 4),
                    "d2h_bytes_per_step": int(4 * 64 * 64 * 2)},
            "gpu_launches": launches, "gpu_launches_by_kernel": counts, "unet_passes": unet_passes,
            "clocks": clocks, "roofline": roofline, "kernels": kernel_table, "saturated_kernels": saturated,
            "batched": batched}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        times, _ = cpu_component_times(args.unet, threads)
        per_image = {"grad_fwd": counts.get("guidance_tail_fwd", 0) / args.steps,
                     "loss_eval": counts.get("guidance_tail_fwd", 0) / args.steps,
                     "bwd": counts.get("guidance_tail_bwd", 0) / args.steps,
                     "cfg_fwd": cfg_passes / args.steps}
        line["cpu_baseline"] = {
            "value": cpu_images_per_s(times, per_image), "unit": "img/s", "cores": threads, "kind": "port",
            "sample": "1 grad-enabled B=1 UNet forward with explicit-softmax hooks + 1 loss evaluation + 1 backward to "
                      "the latents + 1 CFG forward (B=2), fp32 oracle port on the host cores, extrapolated with the "
                      f"UNet-pass counts this run measured per image: {per_image}",
            "component_s": times}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def saturated_rooflines(dev):
    """The HBM-bound guidance kernels timed live at batch sizes that fill the GPU (the pipeline's own launches move
    0.7-11 MB and are launch-latency bound by size): same timing method as the roofline leg (CUDA graph of back-to-back
    launches between two CUDA events on the launching stream, rotating buffer sets larger than L2)."""
    from guided_attention_b200 import microbench
    peak = peaks()[0]
    out = []
    try:
        for direction in ("fwd", "bwd"):
            m = microbench.time_cross_attn(128, 8, 4096, 77, 40, torch.float16, with_acc=False, direction=direction,
                                           device=str(dev))
            out.append({"kernel": m["kernel"], "shape": "B=128 H=8 N=4096 T=77 d=40 fp16", "us": m["us"],
                        "bytes": m["bytes"], "gbs": m["gbs"], "frac": m["gbs"] / peak})
            m = microbench.time_cross_attn(256, 8, 1024, 77, 80, torch.float16, with_acc=True, direction=direction,
                                           device=str(dev))
            out.append({"kernel": m["kernel"], "shape": "B=256 H=8 N=1024 T=77 d=80 fp16 (maps)", "us": m["us"],
                        "bytes": m["bytes"], "gbs": m["gbs"], "frac": m["gbs"] / peak})
            m = microbench.time_tail(16, 5, 2, n_samples=2048, direction=direction, device=str(dev))
            out.append({"kernel": m["kernel"], "shape": "res 16, 5 layers x 2 slices, 2048 samples", "us": m["us"],
                        "bytes": m["bytes"], "gbs": m["gbs"], "frac": m["gbs"] / peak})
    except Exception as e:   # never lose the headline line to an auxiliary measurement
        out.append({"error": f"{type(e).__name__}: {e}"})
    return out


def ncu_traffic(kernel_prefix):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel, from the committed `ncu --set full`
    summary (profiles/r01_ncu_final_summary.json, captured with tools/ncu_all.sh at the shape the bench reports)."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_final_summary.json")
    if not os.path.isfile(path):
        return None
    with open(path) as f:
        rows = json.load(f)
    tot = [r.get("dram_traffic_bytes") for r in rows if r["kernel"].split("::")[-1].replace("void ", "").startswith(kernel_prefix)
           and r["report"].endswith("_b1.ncu-rep")]
    tot = [t for t in tot if t is not None]
    return float(sum(tot)) if tot else None


def roofline_from(prof):
    """Dominant kernel = the (kernel, shape) class with the largest summed device time inside the profiled image.
    Its launch duration is then measured kernel-only: the same launch (same shape, dtype, variant choice) replayed
    back-to-back from a CUDA graph between two CUDA events on the launching stream, on rotating buffer sets larger than
    L2 (`microbench.time_cross_attn` / `time_self_attn`).  The in-pipeline event time is kept next to it: it includes
    the host gaps of the eager profiling pass and only serves to rank the kernels.
    Cross-attention / tail kernels are HBM-bound (algorithmic bytes); the self-attention kernels are tensor-pipe bound
    (algorithmic FLOPs = the 2 GEMMs of the forward / 5 of the backward, not the recomputation the kernels add)."""
    from guided_attention_b200 import microbench
    hbm_peak, tf_peak, how = peaks()
    dev = f"cuda:{torch.cuda.current_device()}"      # every rank measures on its own GPU
    table = []
    for (name, key), d in prof.items():
        avg_us = d["ms"] * 1e3 / d["launches"]
        row = {"kernel": name, "shape": list(map(str, key)), "launches": d["launches"],
               "pipeline_event_us": avg_us, "total_ms": d["ms"]}
        row["flops_issued_per_launch" if name.startswith("self_attn") else "bytes_per_launch"] = d["bytes_per_launch"]
        table.append(row)
    table.sort(key=lambda r: -r["total_ms"])
    if not table:
        return None, table
    dts = {"torch.float16": torch.float16, "torch.bfloat16": torch.bfloat16, "torch.float32": torch.float32}
    for row in table[:8]:
        if row["kernel"].startswith("cross_attn"):
            B, H, N, T, dd = (int(x) for x in row["shape"][:5])
            m = microbench.time_cross_attn(B, H, N, T, dd, dts[row["shape"][5]], with_acc=row["shape"][6] == "True",
                                           direction=row["kernel"].split("_")[-1], device=dev)
            row["kernel_us"], row["gbs"] = m["us"], m["gbs"]
        elif row["kernel"].startswith("self_attn"):
            B, H, N, dd = (int(x) for x in row["shape"][:4])
            m = microbench.time_self_attn(B, H, N, dd, dts[row["shape"][4]], direction=row["kernel"].split("_")[-1],
                                          device=dev)
            row["kernel_us"], row["tflops"], row["tflops_issued"] = m["us"], m["tflops_algorithmic"], m["tflops_issued"]
            row["algorithmic_flops_per_launch"] = m["tflops_algorithmic"] * m["us"] * 1e6
    top = next((r for r in table if "kernel_us" in r), table[0])
    if "tflops" in top:
        roof = {"bound": "tensor", "kernel": top["kernel"], "shape": top["shape"], "achieved": top["tflops"],
                "peak": tf_peak, "unit": "TFLOP/s", "frac": top["tflops"] / tf_peak,
                "traffic": ncu_traffic(top["kernel"]) if top["shape"][:4] == ["1", "8", "4096", "40"] else None,
                "peak_source": how, "avg_launch_us": top["kernel_us"],
                "algorithmic_flops_per_launch": top["algorithmic_flops_per_launch"],
                "issued_tflops": top["tflops_issued"],
                "note": "fused exact self-attention at head_dim 40: per exponential only 4*d = 160 FLOPs of GEMM work "
                        "exist, so the SFU (16 ex2/clk/SM) caps this kernel far below the dense-GEMM peak; the "
                        "cross-attention / tail kernels (HBM-bound) are listed in `kernels` and in profiles/"}
    else:
        roof = {"bound": "hbm", "kernel": top["kernel"], "shape": top["shape"], "achieved": top.get("gbs"),
                "peak": hbm_peak, "unit": "GB/s", "frac": (top["gbs"] / hbm_peak) if "gbs" in top else None,
                "traffic": None, "peak_source": how, "avg_launch_us": top.get("kernel_us"),
                "algorithmic_bytes_per_launch": top.get("bytes_per_launch"),
                "note": "one launch at the pipeline's batch (B=1 text-cond pass, B=2 CFG pass) moves 0.7-11 MB: launch-"
                        "latency bound by size; profiles/ holds the batch sweep where the same kernels run "
                        "bandwidth-bound"}
    return roof, table[:10]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--unet", default="sd14", choices=["sd14", "tiny"], help="tiny is for plumbing checks only")
    ap.add_argument("--denoise-steps", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-saturated", action="store_true", help="skip the large-batch kernel rooflines")
    ap.add_argument("--seeds-per-batch", type=int, default=8,
                    help="also measure the seed-batched extension with this many seeds per UNet pass (0/1 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
