"""GPU tests of the device-side step driver (csrc/step_driver.cu): the refinement loop, threshold tests,
recursion rounds and re-noising of reference pipeline_guided_attention.py:475-581, :925-1053, :1074-1088 running inside
one CUDA graph with conditional nodes.  Bar: the SAME UNet passes and bit-identical latents as the host-driven loop
replaying the same captured programs, and the reference's own `__call__` latents within the existing e2e bound."""
import numpy as np
import pytest
import torch

from oracle.cases import E2E_CASE, make_e2e_inputs
from tests.gpu_harness import setup_prompt

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pipe(dtype, meta_prompt=None, hyper=None, cfg_kw=None):
    from guided_attention_b200.pipeline_guided_attention import GuidedAttention
    from guided_attention_b200.ptp_utils import AttentionStore, register_attention_control
    from guided_attention_b200.substrate import DDIMScheduler
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    case = E2E_CASE
    unet, embeds, _, _ = make_e2e_inputs(case)
    hyper = dict(case["hyper"], **(hyper or {}))
    cfg = setup_prompt(meta_prompt or case["meta_prompt"], hyper, cfg_kw)
    cfg.thresholds = hyper["thresholds"]
    pipe = GuidedAttention(unet=unet.to(DEV, dtype), scheduler=DDIMScheduler(), tokenizer=cfg.stable.tokenizer)
    cfg.stable = pipe
    pipe.use_cuda_graphs = True
    store = AttentionStore()
    register_attention_control(pipe, store)
    return pipe, store, cfg, embeds, case


def _run(pipe, store, cfg, embeds, case, seed, device_side, steps=None, **kw):
    from guided_attention_b200 import ops
    pipe.device_side_control = device_side
    pipe.pass_counts = {"eval": 0, "update": 0, "cfg": 0}
    ops.reset_launch_counts()
    gen = torch.Generator("cpu").manual_seed(seed)
    lat = torch.randn(1, 4, 64, 64, generator=torch.Generator("cpu").manual_seed(seed))
    out = pipe(prompt=cfg.prompt, attention_store=store, attention_res=16, guidance_scale=7.5, generator=gen,
               latents=lat, prompt_embeds=embeds[1:2], negative_prompt_embeds=embeds[0:1],
               num_inference_steps=steps or case["steps"], thresholds=cfg.thresholds, output_type="latent", **kw)
    return out.images.clone(), dict(pipe.pass_counts), dict(ops.launch_counts), pipe.control_mode


def _psnr(a, b):
    mse = float(((a - b) ** 2).mean())
    peak = float(np.abs(b).max())
    return 10 * np.log10(peak * peak / max(mse, 1e-30))


def test_device_driver_equals_host_loop_and_reference_golden(e2e_golden):
    """Tiny UNet, the e2e golden case (refinement that ends early on some rounds, recursion, re-noising).
    fp16 (every kernel on the path is run-to-run deterministic): device driver == host loop BIT FOR BIT, same pass
    counts, same launches of this library's kernels.  fp32 (cuDNN's fp32 convolution backward is not run-to-run
    bit-stable on this stack, host loop vs host loop differs by ~4e-4 too): both within the e2e bound of the
    REFERENCE's own `__call__` latents."""
    doc, arrays = e2e_golden
    pipe, store, cfg, embeds, case = _pipe(torch.float16)
    host, n_host, _, mode_h = _run(pipe, store, cfg, embeds, case, case["latent_seed"], False)
    dev, n_dev, _, mode_d = _run(pipe, store, cfg, embeds, case, case["latent_seed"], True)
    assert (mode_h, mode_d) == ("host-graphs", "device")
    assert n_dev == n_host, (n_dev, n_host)
    assert sum(n_dev.values()) == doc["unet_forwards"]
    assert torch.equal(dev, host)
    # a second image reuses the driver; the counters it reports are per image
    dev2, n_dev2, l_dev2, _ = _run(pipe, store, cfg, embeds, case, 29, True)
    host2, n_host2, l_host2, _ = _run(pipe, store, cfg, embeds, case, 29, False)
    assert torch.equal(dev2, host2) and n_dev2 == n_host2
    assert l_dev2 == l_host2            # the same kernels of this library ran, program by program
    assert not torch.equal(dev2, dev)
    assert pipe.last_step_counters["rounds"] >= case["steps"]

    pipe, store, cfg, embeds, case = _pipe(torch.float32)
    dev32, n32, _, mode = _run(pipe, store, cfg, embeds, case, case["latent_seed"], True)
    assert mode == "device" and sum(n32.values()) == doc["unet_forwards"]
    got, gold = dev32.float().cpu().numpy(), arrays["final_latents"]
    cos = float((got * gold).sum() / (np.linalg.norm(got) * np.linalg.norm(gold)))
    assert cos > 0.99999 and _psnr(got, gold) > 55, (cos, _psnr(got, gold))


@pytest.mark.parametrize("variant", ["early_exit", "never_met", "always_met", "keyword", "avg_within", "no_recursion",
                                      "update_every_step"])
def test_device_driver_control_flow_variants(variant):
    """Thresholds that are met at once / never / after a few refinement iterations, a keyword (custom-loss) group in
    the threshold test, sub-prompt averaging, no recursion, and `only_update_on_threshold_steps = False` (an update on
    every early step): pass counts and latents equal the host loop's."""
    meta, hyper, cfg_kw = None, {}, None
    if variant == "never_met":
        hyper = {"thresholds": {0: 0.01, 1: 0.01, 3: 0.01}, "recurse_steps": 3, "recurse_until": 2}
    elif variant == "always_met":
        hyper = {"thresholds": {0: 1e9, 2: 1e9}}
    elif variant == "early_exit":
        hyper = {"thresholds": {0: 3.4605, 1: 3.459, 2: 3.455}, "recurse_steps": 3, "recurse_until": 3}
    elif variant == "keyword":
        meta = 'a [robot:.55,.3,.4,.55] and a vase with a lamp [CustomLoss:toLeftOf (vase,lamp)]'
        hyper = {"thresholds": {0: 1.6, 2: 1.5}}
    elif variant == "avg_within":
        cfg_kw = {"sub_prompt_avg_within": True}
        hyper = {"thresholds": {0: 1.7, 2: 1.6}}
    elif variant == "no_recursion":
        hyper = {"recurse_steps": 1}
    pipe, store, cfg, embeds, case = _pipe(torch.float16, meta, hyper, cfg_kw)
    if meta is not None:
        from guided_attention_b200.run import synthetic_prompt_embeds
        embeds = synthetic_prompt_embeds(cfg.prompt, embeds.shape[-1])
    kw = {}
    if variant == "update_every_step":
        cfg.only_update_on_threshold_steps = False
        kw = {"max_iter_to_alter": 3}
    host, n_host, _, _ = _run(pipe, store, cfg, embeds, case, 28, False, **kw)
    dev, n_dev, _, mode = _run(pipe, store, cfg, embeds, case, 28, True, **kw)
    assert mode == "device"
    assert n_dev == n_host, (variant, n_dev, n_host, pipe.last_step_counters)
    assert torch.equal(dev, host), variant
    if variant == "never_met":
        assert pipe.last_step_counters["refine_iterations"] == 10 * 3 * 2 + 10     # steps 0,1 x 3 rounds, step 3 x 1
    if variant == "always_met":
        assert pipe.last_step_counters["refine_iterations"] == 0


def test_device_driver_full_size_config2_fp16():
    """BASELINE config 2 (SD-1.4 shape, fp16, the bench's thresholds, 50 steps): one image through the device driver
    and one through the host loop -- identical latents, 190 / 132 / 58 passes; no host read of a loss in between."""
    from guided_attention_b200.pipeline_guided_attention import GuidedAttention
    from guided_attention_b200.ptp_utils import AttentionStore, register_attention_control
    from guided_attention_b200.substrate import DDIMScheduler, UNetConfig, build_unet
    from guided_attention_b200.run import synthetic_prompt_embeds
    hyper = {"strict": False, "inside_loss_scale": .2, "outside_loss_scale": .2, "shrink_factor": .15,
             "thresholds": {0: .4, 2: .8, 4: .9, 8: .9}, "use_optimizer": False, "recurse_until": 14,
             "recurse_steps": 3}
    cfg = setup_prompt(hyper=hyper)
    cfg.thresholds = hyper["thresholds"]
    unet = build_unet(UNetConfig.sd14(), seed=0, dtype=torch.float16, device=DEV)
    pipe = GuidedAttention(unet=unet, scheduler=DDIMScheduler(), tokenizer=cfg.stable.tokenizer)
    cfg.stable = pipe
    pipe.use_cuda_graphs = True
    store = AttentionStore()
    register_attention_control(pipe, store)
    embeds = synthetic_prompt_embeds(cfg.prompt, 768)
    case = dict(E2E_CASE, steps=50)
    dev, n_dev, _, mode = _run(pipe, store, cfg, embeds, case, 28, True)
    host, n_host, _, _ = _run(pipe, store, cfg, embeds, case, 28, False)
    assert mode == "device"
    assert n_dev == n_host == {"eval": 58, "update": 132, "cfg": 58}, (n_dev, n_host)
    assert torch.isfinite(dev.float()).all()
    assert torch.equal(dev, host)


def test_use_optimizer_graphed_paths_match_the_eager_optimizer_loop():
    """`use_optimizer` (reference :495-497, :545-547: the refinement loop steps the latents with
    torch.optim.SGD(lr = step_size / 2.5, momentum = 0.8), a fresh optimizer per refinement): the eager loop runs the
    real torch optimizer; the graphed host loop and the device driver run the captured momentum program.  fp16:
    device == host-graphs bit for bit; both against eager within the graph-vs-eager bound."""
    hyper = {"use_optimizer": True, "thresholds": {0: 0.01, 2: 0.01}, "recurse_steps": 2, "recurse_until": 1}
    pipe, store, cfg, embeds, case = _pipe(torch.float16, None, hyper)
    pipe.use_cuda_graphs = False
    eager, n_eager, _, mode_e = _run(pipe, store, cfg, embeds, case, 28, False)
    pipe.use_cuda_graphs = True
    host, n_host, _, mode_h = _run(pipe, store, cfg, embeds, case, 28, False)
    dev, n_dev, _, mode_d = _run(pipe, store, cfg, embeds, case, 28, True)
    assert (mode_e, mode_h, mode_d) == ("eager", "host-graphs", "device")
    assert n_dev == n_host
    assert n_eager["eval"] + n_eager["cfg"] == sum(n_host.values())      # eager counts every UNet forward as eval / cfg
    assert torch.equal(dev, host)
    a, b = host.float().cpu().numpy(), eager.float().cpu().numpy()
    cos = float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b)))
    assert cos > 0.9999 and _psnr(a, b) > 45, (cos, _psnr(a, b))
    # momentum matters: the plain-SGD image of the same seed is a different image
    pipe2, store2, cfg2, embeds2, case2 = _pipe(torch.float16, None, dict(hyper, use_optimizer=False))
    plain, _, _, _ = _run(pipe2, store2, cfg2, embeds2, case2, 28, True)
    assert _psnr(plain.float().cpu().numpy(), a) < 40
