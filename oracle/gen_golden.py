"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/* by running the UNMODIFIED reference on CPU.

Run in the build container (where /root/reference is mounted):   python -m oracle.gen_golden
The reference ships no tests or golden vectors, so these fixtures -- outputs of the reference's own
code on seeded inputs -- are what pins the oracle (`oracle/oracle.py`) and, through it, the CUDA path.
Inputs are never stored: they are re-drawn from the recorded seeds (a checksum guards against RNG drift).
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.cases import (PARSE_CASES, MASK_CASES, LOSS_CASES, make_loss_inputs, PROCESSOR_CASE,  # noqa: E402
                          make_processor_inputs, E2E_CASE, make_e2e_inputs, PWW_CASE, pww_functional)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _tok():
    from guided_attention_b200.substrate import WhitespaceTokenizer
    return WhitespaceTokenizer()


def _setup(ref, meta_prompt, hyper=None, cfg_kw=None):
    tok = _tok()
    stable = types.SimpleNamespace(tokenizer=tok)
    cfg = ref_loader.make_config(ref, meta_prompt, tempfile.mkdtemp(prefix="ga_golden_"), stable=stable,
                                 **(cfg_kw or {}))
    hp = ref.state.get_hyperparam_states()[0]
    hp.update(hyper or {})
    ref.state.curHyperParams = hp
    ref.run.overrideConfig(cfg)
    ref.run.parseMetaPrompt(cfg)
    ref.state.cur_time_step_iter = 0
    ref.state.cur_seed = 0
    ref.state.sub_iteration = 0
    return cfg, tok


def _meta_to_json(ref, meta_info):
    out = []
    for sub, kind, payload in meta_info:
        if kind == ref.helpers.AnnotationType.BOX:
            payload = [payload.x, payload.y, payload.width, payload.height]
        elif kind == ref.helpers.AnnotationType.COOR:
            payload = list(payload)
        out.append([sub, kind.name, payload])
    return out


def gen_parse(ref):
    res = []
    for mp in PARSE_CASES:
        try:
            cfg, tok = _setup(ref, mp)
            td = {str(k): {"word": v["word"], "kind": v["loss_type"].name, "subprompt": v["subprompt"]}
                  for k, v in cfg.token_dict.items()}
            res.append({"meta_prompt": mp, "prompt": cfg.prompt, "meta_info": _meta_to_json(ref, cfg.meta_info),
                        "custom": {k: v[1] for k, v in cfg.custom_loss.items()}, "token_dict": td})
        except Exception as e:  # the error type is part of the behaviour
            res.append({"meta_prompt": mp, "error": type(e).__name__})
    return res


def gen_parse_fuzz(ref):
    """Reference behaviour on the seeded random meta-prompts of `cases.fuzz_parse_cases` -> reference_parse_fuzz.json."""
    from oracle.cases import fuzz_parse_cases
    res = []
    for mp in fuzz_parse_cases():
        try:
            cfg, tok = _setup(ref, mp)
            td = {str(k): {"word": v["word"], "kind": v["loss_type"].name, "subprompt": v["subprompt"]}
                  for k, v in cfg.token_dict.items()}
            res.append({"meta_prompt": mp, "prompt": cfg.prompt, "meta_info": _meta_to_json(ref, cfg.meta_info),
                        "custom": {k: v[1] for k, v in cfg.custom_loss.items()}, "token_dict": td})
        except Exception as e:
            res.append({"meta_prompt": mp, "error": type(e).__name__})
    return res


def gen_masks(ref):
    out = []
    for box, res, shrink in MASK_CASES:
        ref.state.curHyperParams = {"shrink_factor": shrink}
        r = ref.helpers.Rect(*box, 1).of_size(float(res) if res == 16 else res)
        m = [[1 if ref.helpers.inside_box(jj, ii, r) else 0 for jj in range(res)] for ii in range(res)]
        out.append({"box": list(box), "res": res, "shrink": shrink, "rect_at_res": [r.x, r.y, r.width, r.height],
                    "mask_rows": ["".join(map(str, row)) for row in m]})
    return out


def gen_gaussian(ref):
    out = []
    for k, s in ((3, 0.5), (3, 1.0), (3, 0.25)):
        w = ref.gaussian_smoothing.GaussianSmoothing(channels=1, kernel_size=k, sigma=s, dim=2).weight[0, 0]
        out.append({"kernel_size": k, "sigma": s, "weight": w.double().tolist()})
    return out


def gen_loss(ref, arrays):
    out = []
    for case in LOSS_CASES:
        cfg, tok = _setup(ref, case["meta_prompt"], case.get("hyper"), case.get("cfg"))
        Ps, checksum = make_loss_inputs(case)
        store = ref.ptp_utils.AttentionStore()
        store.num_att_layers = len(Ps)
        for P, place in zip(Ps, case["places"]):
            P.requires_grad_(True)
            store(P, True, place)

        class G(ref.pipeline.GuidedAttention):
            def save_viridis(self, *a, **k):
                pass
        g = G(unet=None, tokenizer=tok)
        g.prompt = cfg.prompt
        ld = g._aggregate_and_get_max_attention_per_token(store, 16, case.get("smooth", True), case.get("sigma", .5),
                                                          3, case.get("normalize_eot", False))
        loss, losses, unscaled = g._compute_loss(ld)
        f = lambda x: float(x) if not isinstance(x, (int, float)) else float(x)  # noqa: E731
        rec = {"name": case["name"], "input_checksum": checksum,
               "token_indices": list(cfg.token_dict.keys()),
               "max": [f(x) for x in ld["max_loss"]], "col": [f(x) for x in ld["col"]],
               "row": [f(x) for x in ld["row"]], "inside": [f(x) for x in ld["inside_loss"]],
               "outside": [f(x) for x in ld["outside_loss"]], "custom": f(ld["custom_loss"]),
               "losses": [[k, f(v)] for k, v in losses], "unscaled": [[k, f(v)] for k, v in unscaled],
               "total": f(loss)}
        thr = case.get("thresholds", {0: 1.0})
        rec["meets"] = {str(i): bool(g.meets_threshold(i, thr, unscaled)) for i in (0, -1, 7)}
        rec["thresholds"] = {str(k): v for k, v in thr.items()}
        if float(loss) != 0:
            grads = torch.autograd.grad(loss, Ps, retain_graph=False)
            same = all(torch.equal(grads[0], gr) for gr in grads[1:])
            rec["grad_same_across_layers"] = bool(same)
            rec["grad_absmean"] = float(grads[0].abs().mean())
            rec["grad_absmax"] = float(grads[0].abs().max())
            # d loss / d Abar = (layers * heads) * d loss / d P[layer][head]
            n_maps = sum(P.shape[0] for P in Ps)
            arrays[f"loss_{case['name']}_dAbar"] = (grads[0][0] * n_maps).reshape(16, 16, -1).numpy().astype(np.float32)
        out.append(rec)
    return out


def gen_processor(ref, arrays):
    from guided_attention_b200.substrate import CrossAttention
    case = PROCESSOR_CASE
    attn, attn_self, x, ctx = make_processor_inputs(case)
    ref.state.curHyperParams = ref.state.get_hyperparam_states()[0]
    ref.state.cur_time_step_iter = 0
    ref.state.config = types.SimpleNamespace(save_individual_CA_maps=False, token_dict={})
    store = ref.ptp_utils.AttentionStore()
    store.num_att_layers = 2
    # torch.cuda.empty_cache() at utils/ptp_utils.py:102 is a no-op without CUDA
    proc = ref.ptp_utils.AttendExciteCrossAttnProcessor(store, "down")
    y_cross = proc(attn, x, encoder_hidden_states=ctx)
    y_self = proc(attn_self, x)
    arrays["proc_cross_out"] = y_cross.detach().numpy()
    arrays["proc_cross_probs"] = store.attention_store["down_cross"][0].detach().numpy()
    arrays["proc_self_out"] = y_self.detach().numpy()
    return {"cross_out_sum": float(y_cross.sum()), "self_out_sum": float(y_self.sum()),
            "store_keys": {k: len(v) for k, v in store.attention_store.items()}, "cur_step": store.cur_step}


def gen_pww(ref, arrays):
    """The reference processor with paint-with-words active (utils/ptp_utils.py:113-138): output, stored
    probabilities and d L / d hidden_states for a seeded linear functional L."""
    from guided_attention_b200.substrate import DDIMScheduler
    case = PWW_CASE
    cfg, tok = _setup(ref, case["meta_prompt"], {"paint_with_words_stop": case["stop"],
                                                 "paint_with_words_weight": case["weight"]})
    sch = DDIMScheduler()
    sch.set_timesteps(case["steps"])
    ref.state.sigmas = (((1 - sch.alphas_cumprod) / sch.alphas_cumprod) ** 0.5).numpy()
    ref.state.timesteps = sch.timesteps
    ref.state.cur_time_step_iter = case["iter"]
    attn, _, x, ctx = make_processor_inputs(case)
    x.requires_grad_(True)
    store = ref.ptp_utils.AttentionStore()
    store.num_att_layers = 1
    proc = ref.ptp_utils.AttendExciteCrossAttnProcessor(store, "down")
    y = proc(attn, x, encoder_hidden_states=ctx)
    P = store.attention_store["down_cross"][0]
    L, _, _ = pww_functional(case, y, P)
    (gx,) = torch.autograd.grad(L, x)
    arrays["pww_out"] = y.detach().numpy()
    arrays["pww_probs"] = P.detach().numpy()
    arrays["pww_grad_x"] = gx.numpy()
    # the same call one iteration past `stop`: the bias must be off
    ref.state.cur_time_step_iter = case["stop"]
    store2 = ref.ptp_utils.AttentionStore()
    store2.num_att_layers = 1
    y_off = ref.ptp_utils.AttendExciteCrossAttnProcessor(store2, "down")(attn, x.detach(), encoder_hidden_states=ctx)
    arrays["pww_out_off"] = y_off.detach().numpy()
    ref.state.cur_time_step_iter = 0
    return {"sigma": float(ref.state.sigmas[int(sch.timesteps[case["iter"]])]), "functional": float(L),
            "token_indices": list(cfg.token_dict.keys())}


def gen_e2e(ref, arrays):
    """The reference's own `GuidedAttention.__call__` + patched UNet forward + refinement loop, on the tiny substrate
    UNet, CPU fp32."""
    case = E2E_CASE
    unet, embeds, latents0, gen = make_e2e_inputs(case)
    hyper = dict(case["hyper"])
    cfg, tok = _setup(ref, case["meta_prompt"], hyper)
    cfg.thresholds = hyper["thresholds"]
    pipe = ref_loader.make_pipeline(ref, unet, tok)

    class FixedTextEncoder:
        """The reference's `_encode_prompt` cannot take `prompt_embeds` (`text_inputs` is unbound on that branch,
        pipeline_guided_attention.py:105-199), so the synthetic embeddings enter through a stand-in text encoder:
        the empty (negative) prompt maps to embeds[0], anything else to embeds[1]."""
        dtype = torch.float32
        config = types.SimpleNamespace()

        def __call__(self, ids, attention_mask=None):
            empty = bool(ids[0, 1] == tok.eos_token_id)
            return (embeds[0:1] if empty else embeds[1:2],)
    pipe.text_encoder = FixedTextEncoder()
    cfg.stable = pipe
    saved = []
    pipe.save_image = lambda latent, tag: saved.append(tag)      # PNG side effects off
    pipe.save_viridis = lambda *a, **k: None
    ref.helpers.log_latent_stats = lambda *a, **k: None          # numpy.quantile diagnostics off
    ref.helpers.log_clear()
    store = ref.ptp_utils.AttentionStore()
    ref.ptp_utils.register_attention_control(pipe, store)
    n_fwd = [0]
    orig_forward = pipe.forward

    def counting_forward(*a, **k):
        n_fwd[0] += 1
        return orig_forward(*a, **k)
    pipe.forward = counting_forward
    out = pipe(prompt=cfg.prompt, attention_store=store, attention_res=16, guidance_scale=case["guidance_scale"],
               generator=gen, latents=latents0.clone(),
               num_inference_steps=case["steps"], max_iter_to_alter=25, run_standard_sd=False,
               thresholds=cfg.thresholds, scale_factor=20, scale_range=(1.0, 0.5), smooth_attentions=True, sigma=0.5,
               kernel_size=3, sd_2_1=False, output_type="latent_probe")
    lines = list(ref.helpers.lines)
    return {"num_att_layers": store.num_att_layers, "unet_forwards": n_fwd[0], "log": [l.strip() for l in lines]}


def main():
    ref = ref_loader.load()
    os.makedirs(GOLDEN, exist_ok=True)
    if "--skip-parse-fuzz" not in sys.argv:
        with open(os.path.join(GOLDEN, "reference_parse_fuzz.json"), "w") as f:
            json.dump({"generator": "oracle/gen_golden.py::gen_parse_fuzz", "cases": gen_parse_fuzz(ref)}, f, indent=0)
    if "--only-parse-fuzz" in sys.argv:
        print("parser fuzz fixture written to", GOLDEN)
        return
    arrays = {}
    doc = {"generator": "oracle/gen_golden.py", "torch": torch.__version__,
           "parse": gen_parse(ref), "masks": gen_masks(ref), "gaussian": gen_gaussian(ref),
           "loss": gen_loss(ref, arrays), "processor": gen_processor(ref, arrays), "pww": gen_pww(ref, arrays)}
    with open(os.path.join(GOLDEN, "reference_kat.json"), "w") as f:
        json.dump(doc, f, indent=1)
    np.savez_compressed(os.path.join(GOLDEN, "reference_kat.npz"), **arrays)
    if "--skip-e2e" not in sys.argv:
        e2e_arrays = {}
        # the reference pipeline returns images only; capture the final latents through decode_latents
        from guided_attention_b200.substrate import StableDiffusionPipelineBase
        captured = {}
        orig = StableDiffusionPipelineBase.decode_latents

        def capture(self, latents):
            captured["latents"] = latents.detach().clone()
            return orig(self, latents)
        StableDiffusionPipelineBase.decode_latents = capture
        try:
            e2e = gen_e2e(ref, e2e_arrays)
        finally:
            StableDiffusionPipelineBase.decode_latents = orig
        e2e_arrays["final_latents"] = captured["latents"].numpy().astype(np.float32)
        with open(os.path.join(GOLDEN, "reference_e2e.json"), "w") as f:
            json.dump(e2e, f, indent=1)
        np.savez_compressed(os.path.join(GOLDEN, "reference_e2e.npz"), **e2e_arrays)
    print("golden fixtures written to", GOLDEN)


if __name__ == "__main__":
    main()
