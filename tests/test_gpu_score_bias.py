"""GPU parity tests of the optional score bias of the cross-attention kernels (SURVEY 8 f4 + the a1 `attention_mask`
contract): paint-with-words against fixtures produced by the reference processor itself
(reference utils/ptp_utils.py:113-138; tests/golden/reference_kat.* `pww_*`, generator oracle/gen_golden.py::gen_pww),
the additive attention mask against the oracle, and the pipeline with paint-with-words on against the oracle pipeline.

Tolerances (north_star): fp32 <= 1e-3 relative, fp16 <= 2e-2."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle.cases import PWW_CASE, E2E_CASE, make_processor_inputs, make_e2e_inputs, pww_functional
from tests.gpu_harness import setup_prompt, oracle_tokens, oracle_hyper, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FP32_RTOL, FP16_RTOL = 1e-3, 2e-2


def _pww_setup(iteration):
    from guided_attention_b200 import shared_state as S
    from guided_attention_b200.substrate import DDIMScheduler
    case = PWW_CASE
    cfg = setup_prompt(case["meta_prompt"], {"paint_with_words_stop": case["stop"],
                                             "paint_with_words_weight": case["weight"]})
    sch = DDIMScheduler()
    sch.set_timesteps(case["steps"])
    S.sigmas = (((1 - sch.alphas_cumprod) / sch.alphas_cumprod) ** 0.5).numpy()
    S.timesteps = sch.timesteps
    S.cur_time_step_iter = iteration
    return cfg


@pytest.mark.parametrize("dtype,rtol", [(torch.float32, FP32_RTOL), (torch.float16, FP16_RTOL)])
def test_paint_with_words_against_reference_golden(kat, kat_arrays, dtype, rtol):
    """The product processor with `paint_with_words_stop` > cur_time_step_iter against the REFERENCE's processor:
    layer output, per-head probabilities, head-summed accumulator, and the gradient of a seeded linear functional
    w.r.t. the hidden states (it runs through the global max of the scores)."""
    from guided_attention_b200.ptp_utils import AttendExciteCrossAttnProcessor, AttentionStore
    case, rec = PWW_CASE, kat["pww"]
    cfg = _pww_setup(case["iter"])
    assert list(cfg.token_dict.keys()) == rec["token_indices"]
    attn, _, x, ctx = make_processor_inputs(case)
    attn = attn.to(DEV, dtype)
    store = AttentionStore()
    store.num_att_layers = 1
    proc = AttendExciteCrossAttnProcessor(store, "down")
    xd = x.to(DEV, dtype).requires_grad_(True)
    y = proc(attn, xd, encoder_hidden_states=ctx.to(DEV, dtype))
    maps = store.get_average_attention()["down_cross"][0]
    assert rel_err(y.detach().float().cpu().numpy(), kat_arrays["pww_out"]) < rtol
    gold_P = kat_arrays["pww_probs"]
    assert rel_err(maps.probs().float().cpu().numpy(), gold_P) < rtol
    B, H = case["batch"], case["heads"]
    assert rel_err(maps.acc.detach().cpu().numpy(), gold_P.reshape(B, H, *gold_P.shape[1:]).sum(1)) < rtol
    # the same functional the fixture differentiated: <y, R1> + <sum_h P, R2>
    _, r1, r2 = pww_functional(case, torch.from_numpy(kat_arrays["pww_out"]), torch.from_numpy(gold_P))
    L = (y.float() * r1.to(DEV)).sum() + (maps.acc * r2.to(DEV)).sum()
    assert float(L) == pytest.approx(rec["functional"], rel=rtol, abs=rtol * 50)
    (gx,) = torch.autograd.grad(L, xd)
    assert rel_err(gx.float().cpu().numpy(), kat_arrays["pww_grad_x"]) < (3 * rtol if dtype == torch.float16 else rtol)
    # the bias really is in play: the unbiased output is far outside the tolerance
    assert rel_err(kat_arrays["pww_out_off"], kat_arrays["pww_out"]) > 10 * rtol


def test_paint_with_words_is_off_from_stop_on(kat_arrays):
    """cur_time_step_iter == stop: the reference skips the bias (utils/ptp_utils.py:114); here the coefficient is 0."""
    from guided_attention_b200.ptp_utils import AttendExciteCrossAttnProcessor, AttentionStore
    case = PWW_CASE
    _pww_setup(case["stop"])
    attn, _, x, ctx = make_processor_inputs(case)
    store = AttentionStore()
    store.num_att_layers = 1
    y = AttendExciteCrossAttnProcessor(store, "down")(attn.to(DEV), x.to(DEV), encoder_hidden_states=ctx.to(DEV))
    assert rel_err(y.cpu().numpy(), kat_arrays["pww_out_off"]) < FP32_RTOL


def test_paint_with_words_gradient_through_the_max_vs_oracle():
    """ops level, SD-1.4 16x16 shape (8 heads x 160), fp32: forward, accumulator and dQ against oracle autograd, with
    the upstream gradient concentrated on the biased columns so the max term is a visible share of dQ."""
    from guided_attention_b200 import ops
    cfg = _pww_setup(0)
    H, d, N, T, B = 8, 160, 256, 77, 2
    g = torch.Generator("cpu").manual_seed(5)
    q = torch.randn(B, N, H * d, generator=g)
    k = torch.randn(B, T, H * d, generator=g)
    v = torch.randn(B, T, H * d, generator=g)
    scale = d ** -0.5
    tokens = oracle_tokens(cfg)
    from guided_attention_b200 import shared_state as S
    pww = O.PaintWithWords(tokens, float(S.get_sigma()), PWW_CASE["weight"], 0.15)
    qo = q.clone().requires_grad_(True)
    p, o = O.cross_attention(O.head_to_batch(qo, H), O.head_to_batch(k, H), O.head_to_batch(v, H), scale, pww=pww)
    acc_o = p.reshape(B, H, N, T).sum(1)
    w_acc = torch.zeros(B, N, T)
    for t in tokens:
        w_acc[:, :, t.index] = torch.randn(B, N, generator=g)
    w_out = torch.randn(B, N, H * d, generator=g) * 0.1
    Lo = (O.batch_to_head(o, H) * w_out).sum() + (acc_o * w_acc).sum()
    (go,) = torch.autograd.grad(Lo, qo)

    from guided_attention_b200.ptp_utils import PaintWithWords
    st = PaintWithWords()
    st.update(torch.device(DEV))
    masks, cols = st.masks_for(N, torch.device(DEV))
    assert cols == [t.index for t in tokens if t.kind == O.BOX]
    bias = ops.ScoreBias(pww_masks=masks, pww_columns=cols, pww_coef=st.coef)
    qd = q.to(DEV).requires_grad_(True)
    out, acc = ops.cross_attention(qd, k.to(DEV), v.to(DEV), H, scale, want_acc=True, bias=bias)
    assert rel_err(out.detach().cpu().numpy(), O.batch_to_head(o, H).detach().numpy()) < FP32_RTOL
    assert rel_err(acc.detach().cpu().numpy(), acc_o.detach().numpy()) < FP32_RTOL
    L = (out * w_out.to(DEV)).sum() + (acc * w_acc.to(DEV)).sum()
    (gd,) = torch.autograd.grad(L, qd)
    assert rel_err(gd.cpu().numpy(), go.numpy()) < FP32_RTOL
    # the row that owns the max got the rank-1 term: without it the error there is visible
    packed = int(bias.smax.item())
    flat = 0xFFFFFFFF - (packed & 0xFFFFFFFF)
    s_all = O.attention_scores(O.head_to_batch(q, H), O.head_to_batch(k, H), scale)
    assert flat == int(s_all.argmax())


@pytest.mark.parametrize("dtype,rtol", [(torch.float32, FP32_RTOL), (torch.float16, FP16_RTOL)])
def test_attention_mask_on_cross_layers_vs_oracle(dtype, rtol):
    """`attention_mask` (reference utils/ptp_utils.py:135-136) is added to the scores as given, broadcast like
    `scores + mask`: a (B*H, 1, T) padding-style mask and a full (B*H, N, T) one; forward, maps and dQ."""
    from guided_attention_b200 import ops
    H, d, N, T, B = 4, 32, 200, 77, 2
    g = torch.Generator("cpu").manual_seed(11)
    q = torch.randn(B, N, H * d, generator=g)
    k = torch.randn(B, T, H * d, generator=g)
    v = torch.randn(B, T, H * d, generator=g)
    scale = d ** -0.5
    pad = torch.zeros(B * H, 1, T)
    pad[:, :, 60:] = -10000.0
    full = torch.randn(B * H, N, T, generator=g)
    for mask in (pad, full):
        qo = q.to(dtype).float().requires_grad_(True)
        p, o = O.cross_attention(O.head_to_batch(qo, H), O.head_to_batch(k.to(dtype).float(), H),
                                 O.head_to_batch(v.to(dtype).float(), H), scale, attention_mask=mask)
        wo = torch.randn(B, N, H * d, generator=g)
        wa = torch.randn(B, N, T, generator=g)
        (go,) = torch.autograd.grad((O.batch_to_head(o, H) * wo).sum() + (p.reshape(B, H, N, T).sum(1) * wa).sum(), qo)
        qd = q.to(DEV, dtype).requires_grad_(True)
        out, acc = ops.cross_attention(qd, k.to(DEV, dtype), v.to(DEV, dtype), H, scale, want_acc=True,
                                       bias=ops.ScoreBias(mask=mask.to(DEV)))
        assert rel_err(out.detach().float().cpu().numpy(), O.batch_to_head(o, H).detach().numpy()) < rtol
        assert rel_err(acc.detach().cpu().numpy(), p.reshape(B, H, N, T).sum(1).detach().numpy()) < rtol
        (gd,) = torch.autograd.grad((out.float() * wo.to(DEV)).sum() + (acc * wa.to(DEV)).sum(), qd)
        assert rel_err(gd.float().cpu().numpy(), go.numpy()) < (5 * rtol if dtype == torch.float16 else rtol)
    # a mask that does not broadcast fails like the reference's `attention_scores + attention_mask`
    with pytest.raises(RuntimeError):
        ops.cross_attention(q.to(DEV), k.to(DEV), v.to(DEV), H, scale, bias=ops.ScoreBias(mask=torch.zeros(3, 5, device=DEV)))


def test_stored_maps_are_tensor_like_on_demand():
    """`AttentionStore.get_average_attention()` keeps the reference's six keys and its items can be consumed by code
    written against the reference's probability tensors (utils/ptp_utils.py:282-288): reshape / cat / indexing
    materialise the per-head maps lazily."""
    from guided_attention_b200.ptp_utils import AttendExciteCrossAttnProcessor, AttentionStore
    from oracle.cases import PROCESSOR_CASE
    setup_prompt()
    attn, _, x, ctx = make_processor_inputs(PROCESSOR_CASE)
    store = AttentionStore()
    store.num_att_layers = 1
    AttendExciteCrossAttnProcessor(store, "down")(attn.to(DEV), x.to(DEV), encoder_hidden_states=ctx.to(DEV))
    kept = store.get_average_attention()
    assert sorted(kept) == sorted(AttentionStore.get_empty_store())
    item = kept["down_cross"][0]
    # the reference's own aggregate_attention body, verbatim semantics
    out = []
    for it in kept["down_cross"]:
        if it.shape[1] == 8 ** 2:
            out.append(it.reshape(1, -1, 8, 8, it.shape[-1])[0])
    out = torch.cat(out, dim=0)
    agg = out.sum(0) / out.shape[0]
    want = item.acc.sum(0) / item.shape[0]
    assert rel_err(agg.reshape(64, -1).cpu().numpy(), want.cpu().numpy()) < FP32_RTOL
    assert torch.cat([item, item]).shape[0] == 2 * item.shape[0] and item[0].shape == (64, 77)
    assert float(item.sum()) == pytest.approx(item.shape[0] * 64, rel=1e-3)     # rows of P sum to one


def _psnr(a, b):
    mse = float(((a - b) ** 2).mean())
    peak = float(np.abs(b).max())
    return 10 * np.log10(peak * peak / max(mse, 1e-30))


@pytest.mark.parametrize("graphs", [False, True])
def test_pipeline_with_paint_with_words_matches_oracle_pipeline(graphs):
    """Whole guided images on the tiny UNet, fp32, `paint_with_words_stop` = 2 (bias on for iterations 0 and 1, off
    after), eager and CUDA-graph execution, against the oracle pipeline (CPU) with the same setting: same number of
    UNet passes, PSNR > 55 dB, cosine > 0.9999."""
    from guided_attention_b200.pipeline_guided_attention import GuidedAttention
    from guided_attention_b200.ptp_utils import AttentionStore, register_attention_control
    from guided_attention_b200.substrate import DDIMScheduler
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    case = E2E_CASE
    hyper = dict(case["hyper"], paint_with_words_stop=2, paint_with_words_weight=1.0)
    unet, embeds, latents, gen = make_e2e_inputs(case)
    cfg = setup_prompt(case["meta_prompt"], hyper)
    cfg.thresholds = hyper["thresholds"]
    opipe = O.OraclePipeline(unet, DDIMScheduler(), oracle_tokens(cfg), oracle_hyper(cfg),
                             recurse_steps=hyper["recurse_steps"], recurse_until=hyper["recurse_until"], pww_stop=2,
                             pww_weight=1.0)
    trace = opipe(embeds, latents.clone(), case["latent_seed"], num_inference_steps=case["steps"],
                  guidance_scale=case["guidance_scale"], thresholds=hyper["thresholds"])
    gold = trace.latents.numpy()
    plain = O.OraclePipeline(make_e2e_inputs(case)[0], DDIMScheduler(), oracle_tokens(cfg), oracle_hyper(cfg),
                             recurse_steps=hyper["recurse_steps"], recurse_until=hyper["recurse_until"])
    no_pww = plain(embeds, latents.clone(), case["latent_seed"], num_inference_steps=case["steps"],
                   guidance_scale=case["guidance_scale"], thresholds=hyper["thresholds"]).latents.numpy()
    assert _psnr(no_pww, gold) < 45          # the bias changes the image: the comparison below is not vacuous

    unet_d = make_e2e_inputs(case)[0].to(DEV)
    pipe = GuidedAttention(unet=unet_d, scheduler=DDIMScheduler(), tokenizer=cfg.stable.tokenizer)
    cfg.stable = pipe
    pipe.use_cuda_graphs = graphs
    store = AttentionStore()
    register_attention_control(pipe, store)
    out = pipe(prompt=cfg.prompt, attention_store=store, attention_res=16, guidance_scale=case["guidance_scale"],
               generator=gen, latents=latents.clone(), prompt_embeds=embeds[1:2], negative_prompt_embeds=embeds[0:1],
               num_inference_steps=case["steps"], thresholds=cfg.thresholds, output_type="latent")
    got = out.images.float().cpu().numpy()
    assert pipe.pass_counts["eval"] + pipe.pass_counts["cfg"] + (pipe.pass_counts["update"] if graphs else 0) \
        == trace.unet_forwards
    cos = float((got * gold).sum() / (np.linalg.norm(got) * np.linalg.norm(gold)))
    assert cos > 0.9999 and _psnr(got, gold) > 55, (cos, _psnr(got, gold))
