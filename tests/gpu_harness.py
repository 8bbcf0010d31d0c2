"""Shared harness for the GPU parity tests, `smoke()` and bench.py's self-check: builds a seeded synthetic stack of
cross-attention layers, runs the fused CUDA path through the C ABI and the CPU oracle on the same inputs."""
from __future__ import annotations

import tempfile
import types

import numpy as np
import torch

from oracle import oracle as O
from oracle.cases import DEFAULT_PROMPT


def setup_prompt(meta_prompt=DEFAULT_PROMPT, hyper=None, cfg_kw=None, register=True):
    """Product-side equivalent of reference run.setup + parseMetaPrompt, with the whitespace tokenizer."""
    from guided_attention_b200 import run as R
    return R.setup_prompt(meta_prompt, hyper, cfg_kw, register)


def oracle_tokens(cfg):
    from guided_attention_b200.helpers import AnnotationType as AT
    kind = {AT.COOR: O.COOR, AT.BOX: O.BOX, AT.KEYWORD: O.KEYWORD}
    out = []
    for idx, info in cfg.token_dict.items():
        pl = info['loss']
        if info['loss_type'] == AT.BOX:
            pl = (pl.x, pl.y, pl.width, pl.height)
        out.append(O.TokenSpec(index=idx, kind=kind[info['loss_type']], payload=pl, subprompt=info['subprompt'],
                               word=info['word']))
    return out


def oracle_hyper(cfg):
    from guided_attention_b200 import shared_state as S
    hp = S.curHyperParams
    return O.HyperParams(strict=hp["strict"], inside_loss_scale=hp["inside_loss_scale"],
                         outside_loss_scale=hp["outside_loss_scale"], shrink_factor=hp["shrink_factor"],
                         bb_center_weight=hp.get("bb_center_weight", .05),
                         sub_prompt_avg_within=cfg.sub_prompt_avg_within)


def make_layers(res, heads, head_dim, layers, batch, T, seed, gain=1.0):
    """fp32 CPU q/k/v per layer in projection layout (B, N, H*d) / (B, T, H*d)."""
    g = torch.Generator("cpu").manual_seed(seed)
    out = []
    N, C = res * res, heads * head_dim
    for _ in range(layers):
        q = gain * torch.randn(batch, N, C, generator=g)
        k = gain * torch.randn(batch, T, C, generator=g)
        v = torch.randn(batch, T, C, generator=g)
        out.append((q, k, v))
    return out


def oracle_forward(layers_qkv, heads, scale, cfg, res, dtype=torch.float32, last_idx=-1, smooth=True):
    """CPU oracle: explicit softmax attention per layer -> aggregate -> guidance loss; returns result + grads wrt q."""
    qs, Ps, outs = [], [], []
    for q, k, v in layers_qkv:
        q = q.to(dtype).float().requires_grad_(True)   # oracle computes in fp32 on the (rounded) inputs
        k2, v2 = k.to(dtype).float(), v.to(dtype).float()
        p, o = O.cross_attention(O.head_to_batch(q, heads), O.head_to_batch(k2, heads), O.head_to_batch(v2, heads),
                                 scale)
        qs.append(q); Ps.append(p); outs.append(O.batch_to_head(o, heads))
    store = {"down_cross": Ps, "mid_cross": [], "up_cross": []}
    abar = O.aggregate_attention(store, res)
    r = O.guidance_loss(abar, oracle_tokens(cfg), res, oracle_hyper(cfg), smooth_attentions=smooth, last_idx=last_idx)
    return r, qs, Ps, outs, abar


def cuda_forward(layers_qkv, heads, scale, cfg, res, dtype, impl=None, device="cuda:0", normalize_eot=False,
                 smooth=True):
    """Product path: K1 per layer (with accumulation) -> fused tail; returns losses_dict + leaf queries + outputs."""
    from guided_attention_b200 import ops
    from guided_attention_b200.pipeline_guided_attention import GuidedAttention
    from guided_attention_b200.ptp_utils import HeadSummedMaps
    from guided_attention_b200.substrate import WhitespaceTokenizer
    pipe = GuidedAttention(unet=None, tokenizer=cfg.stable.tokenizer)
    pipe.prompt = cfg.prompt
    qs, maps, outs = [], [], []
    for q, k, v in layers_qkv:
        qd = q.to(device=device, dtype=dtype).requires_grad_(True)
        kd, vd = k.to(device=device, dtype=dtype), v.to(device=device, dtype=dtype)
        o, acc = ops.cross_attention(qd, kd, vd, heads, scale, want_acc=True, impl=impl)
        qs.append(qd); outs.append(o)
        maps.append(HeadSummedMaps(acc, heads, qd, kd, scale))
    ld = pipe._compute_max_attention_per_index(maps, smooth_attentions=smooth, sigma=0.5, kernel_size=3,
                                               normalize_eot=normalize_eot)
    loss, losses, unscaled = pipe._compute_loss(ld)
    return ld, loss, losses, unscaled, qs, outs, maps


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rms_rel_err(a, b):
    """||a - b||_2 / ||b||_2: unlike `rel_err` (max-norm, dominated by the largest entries) this weighs every entry."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-30))


def tails_ok(a, b, rtol, atol=1e-7):
    """Element-wise |a - b| <= atol + rtol |b|: a RELATIVE bound on every entry, so the small-probability tails of a
    map are checked as strictly as its peaks.  Returns (ok, worst relative excess) for the assertion message."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    excess = np.abs(a - b) / (atol + rtol * np.abs(b))
    return bool(excess.max() <= 1.0), float(excess.max())


def record_metric(name, values):
    """Appends measured parity numbers to gpurun_out/r02_parity_metrics.jsonl (brought back from the GPU box) so the
    bounds asserted in the tests can be stated next to what was measured."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "r02_parity_metrics.jsonl"), "a") as f:
            f.write(json.dumps({"name": name, **values}) + "\n")
    except OSError:
        pass


def run_microcase(res=16, heads=8, head_dim=40, layers=5, batch=1, dtype=torch.float16, seed=0, impl=None,
                  meta_prompt=DEFAULT_PROMPT, gain=1.0):
    cfg = setup_prompt(meta_prompt)
    scale = head_dim ** -0.5
    lay = make_layers(res, heads, head_dim, layers, batch, 77, seed, gain)
    r, oq, oP, oo, abar = oracle_forward(lay, heads, scale, cfg, res, dtype=dtype)
    og = torch.autograd.grad(r.loss, oq)
    ld, loss, losses, unscaled, qs, outs, maps = cuda_forward(lay, heads, scale, cfg, res, dtype, impl=impl)
    grads = torch.autograd.grad(loss, qs)
    torch.cuda.synchronize()
    info = {
        "loss": float(loss.detach()), "oracle_loss": float(r.loss.detach()),
        "loss_rel_err": abs(float(loss.detach()) - float(r.loss.detach())) / abs(float(r.loss.detach())),
        "out_rel_err": max(rel_err(o.detach().float().cpu().numpy(), oo_.detach().numpy()) for o, oo_ in zip(outs, oo)),
        "grad_rel_err": max(rel_err(g.float().cpu().numpy(), og_.numpy()) for g, og_ in zip(grads, og)),
        "argmax_equal": [int(a) for a in ld["_argmax"].cpu().tolist()] == r.argmax,
    }
    return info
