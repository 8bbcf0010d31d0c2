"""TEST INFRASTRUCTURE ONLY -- seeded case definitions shared by the golden generator (`oracle/gen_golden.py`, runs the
reference) and by the tests (which re-draw the same inputs and compare oracle / CUDA results with the fixtures).
Nothing here reads /root/reference.
"""
from __future__ import annotations

import torch

DEFAULT_PROMPT = 'a [robot:.6,.3,.4,.55] and a [blue vase:.2,.3,.4,.55]'

# ------------------------------------------------------------------------------------------------ parser cases
PARSE_CASES = [
    DEFAULT_PROMPT,
    'a [cat:.3,.6] and a [dog:.55,.2,.4,.5] on grass',
    'a photo of a [red sports car:.1,.4,.5,.4] near a [tree:.7,.1,.25,.8]',
    '[robot:.1,.2,.3,.4] vase',                      # trailing bare word is dropped
    'vase',                                          # single word -> empty prompt
    'a  [ robot : .6 , .3 , .4 , .55 ]  and  more',  # blanks inside the bracket
    'a [robot:.6,.3,.4] here',                       # 3 numbers: word kept, no annotation
    'x[robot:.5,.5] y z',                            # bracket glued to a word
    'a [big [nested] robot:.6,.3,.4,.55] and it',    # nested brackets
    'a [robot:.6,.3,.4,.55',                         # unmatched bracket
    'a [robot .6,.3,.4,.55] b',                      # no colon -> ValueError
    'a [cat:.2,.3,.3,.4] and a dog with a bird [CustomLoss:toLeftOf (dog,bird)]',
    'a [cat:.2,.3,.3,.4] left of a [dog:.6,.3,.3,.4] [CustomLoss:toLeftOf (cat,dog)]',  # keyword overwrites box
    'a [cat:.2,.3,.3,.4] [CustomLoss:unknown (cat,dog)]',   # KeyError
    'a [cat:a,b] x',                                 # ValueError from float()
    'time: a [robot:.6,.3,.4,.55] now',              # a colon before the bracket wins
]

def fuzz_parse_cases(n: int = 400, seed: int = 20261018):
    """Seeded random meta-prompts from the bracket grammar: plain words, [words:x,y] crosshairs,
    [words:x,y,w,h] boxes, 1/3/5-number payloads, blanks, nested and unmatched brackets, missing colons, non-numeric
    payloads, glued brackets, repeated sub-prompts and a trailing [CustomLoss:...] -- including the malformed ones: the
    error TYPE the reference raises is part of the fixture."""
    import random
    rng = random.Random(seed)
    words = ["a", "the", "robot", "vase", "cat", "dog", "bird", "blue", "red", "big", "car", "tree", "on", "and",
             "of", "grass", "photo", "near", "left", "table"]

    def num():
        r = rng.random()
        if r < .08:
            return rng.choice(["a", "", "1e-1", "-.2", "1.", "0", "1"])
        v = rng.random()
        return rng.choice(["%.2f" % v, ("%.2f" % v).lstrip("0") or "0", "%.3f" % v, repr(round(v, 1))])

    def phrase(k):
        return " ".join(rng.choice(words) for _ in range(k))

    def bracket():
        body = phrase(rng.randint(1, 3))
        if rng.random() < .06:
            body = body + " [" + phrase(1) + "]"                   # nested bracket inside the sub-prompt
        n_num = rng.choice([2, 2, 4, 4, 4, 4, 1, 3, 5])
        sep = rng.choice([",", ",", ", ", " ,"])
        payload = sep.join(num() for _ in range(n_num))
        colon = ":" if rng.random() > .05 else rng.choice(["", " ", ";"])
        pad = rng.choice(["", "", " "])
        close = "]" if rng.random() > .04 else ""
        return "[" + pad + body + pad + colon + pad + payload + pad + close

    cases = []
    for _ in range(n):
        parts = []
        for _ in range(rng.randint(1, 5)):
            r = rng.random()
            if r < .45:
                parts.append(phrase(rng.randint(1, 3)))
            else:
                parts.append(bracket())
        glue = " " if rng.random() > .08 else ""
        text = glue.join(parts) if glue == "" else " ".join(parts)
        if rng.random() < .12:
            args = ",".join(rng.choice(words) for _ in range(2))
            name = rng.choice(["toLeftOf", "toLeftOf", "unknown"])
            text += " [CustomLoss:" + name + " (" + args + ")]"
        if rng.random() < .05:
            text = "  " + text + " "
        cases.append(text)
    return cases


# ------------------------------------------------------------------------------------------------- mask cases
# (unit box, res, shrink)
MASK_CASES = [
    ((.6, .3, .4, .55), 16, .15), ((.2, .3, .4, .55), 16, .15),
    ((.6, .3, .4, .55), 32, .15), ((.2, .3, .4, .55), 32, .15),
    ((.6, .3, .4, .55), 24, .15), ((.6, .3, .4, .55), 64, .15), ((.6, .3, .4, .55), 8, .15),
    ((.6, .3, .4, .55), 16, 0.0), ((.1, .1, .8, .8), 16, .25), ((0., 0., 1., 1.), 16, 0.),
    ((.25, .25, .5, .5), 16, 0.), ((.25, .25, .5, .5), 32, .125), ((.3, .3, .3, .3), 16, .1),
    ((.5, .5, .02, .02), 16, .15),   # no inside pixel
    ((.09375, .09375, .8125, .8125), 16, 0.),  # edges exactly on pixel centres (1.5 .. 14.5)
    ((.7, .1, .25, .8), 96, .15), ((.05, .45, .9, .1), 12, .15),
]

# ------------------------------------------------------------------------------------------------- loss cases
_PLACES5 = ["down", "down", "up", "up", "up"]
LOSS_CASES = [
    dict(name="kat_default", meta_prompt=DEFAULT_PROMPT, seed=0, bh=8, layers=5, places=_PLACES5, gain=1.0),
    dict(name="peaky", meta_prompt=DEFAULT_PROMPT, seed=1, bh=8, layers=5, places=_PLACES5, gain=4.0),
    dict(name="strict", meta_prompt=DEFAULT_PROMPT, seed=2, bh=8, layers=5, places=_PLACES5, gain=4.0,
         hyper={"strict": True}),
    dict(name="coor_mixed", meta_prompt='a [cat:.3,.6] and a [dog:.55,.2,.4,.5] on grass', seed=3, bh=8, layers=5,
         places=_PLACES5, gain=3.0),
    dict(name="avg_within", meta_prompt='a photo of a [red sports car:.1,.4,.5,.4] near a [tree:.7,.1,.25,.8]',
         seed=4, bh=8, layers=5, places=_PLACES5, gain=3.0, cfg={"sub_prompt_avg_within": True}),
    dict(name="no_center", meta_prompt=DEFAULT_PROMPT, seed=5, bh=8, layers=5, places=_PLACES5, gain=2.0,
         hyper={"bb_center_weight": 0}),
    dict(name="custom_left", meta_prompt='a [cat:.2,.3,.3,.4] and a dog with a bird [CustomLoss:toLeftOf (dog,bird)]',
         seed=6, bh=8, layers=5, places=_PLACES5, gain=3.0),
    dict(name="keyword_overwrites_box",
         meta_prompt='a [cat:.2,.3,.3,.4] left of a [dog:.6,.3,.3,.4] [CustomLoss:toLeftOf (cat,dog)]',
         seed=7, bh=8, layers=5, places=_PLACES5, gain=3.0),
    dict(name="batch2_mid", meta_prompt=DEFAULT_PROMPT, seed=8, bh=16, layers=4, places=["down", "mid", "up", "up"],
         gain=3.0),
    dict(name="normalize_eot", meta_prompt=DEFAULT_PROMPT, seed=9, bh=8, layers=5, places=_PLACES5, gain=3.0,
         normalize_eot=True),
    dict(name="no_smooth", meta_prompt=DEFAULT_PROMPT, seed=10, bh=8, layers=5, places=_PLACES5, gain=3.0,
         smooth=False),
    dict(name="loose_threshold", meta_prompt=DEFAULT_PROMPT, seed=11, bh=8, layers=5, places=_PLACES5, gain=6.0,
         thresholds={0: 3.0, 4: 1.0}),
]


def make_loss_inputs(case, res=16, T=77):
    """layers x softmax(gain * randn(bh, res^2, T)) drawn in order from torch.manual_seed(seed) on CPU, plus a
    low-frequency spatial bias so the maps are not exchangeable across pixels."""
    g = torch.Generator("cpu").manual_seed(case["seed"])
    Ps = []
    yy, xx = torch.meshgrid(torch.arange(res, dtype=torch.float32), torch.arange(res, dtype=torch.float32),
                            indexing="ij")
    for _ in range(case["layers"]):
        logits = case["gain"] * torch.randn(case["bh"], res * res, T, generator=g)
        if case["gain"] != 1.0:
            phase = torch.rand(T, generator=g) * 6.283
            bias = torch.sin(xx.reshape(-1, 1) * 0.4 + phase[None, :]) + torch.cos(yy.reshape(-1, 1) * 0.3 + phase)
            logits = logits + case["gain"] * 0.5 * bias[None]
        Ps.append(torch.softmax(logits, dim=-1))
    checksum = float(sum(float(P.double().sum()) + float((P.double() ** 2).sum()) for P in Ps))
    return Ps, checksum


# -------------------------------------------------------------------------------------------- processor case
PROCESSOR_CASE = dict(seed=123, query_dim=64, cross_dim=48, heads=4, dim_head=16, batch=2, n=64, T=77)


def make_processor_inputs(case):
    from guided_attention_b200.substrate import CrossAttention
    state = torch.random.get_rng_state()
    torch.manual_seed(case["seed"])
    attn = CrossAttention(case["query_dim"], case["cross_dim"], case["heads"], case["dim_head"])
    attn_self = CrossAttention(case["query_dim"], None, case["heads"], case["dim_head"])
    x = torch.randn(case["batch"], case["n"], case["query_dim"])
    ctx = torch.randn(case["batch"], case["T"], case["cross_dim"])
    torch.random.set_rng_state(state)
    for p in list(attn.parameters()) + list(attn_self.parameters()):
        p.requires_grad_(False)
    return attn, attn_self, x, ctx


# paint-with-words (utils/ptp_utils.py:113-138): the reference processor on a 16x16 cross layer at denoising iteration
# `iter` < `stop`, plus the gradient of a seeded linear functional of (output, stored probabilities) w.r.t. the
# hidden states -- it exercises the gradient through the global max of the scores.
PWW_CASE = dict(seed=321, query_dim=64, cross_dim=48, heads=4, dim_head=16, batch=2, n=256, T=77,
                meta_prompt=DEFAULT_PROMPT, stop=3, iter=1, weight=1.5, steps=50, functional_seed=77)


def pww_functional(case, y, probs):
    """L = <y, R1> + <sum_h P[b, h], R2[b]> with seeded R1, R2: linear in the layer output and in the head-summed maps
    (what the AttentionStore accumulators hold), so the product path can evaluate the same functional."""
    g = torch.Generator("cpu").manual_seed(case["functional_seed"])
    B, H = case["batch"], case["heads"]
    r1 = torch.randn(y.shape, generator=g).to(y)
    r2 = torch.randn((B,) + tuple(probs.shape[1:]), generator=g).to(probs)
    head_sum = probs.reshape((B, H) + tuple(probs.shape[1:])).sum(1)
    return (y * r1).sum() + (head_sum * r2).sum(), r1, r2


# -------------------------------------------------------------------------------------------------- e2e case
E2E_CASE = dict(meta_prompt=DEFAULT_PROMPT, unet_seed=0, embed_seed=1234, latent_seed=28, steps=5,
                guidance_scale=7.5, embed_gain=4.0,
                hyper={"strict": False, "inside_loss_scale": .2, "outside_loss_scale": .2, "shrink_factor": .15,
                       "thresholds": {0: 3.46, 2: 3.45}, "use_optimizer": False, "recurse_until": 1,
                       "recurse_steps": 2})


def make_e2e_inputs(case):
    from guided_attention_b200.substrate import UNetConfig, build_unet
    cfg = UNetConfig.tiny()
    unet = build_unet(cfg, seed=case["unet_seed"])
    embeds = case["embed_gain"] * torch.randn(2, 77, cfg.cross_attention_dim,
                                              generator=torch.Generator("cpu").manual_seed(case["embed_seed"]))
    gen = torch.Generator("cpu").manual_seed(case["latent_seed"])
    latents = torch.randn(1, 4, 64, 64, generator=gen)
    return unet, embeds, latents, torch.Generator("cpu").manual_seed(case["latent_seed"])
