"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's cross-attention guidance path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module, and only as the checker or the reported CPU baseline.  The product package never imports it.

Every function cites the reference lines it restates (paths relative to /root/reference).  The restatement is
loop-free PyTorch on CPU (fp32 by default, fp64 on request), parametrised in `res` (the reference hard-codes 16:
`16 -> res`, `15 -> res - 1`), and differentiable through torch autograd so that gradients can be compared too.

Parity pinning: the reference ships no tests or golden vectors, so this oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF, produced in the build container by importing the unmodified reference under import stubs
(`oracle/ref_loader.py`) -- see `oracle/gen_golden.py` (generator) and `tests/golden/*.json|npz` (fixtures), checked by
`tests/test_oracle_vs_golden.py`; when /root/reference is mounted, `tests/test_oracle_vs_reference_live.py` also
compares live.  What cannot be pinned: everything owned by diffusers 0.12.1 / CLIP (absent offline) -- those pieces
are *definitions* in `guided_attention_b200/substrate`, shared by oracle and product.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

COOR, BOX, KEYWORD = 0, 1, 2


# ------------------------------------------------------------------------------------------------------ specs
@dataclass
class TokenSpec:
    """One tracked text token (one entry of the reference's `config.token_dict`, run.py:81-91)."""
    index: int                 # position in the tokenised prompt (BOS = 0)
    kind: int                  # COOR / BOX / KEYWORD
    payload: object = None     # (x, y) for COOR, (x, y, w, h) unit-square box for BOX, None for KEYWORD
    subprompt: str = ""
    word: str = ""


@dataclass
class HyperParams:
    """reference utils/shared_state.py:21 + pipeline_guided_attention.py:430 defaults."""
    strict: bool = False
    inside_loss_scale: float = .2
    outside_loss_scale: float = .2
    shrink_factor: float = .15
    bb_center_weight: float = .05
    sub_prompt_avg_within: bool = False


# --------------------------------------------------------------------------------------------- a1: attention
@dataclass
class PaintWithWords:
    """State of the optional paint-with-words score bias (reference utils/ptp_utils.py:113-138): active for 77-key
    layers while `cur_time_step_iter < paint_with_words_stop`.  `sigma` is `state.get_sigma()` =
    sigmas[timesteps[cur_time_step_iter]] (utils/shared_state.py:26-27)."""
    tokens: Sequence["TokenSpec"]
    sigma: float
    weight: float = 1.0           # curHyperParams["paint_with_words_weight"]
    shrink_factor: float = .15    # inside_box reads curHyperParams["shrink_factor"] (utils/helpers.py:168-169)


def pww_mask(pww: PaintWithWords, n_query: int, n_ctx: int = 77) -> torch.Tensor:
    """(n_query, n_ctx) fp32: `weight` where pixel (ii, jj) of the hw x hw grid lies inside the (shrunk) box of a BOX
    token, in that token's column.  reference utils/ptp_utils.py:119-133 (`rect.of_size(hw)`, `inside_box(jj, ii)`)."""
    hw = int(n_query ** .5)
    mask = torch.zeros((hw, hw, n_ctx))
    for tok in pww.tokens:
        if tok.kind == BOX:
            m = inside_box_mask(rect_of_size(tuple(tok.payload), hw), hw, pww.shrink_factor)
            mask[:, :, tok.index] = torch.from_numpy(m.astype(np.float32)) * pww.weight
    return mask.reshape(n_query, n_ctx)


def attention_scores(q: torch.Tensor, k: torch.Tensor, scale: float, attention_mask: Optional[torch.Tensor] = None,
                     pww: Optional[PaintWithWords] = None) -> torch.Tensor:
    """scale * Q K^T (+ attention_mask) (+ mask * 0.4 * max(scores) * ln(1 + sigma)); the max is over the whole
    (B*H, N, T) tensor, taken after the attention_mask was added, and stays in the autograd graph.
    reference utils/ptp_utils.py:103-138."""
    s = scale * torch.bmm(q, k.transpose(-1, -2))
    if attention_mask is not None:
        s = s + attention_mask
    if pww is not None and k.shape[1] == 77:
        m = pww_mask(pww, q.shape[1], k.shape[1]).to(s.device)
        s = s + m[None] * .4 * s.max() * float(np.log(1 + pww.sigma))
    return s


def attention_probs(q: torch.Tensor, k: torch.Tensor, scale: float, attention_mask: Optional[torch.Tensor] = None,
                    pww: Optional[PaintWithWords] = None) -> torch.Tensor:
    """softmax over keys; q (BH, N, d), k (BH, T, d).  reference utils/ptp_utils.py:103-109, 143-144."""
    return torch.softmax(attention_scores(q, k, scale, attention_mask, pww), dim=-1)


def cross_attention(q, k, v, scale, attention_mask=None, pww=None):
    """P and O = P V.  reference utils/ptp_utils.py:82-85."""
    p = attention_probs(q, k, scale, attention_mask, pww)
    return p, torch.bmm(p, v)


def head_to_batch(t: torch.Tensor, heads: int) -> torch.Tensor:
    """(B, S, H*d) -> (B*H, S, d), batch-major (diffusers CrossAttention.head_to_batch_dim, [memory])."""
    b, s, c = t.shape
    return t.reshape(b, s, heads, c // heads).permute(0, 2, 1, 3).reshape(b * heads, s, c // heads)


def batch_to_head(t: torch.Tensor, heads: int) -> torch.Tensor:
    bh, s, d = t.shape
    return t.reshape(bh // heads, heads, s, d).permute(0, 2, 1, 3).reshape(bh // heads, s, d * heads)


class OracleProcessor:
    """Restates `AttendExciteCrossAttnProcessor.__call__` (utils/ptp_utils.py:66-93).  `pww` is a callable returning the
    current `PaintWithWords` state or None (off, the reference's default)."""

    def __init__(self, store, place_in_unet, pww=None):
        self.store = store
        self.place = place_in_unet
        self.pww = pww

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None):
        is_cross = encoder_hidden_states is not None
        ctx = encoder_hidden_states if is_cross else hidden_states
        q = head_to_batch(attn.to_q(hidden_states), attn.heads)
        k = head_to_batch(attn.to_k(ctx), attn.heads)
        v = head_to_batch(attn.to_v(ctx), attn.heads)
        pww = self.pww() if (self.pww is not None and is_cross) else None
        p, o = cross_attention(q, k, v, attn.scale, pww=pww)
        self.store(p, is_cross, self.place)
        return attn.to_out[1](attn.to_out[0](batch_to_head(o, attn.heads)))


# ------------------------------------------------------------------------------------------------- a2: store
class OracleStore:
    """Restates AttentionControl/AttentionStore (utils/ptp_utils.py:178-210, 219-270), save_individual_CA_maps off."""

    KEYS = ("down_cross", "mid_cross", "up_cross", "down_self", "mid_self", "up_self")

    def __init__(self):
        self.num_att_layers = -1
        self.reset()

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0
        self.step_store = {k: [] for k in self.KEYS}
        self.attention_store = {}

    def __call__(self, attn, is_cross, place):
        if attn.shape[1] <= 32 ** 2:
            self.step_store[f"{place}_{'cross' if is_cross else 'self'}"].append(attn)
        self.cur_att_layer += 1
        if self.cur_att_layer == self.num_att_layers:
            self.cur_att_layer = 0
            self.cur_step += 1
            self.attention_store = self.step_store
            self.step_store = {k: [] for k in self.KEYS}


def register(unet, store, pww=None) -> int:
    """Restates `register_attention_control` (utils/ptp_utils.py:149-175)."""
    procs, n = {}, 0
    for name in unet.attn_processors.keys():
        if name.startswith("mid_block"):
            place = "mid"
        elif name.startswith("up_blocks"):
            place = "up"
        elif name.startswith("down_blocks"):
            place = "down"
        else:
            continue
        n += 1
        procs[name] = OracleProcessor(store, place, pww)
    unet.set_attn_processor(procs)
    store.num_att_layers = n
    return n


# --------------------------------------------------------------------------------------------- a3: aggregate
def aggregate_attention(store_dict: Dict[str, List[torch.Tensor]], res: int,
                        from_where: Sequence[str] = ("up", "down", "mid"), is_cross: bool = True) -> torch.Tensor:
    """Mean over layers x (batch*heads) of the stored maps with N == res^2 -> (res, res, T).
    reference utils/ptp_utils.py:273-289 (select = 0)."""
    out = []
    for loc in from_where:
        for item in store_dict[f"{loc}_{'cross' if is_cross else 'self'}"]:
            if item.shape[1] == res * res:
                out.append(item.reshape(-1, res, res, item.shape[-1]))
    out = torch.cat(out, dim=0)   # raises on an empty list exactly like the reference
    return out.sum(0) / out.shape[0]


# ------------------------------------------------------------------------------------------------ a4: renorm
def renorm(abar: torch.Tensor, last_idx: int = -1) -> torch.Tensor:
    """softmax(100 * Abar[:, :, 1:last]) over the text tokens.  reference pipeline_guided_attention.py:209-219."""
    return torch.softmax(abar[:, :, 1:last_idx] * 100, dim=-1)


# --------------------------------------------------------------------------------------------- a5: smoothing
def gaussian_weights_2d(kernel_size: int = 3, sigma: float = 0.5) -> torch.Tensor:
    """The reference's (non-standard exponent) Gaussian, fp32, normalised to sum 1.
    reference utils/gaussian_smoothing.py:21-43:  g(k) = 1/(s*sqrt(2pi)) * exp(-((k-mean)/(2s))^2)."""
    ax = torch.arange(kernel_size, dtype=torch.float32)
    mean = (kernel_size - 1) / 2
    g = 1 / (sigma * math.sqrt(2 * math.pi)) * torch.exp(-((ax - mean) / (2 * sigma)) ** 2)
    k2 = g[:, None] * g[None, :]
    return k2 / k2.sum()


def smooth(img: torch.Tensor, kernel_size: int = 3, sigma: float = 0.5) -> torch.Tensor:
    """reflect-pad by 1 then depth-wise conv.  reference pipeline_guided_attention.py:251-254."""
    w = gaussian_weights_2d(kernel_size, sigma).to(img.dtype)
    x = F.pad(img[None, None], (1, 1, 1, 1), mode="reflect")
    return F.conv2d(x, w[None, None])[0, 0]


# ----------------------------------------------------------------------------------------- a7: mask and boxes
def rect_of_size(box: Tuple[float, float, float, float], res) -> Tuple[float, float, float, float]:
    """reference utils/helpers.py:28-30 (Rect.of_size from size 1): one ratio, four float64 products."""
    ratio = float(res / 1)
    x, y, w, h = box
    return (x * ratio, y * ratio, w * ratio, h * ratio)


def inside_box_mask(box_at_res: Tuple[float, float, float, float], res: int, shrink: float) -> np.ndarray:
    """(res, res) uint8; pixel (row ii, col jj) centre (jj+.5, ii+.5) inside the shrunk box, bounds inclusive,
    Python float64 in the reference's operation order.  reference utils/helpers.py:164-173."""
    x, y, w, h = box_at_res
    off_x = shrink * w
    off_y = shrink * h
    m = np.zeros((res, res), dtype=np.uint8)
    for ii in range(res):
        cy = ii + 0.5
        for jj in range(res):
            cx = jj + 0.5
            if cx >= (x + off_x) and cx <= (x + w - off_x):
                if cy >= (y + off_y) and cy <= (y + h - off_y):
                    m[ii, jj] = 1
    return m


def strict_weights(box_at_res, res: int, mask: np.ndarray) -> np.ndarray:
    """Strict-mode weights, normalised inside / outside.  reference utils/helpers.py:158-161, 175-184, 216-246."""
    x, y, w, h = box_at_res
    cx0, cy0 = x + w / 2.0, y + h / 2.0
    wts = np.ones((res, res), dtype=np.float32)
    for ii in range(res):
        for jj in range(res):
            if mask[ii, jj]:
                px, py = jj + 0.5, ii + 0.5
                dist = math.sqrt(math.pow(2 * (cx0 - px) / w, 2) + math.pow(2 * (cy0 - py) / h, 2)) / math.sqrt(2)
                wts[ii, jj] = np.interp(dist, [0, .333, .666, 1.0], [3, 2.5, 1, .2])
    s_in, s_out = np.float32(0), np.float32(0)
    for ii in range(res):
        for jj in range(res):
            if mask[ii, jj]:
                s_in = np.float32(s_in + wts[ii, jj])
            else:
                s_out = np.float32(s_out + wts[ii, jj])
    m = mask.astype(bool)
    out = wts.copy()
    out[m] = wts[m] / s_in
    out[~m] = wts[~m] / s_out
    return out


def box_losses(p: torch.Tensor, box_at_res, res: int, hp: HyperParams):
    """(inside, outside) for the pixel-normalised map p.  reference utils/helpers.py:215-277."""
    mask_np = inside_box_mask(box_at_res, res, hp.shrink_factor)
    n_in = int(mask_np.sum())
    at_most = 1.0 / n_in          # ZeroDivisionError when the shrunk box covers no pixel centre, like the reference
    mask = torch.from_numpy(mask_np.astype(bool))
    if hp.strict:
        w = torch.from_numpy(strict_weights(box_at_res, res, mask_np)).to(p.dtype)
        inside = (w * 2. * torch.clamp(at_most - p, min=0))[mask].sum()
        outside = (w * torch.clamp(p, min=0))[~mask].sum()
    else:
        inside = 1. - p[mask].sum()
        outside = p[~mask].sum()
    return inside, outside, mask_np


# ------------------------------------------------------------------------------------ a6 + a8: stats and loss
def center_of_mass(p: torch.Tensor, res: int):
    """col = sum (jj+.5) p[ii,jj], row = sum (ii+.5) p[ii,jj].  reference pipeline_guided_attention.py:263-268."""
    coords = torch.arange(res, dtype=p.dtype) + 0.5
    return (p * coords[None, :]).sum(), (p * coords[:, None]).sum()


def centering_loss(center_xy, col, row, res: int):
    """|col - cx*res|/(res-1) + 4 |row - cy*res|/(res-1).  reference pipeline_guided_attention.py:390-395."""
    return (col - center_xy[0] * res).abs() / (res - 1.) + 4. * (row - center_xy[1] * res).abs() / (res - 1.)


@dataclass
class GuidanceResult:
    loss: torch.Tensor                       # scalar, differentiable
    scaled: List[torch.Tensor]               # per token (reference `losses`)
    unscaled: List[torch.Tensor]             # per token (reference `unscaled_losses`)
    max: List[torch.Tensor]
    argmax: List[int]
    sum: List[torch.Tensor]
    col: List[torch.Tensor]
    row: List[torch.Tensor]
    inside: List[torch.Tensor]
    outside: List[torch.Tensor]
    masks: List[Optional[np.ndarray]]
    attention_for_text: torch.Tensor = None  # (res, res, T') after renorm
    smoothed: List[torch.Tensor] = field(default_factory=list)
    custom: Optional[torch.Tensor] = None


def guidance_loss(abar: torch.Tensor, tokens: Sequence[TokenSpec], res: int, hp: HyperParams = HyperParams(),
                  smooth_attentions: bool = True, sigma: float = 0.5, kernel_size: int = 3, last_idx: int = -1,
                  custom_losses=()) -> GuidanceResult:
    """Rows a4-a8: renorm -> per tracked token smoothing, max, centre of mass, box losses -> loss assembly.
    reference pipeline_guided_attention.py:201-296 + :398-451 + :358-387.
    `custom_losses`: callables A -> scalar tensor (reference :286-289, appended under sub-prompt None)."""
    A = renorm(abar, last_idx)
    r = GuidanceResult(loss=None, scaled=[], unscaled=[], max=[], argmax=[], sum=[], col=[], row=[], inside=[],
                       outside=[], masks=[], attention_for_text=A)
    zero = torch.zeros((), dtype=abar.dtype)
    for tk in tokens:
        img = A[:, :, tk.index - 1]
        if smooth_attentions:
            img = smooth(img, kernel_size, sigma)
        r.smoothed.append(img)
        r.max.append(img.max())
        r.argmax.append(int(img.flatten().argmax()))
        s = img.sum()
        r.sum.append(s)
        p = img / s
        col, row = center_of_mass(p, res)
        r.col.append(col)
        r.row.append(row)
        if tk.kind == BOX:
            box = rect_of_size(tk.payload, float(res))
            ins, outs, m = box_losses(p, box, res, hp)
            r.inside.append(ins); r.outside.append(outs); r.masks.append(m)
        else:
            r.inside.append(zero); r.outside.append(zero); r.masks.append(None)
    # loss assembly (reference :398-451)
    for i, tk in enumerate(tokens):
        if tk.kind == COOR:
            li = centering_loss(tk.payload, r.col[i], r.row[i], res)
            r.scaled.append(li); r.unscaled.append(li)
        elif tk.kind == BOX:
            ins, outs = r.inside[i], r.outside[i]
            li = hp.inside_loss_scale * ins + hp.outside_loss_scale * outs * 3
            if hp.bb_center_weight > 0:
                x, y, w, h = tk.payload
                li = li + hp.bb_center_weight * centering_loss((x + w / 2.0, y + h / 2.0), r.col[i], r.row[i], res)
            r.scaled.append(li); r.unscaled.append(ins + outs)
        else:  # KEYWORD contributes nothing here (reference has no branch for it)
            r.scaled.append(None); r.unscaled.append(None)
    custom = zero.clone()
    for fn in custom_losses:
        custom = custom + fn(A)
    r.custom = custom
    r.loss = group_total([(tk, l) for tk, l in zip(tokens, r.scaled) if l is not None], hp) + custom
    return r


def group_by_subprompt(items, hp: HyperParams):
    """{sub-prompt: sum (or mean if sub_prompt_avg_within)}.  reference pipeline_guided_attention.py:358-387."""
    groups: Dict[object, list] = {}
    for tk, val in items:
        groups.setdefault(tk.subprompt if tk is not None else None, []).append(val)
    out = {}
    for key, vals in groups.items():
        tot = 0.
        for v in vals:
            tot = tot + (v / len(vals) if hp.sub_prompt_avg_within else v)
        out[key] = tot
    return out


def group_total(items, hp: HyperParams):
    tot = 0.
    for v in group_by_subprompt(items, hp).values():
        tot = tot + v
    return tot


def meets_threshold(i: int, thresholds: Dict[int, float], tokens, unscaled, custom, hp: HyperParams) -> bool:
    """Every sub-prompt's summed unscaled loss <= thresholds[i] (last value for i == -1); True when i is not a key.
    reference pipeline_guided_attention.py:1074-1088."""
    if (i not in thresholds and i != -1) or len(thresholds) == 0:
        return True
    thresh = list(thresholds.values())[-1] if i == -1 else thresholds[i]
    items = [(tk, u) for tk, u in zip(tokens, unscaled) if u is not None]
    if custom is not None:
        items.append((None, custom))
    for v in group_by_subprompt(items, hp).values():
        if float(v) > thresh:
            return False
    return True


# ------------------------------------------------------------------------------------------ a10: ToLeftOf
def to_left_of(A: torch.Tensor, left_cols: Sequence[int], right_cols: Sequence[int]) -> torch.Tensor:
    """9 * max(0, (cx_left + 0.2 W - cx_right) / W); the right centre is divided by len(left) (reference quirk).
    reference run.py:174-200, 216-224 (16 -> res)."""
    res = A.shape[1]
    coords = torch.arange(res, dtype=A.dtype) + 0.5

    def cx(c):
        m = A[:, :, c]
        return ((m / m.sum()) * coords[None, :]).sum()

    left = sum(cx(c) / len(left_cols) for c in left_cols)
    right = sum(cx(c) / len(left_cols) for c in right_cols)
    return torch.clamp((left + .2 * res - right) / res * 9, min=0)


# --------------------------------------------------------------------------------- a9 + pipeline loop (caller)
def update_latent(latents, loss, step_size):
    """latents - step * d loss / d latents.  reference pipeline_guided_attention.py:455-470."""
    g = torch.autograd.grad(loss, [latents], retain_graph=True)[0]
    return latents - step_size * g


@dataclass
class PipelineTrace:
    latents: torch.Tensor
    losses: List[Tuple[int, int, float]] = field(default_factory=list)   # (step, sub_iteration, loss)
    unet_forwards: int = 0


class OraclePipeline:
    """Restates the guided denoising loop, `GuidedAttention.__call__` (pipeline_guided_attention.py:746-1072) and
    `_perform_iterative_refinement_step` (:475-581), with diagnostics off and `use_optimizer`=False."""

    def __init__(self, unet, scheduler, tokens: Sequence[TokenSpec], hp: HyperParams = HyperParams(),
                 recurse_steps: int = 3, recurse_until: int = 14, custom_losses=(), pww_stop: int = 0,
                 pww_weight: float = 1.0):
        self.unet, self.scheduler = unet, scheduler
        self.tokens, self.hp = list(tokens), hp
        self.recurse_steps, self.recurse_until = max(recurse_steps, 1), recurse_until
        self.custom_losses = custom_losses
        self.store = OracleStore()
        self.pww_stop, self.pww_weight, self._iter = pww_stop, pww_weight, 0
        register(unet, self.store, self._pww_state if pww_stop > 0 else None)
        self.n_fwd = 0

    def _pww_state(self):
        """paint-with-words is on while cur_time_step_iter < paint_with_words_stop (utils/ptp_utils.py:113-114);
        sigma = sigmas[timesteps[i]] with sigmas = sqrt((1 - a_bar) / a_bar) (pipeline :887-890, shared_state.py:26-27)."""
        if self._iter >= self.pww_stop:
            return None
        ac = self.scheduler.alphas_cumprod
        sigmas = (((1 - ac) / ac) ** 0.5).numpy()
        return PaintWithWords(self.tokens, float(sigmas[int(self.scheduler.timesteps[self._iter])]), self.pww_weight,
                              self.hp.shrink_factor)

    def _loss(self, attention_res, smooth_attentions, sigma, kernel_size, last_idx):
        abar = aggregate_attention(self.store.attention_store, attention_res)
        # the loss restatement is host code; when a test drives the UNet on a device (full-size configs whose explicit
        # self-attention maps would take minutes on the host) the averaged map comes back here, autograd-connected
        abar = abar.cpu()
        return guidance_loss(abar, self.tokens, attention_res, self.hp, smooth_attentions, sigma, kernel_size,
                             last_idx, self.custom_losses)

    def _unet(self, x, t, emb):
        self.n_fwd += 1
        return self.unet(x, t, encoder_hidden_states=emb).sample

    def _meets(self, i, thresholds, r):
        return meets_threshold(i, thresholds, self.tokens, r.unscaled, r.custom, self.hp)

    def _refine(self, latents, emb, step_size, t, i, thresholds, lk, trace, max_refinement_steps=10):
        it, r = 0, None
        while r is None or not self._meets(i, thresholds, r):
            it += 1
            latents = latents.clone().detach().requires_grad_(True)
            self._unet(latents, t, emb[1][None])
            r = self._loss(**lk)
            trace.losses.append((i, it, float(r.loss)))
            if float(r.loss) != 0:
                latents = update_latent(latents, r.loss, step_size)
            if it >= max_refinement_steps:
                break
        latents = latents.clone().detach().requires_grad_(True)
        self._unet(latents, t, emb[1][None])
        r = self._loss(**lk)
        return r, latents

    @torch.no_grad()
    def __call__(self, prompt_embeds: torch.Tensor, latents: torch.Tensor, seed: int, attention_res: int = 16,
                 num_inference_steps: int = 50, guidance_scale: float = 7.5, max_iter_to_alter: int = 25,
                 run_standard_sd: bool = False, thresholds: Dict[int, float] = None, scale_factor: int = 20,
                 scale_range=(1., .5), smooth_attentions=True, sigma=0.5, kernel_size=3, last_idx=-1,
                 only_update_on_threshold_steps=True, max_steps: Optional[int] = None,
                 max_refinement_steps: int = 10) -> PipelineTrace:
        thresholds = {0: float("inf")} if not thresholds else thresholds
        sch = self.scheduler
        sch.set_timesteps(num_inference_steps)
        scale = np.linspace(scale_range[0], scale_range[1], len(sch.timesteps))
        renoise_gen = torch.Generator("cpu").manual_seed(seed) if self.recurse_steps > 1 else None
        lk = dict(attention_res=attention_res, smooth_attentions=smooth_attentions, sigma=sigma,
                  kernel_size=kernel_size, last_idx=last_idx)
        trace = PipelineTrace(latents=latents)
        for i, t in enumerate(sch.timesteps):
            if max_steps is not None and i >= max_steps:
                break
            for recurse_step in range(self.recurse_steps):
                updated = False
                self._iter = i                      # state.cur_time_step_iter = i (:928)
                with torch.enable_grad():
                    latents = latents.clone().detach().requires_grad_(True)
                    self._unet(latents, t, prompt_embeds[1][None])
                    r = self._loss(**lk)
                    if not run_standard_sd:
                        trace.losses.append((i, 0, float(r.loss)))
                        stale = r   # the reference keeps testing the PRE-refinement unscaled losses below (:1001)
                        step_size = scale_factor * np.sqrt(scale[i])
                        if not self._meets(i, thresholds, r):
                            updated = True
                            r, latents = self._refine(latents, prompt_embeds, step_size, t, i, thresholds, lk, trace,
                                                      max_refinement_steps)
                        if (not only_update_on_threshold_steps and i < max_iter_to_alter) or (i in thresholds):
                            if not self._meets(-1, thresholds, stale):
                                updated = True
                                if float(r.loss) != 0:
                                    latents = update_latent(latents, r.loss, step_size)
                x2 = sch.scale_model_input(torch.cat([latents] * 2), t)
                noise = self._unet(x2, t, prompt_embeds)
                n_u, n_t = noise.chunk(2)
                noise = n_u + guidance_scale * (n_t - n_u)
                latents = sch.step(noise, t, latents).prev_sample
                if i > self.recurse_until or not updated:
                    break
                if recurse_step != self.recurse_steps - 1:
                    ti = int(t)
                    prev_t = ti - sch.config.num_train_timesteps // sch.num_inference_steps
                    if prev_t > 0:
                        bt = sch.alphas_cumprod[ti] / sch.alphas_cumprod[prev_t]
                        latents = bt.sqrt() * latents + (1 - bt).sqrt() * torch.randn(
                            latents.shape, generator=renoise_gen).to(latents.device)
        trace.latents = latents.detach()
        trace.unet_forwards = self.n_fwd
        return trace
