"""Attention hooks -- drop-in mirror of the reference's `utils/ptp_utils.py:59-289`.

Same public names, signatures and bookkeeping (`register_attention_control`, `AttendExciteCrossAttnProcessor`,
`AttentionControl`, `EmptyControl`, `AttentionStore`, `aggregate_attention`), different machinery underneath:

  * for cross-attention layers the processor calls ONE fused sm_100a kernel (K1, `ops.cross_attention`) that computes
    softmax(scale QK^T)V against the 77-token context and, for layers the store keeps (N <= 32^2), writes the head-sum of
    the probabilities straight into that layer's fp32 accumulator.  The (B*H, N, 77) probability tensor the reference
    materialises and retains for autograd (`utils/ptp_utils.py:82-85`) never exists; the backward (K2) recomputes it
    from the saved row log-sum-exp and injects the map gradient.
  * what the controller receives for a cross layer is therefore a `HeadSummedMaps` handle instead of a probability
    tensor.  It has the `.shape` the reference tensor would have, so `AttentionStore.forward`'s size test is unchanged,
    and `.probs()` materialises the per-head maps on demand (API compatibility, off the hot path).
  * self-attention layers are off the guidance path but on the autograd path: with 16-bit operands they run the fused
    exact self-attention kernels (`ops.self_attention`, forward + backward, no (B*H, N, N) tensor); with fp32 operands,
    or when `AttentionStore(save_self_attention=True)` asks for the maps, they run the reference's explicit math in
    PyTorch (the reference stores the self maps unconditionally although nothing reads them:
    `pipeline_guided_attention.py:309` hard-wires the reader off).
"""
from __future__ import annotations

import abc
import os
from typing import List

import torch
import torch.nn.functional as F

from . import ops
from . import shared_state as state


class HeadSummedMaps:
    """Handle to one cross-attention layer's maps for one UNet forward.

    acc   (B, N, T) fp32, acc[b] = sum_h P[b, h]; autograd-connected to the layer's queries through K1/K2.
    shape (B*H, N, T): the shape of the probability tensor the reference hands to the controller.
    """

    _cache = None

    def __init__(self, acc: torch.Tensor, heads: int, q: torch.Tensor, k: torch.Tensor, scale: float, bias=None):
        self.acc, self.heads, self.scale = acc, heads, scale
        self._q, self._k, self._bias = q, k, bias
        self.shape = torch.Size((acc.shape[0] * heads, acc.shape[1], acc.shape[2]))
        self.dtype, self.device = q.dtype, acc.device

    @property
    def n_maps(self) -> int:
        return self.shape[0]

    def probs(self) -> torch.Tensor:
        """Per-head probabilities (B*H, N, T), recomputed by `ga_attn_probs`; detached."""
        return ops.attention_probs(self._q.detach(), self._k.detach(), self.heads, self.scale, bias=self._bias)

    # -- tensor-like on demand ("lazily materialised only if someone asks").  Code written against the
    # reference's store -- e.g. its own aggregate_attention, `item.reshape(1, -1, res, res, item.shape[-1])`
    # (utils/ptp_utils.py:286), or `torch.cat(out, dim=0)` (:287) -- sees the (B*H, N, T) probability tensor: any
    # attribute this handle does not have, and any torch function it is passed to, goes to the materialised maps.
    def _materialised(self) -> torch.Tensor:
        if self._cache is None:
            self._cache = self.probs()
        return self._cache

    def __getattr__(self, name):
        if name.startswith("__") or name in ("acc", "heads", "scale", "_q", "_k", "_bias", "shape", "dtype", "device"):
            raise AttributeError(name)
        return getattr(self._materialised(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def conv(x):
            if isinstance(x, HeadSummedMaps):
                return x._materialised()
            if isinstance(x, (list, tuple)):
                return type(x)(conv(y) for y in x)
            return x
        return func(*conv(args), **{k: conv(v) for k, v in (kwargs or {}).items()})

    def __getitem__(self, idx):
        return self._materialised()[idx]

    def __len__(self):
        return self.shape[0]

    def mean_map(self) -> torch.Tensor:
        """(N, T) mean over batch x heads, differentiable."""
        return self.acc.sum(0) / self.n_maps

    def detach(self):
        return HeadSummedMaps(self.acc.detach(), self.heads, self._q.detach(), self._k.detach(), self.scale, self._bias)


class AttendExciteCrossAttnProcessor:
    """Same call contract as the reference processor (utils/ptp_utils.py:59-93)."""

    def __init__(self, attnstore, place_in_unet):
        super().__init__()
        self.attnstore = attnstore
        self.place_in_unet = place_in_unet

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None):
        batch_size, sequence_length, _ = hidden_states.shape
        attention_mask = attn.prepare_attention_mask(attention_mask, sequence_length)
        is_cross = encoder_hidden_states is not None
        query = attn.to_q(hidden_states)

        if is_cross:
            key, value = self._text_kv(attn, encoder_hidden_states)
            keep = self.attnstore.wants_maps(sequence_length) if hasattr(self.attnstore, "wants_maps") \
                else sequence_length <= 32 ** 2
            bias = self._score_bias(attention_mask, sequence_length, key.shape[1], query.device)
            out, acc = ops.cross_attention(query, key, value, attn.heads, attn.scale, want_acc=keep, bias=bias)
            if keep:
                maps = HeadSummedMaps(acc, attn.heads, query, key, attn.scale, bias)
            else:
                maps = _ShapeOnly(batch_size * attn.heads, sequence_length, key.shape[1])
            self.attnstore(maps, True, self.place_in_unet)
            hidden_states = out
        else:
            key = attn.to_k(hidden_states)
            value = attn.to_v(hidden_states)
            want_maps = bool(getattr(self.attnstore, "save_self_attention", False))
            fused = (not want_maps and attention_mask is None and hidden_states.is_cuda
                     and ops.self_attention_supported(query.dtype, query.shape[-1] // attn.heads)
                     and not (attn.upcast_attention or attn.upcast_softmax))
            if fused:
                # self-attention, 16-bit operands: fused exact attention; the (B*H, N, N) probabilities
                # the reference materialises and keeps for autograd never exist.  The controller still gets its call
                # (layer counting, utils/ptp_utils.py:194-201) with a shape-only stand-in: nothing reads these maps
                # (pipeline_guided_attention.py:309 hard-wires the reader off).
                hidden_states = ops.self_attention(query, key, value, attn.heads, attn.scale)
                self.attnstore(_ShapeOnly(batch_size * attn.heads, sequence_length, sequence_length), False,
                               self.place_in_unet)
            else:
                # fp32 operands / stored self maps: the reference's explicit math (utils/ptp_utils.py:77-85) in PyTorch
                query = attn.head_to_batch_dim(query)
                key = attn.head_to_batch_dim(key)
                value = attn.head_to_batch_dim(value)
                attention_probs = attn.get_attention_scores(query, key, attention_mask)
                self.attnstore(attention_probs, False, self.place_in_unet)
                hidden_states = attn.batch_to_head_dim(torch.bmm(attention_probs, value))

        hidden_states = attn.to_out[0](hidden_states)
        hidden_states = attn.to_out[1](hidden_states)
        return hidden_states


def _score_bias_method(self, attention_mask, n_query, n_ctx, device):
    """The two optional additive score terms of the reference processor (utils/ptp_utils.py:113-138), handed to the
    kernels as one `ops.ScoreBias`: the call's `attention_mask` (added as is, broadcast like `scores + mask`) and the
    paint-with-words bias.  None when neither applies (the default: the tcgen05 kernels run)."""
    pww = PaintWithWords.of(self.attnstore)
    use_pww = pww is not None and n_ctx == 77 and pww.n_tokens() > 0
    if attention_mask is None and not use_pww:
        return None
    if not use_pww:
        return ops.ScoreBias(mask=attention_mask)
    pww.update(device)
    masks, columns = pww.masks_for(n_query, device)
    return ops.ScoreBias(mask=attention_mask, pww_masks=masks, pww_columns=columns, pww_coef=pww.coef)


class PaintWithWords:
    """Host-side state of the paint-with-words bias (reference utils/ptp_utils.py:113-138), one per controller.

    The reference switches the bias on while `cur_time_step_iter < paint_with_words_stop` and scales it with
    ln(1 + sigma_t); both depend on the denoising step, which a captured CUDA graph cannot see.  So whenever the feature
    is enabled (`paint_with_words_stop` > 0) every 77-key layer runs the biased kernels and the step-dependent factor
    `w * 0.4 * ln(1 + sigma_t)` lives in a one-element DEVICE tensor (`coef`): 0 past `stop` (S + 0 * max(S) == S
    exactly).  `update()` refreshes it from `shared_state`; the graph runner calls it before every replay."""

    def __init__(self):
        self.coef, self._value, self._masks = None, None, {}

    @staticmethod
    def of(controller):
        hp = state.curHyperParams or {}
        if not hp.get("paint_with_words_stop", 0):
            return None
        if getattr(controller, "pww", None) is None:
            controller.pww = PaintWithWords()
        return controller.pww

    @staticmethod
    def n_tokens():
        from .helpers import AnnotationType
        td = getattr(state.config, "token_dict", None) or {}
        return sum(1 for v in td.values() if v['loss_type'] == AnnotationType.BOX)

    @staticmethod
    def current_coef() -> float:
        import numpy as np
        hp = state.curHyperParams
        i = state.cur_time_step_iter
        if i is None or not (i < hp.get("paint_with_words_stop", 0)):
            return 0.0
        return float(hp.get("paint_with_words_weight", 1.0) * .4 * np.log(1 + state.get_sigma()))

    def update(self, device):
        val = self.current_coef()
        if self.coef is None or self.coef.device != torch.device(device):
            self.coef = torch.zeros(1, dtype=torch.float32, device=device)
            self._value = 0.0
        if val != self._value:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("paint-with-words coefficient changed during CUDA-graph capture: call "
                                   "PaintWithWords.update() before capturing / replaying")
            self.coef.fill_(val)
            self._value = val

    def masks_for(self, n_query: int, device):
        """((n_box, n_query) uint8 device masks at this layer's resolution, token-index column of each) -- rasterised
        by K5 exactly like `helpers.inside_box(jj, ii, rect.of_size(hw))` (utils/ptp_utils.py:119-131), once per
        (token table, resolution)."""
        from .helpers import AnnotationType
        hw = int(n_query ** .5)
        shrink = float(state.curHyperParams["shrink_factor"])
        boxes, columns = [], []
        for idx, info in state.config.token_dict.items():
            if info['loss_type'] == AnnotationType.BOX:
                r = info['loss']
                boxes.append((r.x / r.size, r.y / r.size, r.width / r.size, r.height / r.size)
                             if r.size != 1 else (r.x, r.y, r.width, r.height))
                columns.append(idx)
        key = (tuple(boxes), tuple(columns), hw, n_query, shrink, str(device))
        if key not in self._masks:
            if len(self._masks) > 64:
                self._masks.clear()
            m = torch.zeros((len(boxes), n_query), dtype=torch.uint8, device=device)
            m[:, :hw * hw] = ops.rasterize_boxes(boxes, hw, shrink, device).reshape(len(boxes), hw * hw)
            self._masks[key] = m
        return self._masks[key], columns


def _text_kv_method(self, attn, encoder_hidden_states):
    """K and V of a cross-attention layer.  They depend on the text embeddings and the layer's weights only, yet the
    reference recomputes them in every one of the ~500 UNet passes of an image (utils/ptp_utils.py:72-75).  When the
    caller owns the embedding buffer for the lifetime of the run and says so by installing `controller.text_kv`
    (`TextKVCache`; the pipeline's CUDA-graph programs do, the eager path does not), the projections are computed once
    per (layer, buffer) and refreshed in place when the buffer's contents change."""
    cache = getattr(self.attnstore, "text_kv", None)
    if cache is None or encoder_hidden_states.requires_grad or torch.is_grad_enabled() and (
            attn.to_k.weight.requires_grad or attn.to_v.weight.requires_grad):
        return attn.to_k(encoder_hidden_states), attn.to_v(encoder_hidden_states)
    return cache.get(attn, encoder_hidden_states)


AttendExciteCrossAttnProcessor._text_kv = _text_kv_method
AttendExciteCrossAttnProcessor._score_bias = _score_bias_method


class TextKVCache:
    """(layer, embedding buffer) -> (K, V).  Entries keep the tensors they were computed from, so `refresh()` can
    recompute them IN PLACE after the owner overwrote the buffer (captured CUDA graphs keep reading the same memory)."""

    def __init__(self):
        self.entries = {}

    def get(self, attn, ehs):
        key = (id(attn), ehs.data_ptr(), tuple(ehs.shape), ehs.dtype)
        e = self.entries.get(key)
        if e is None:
            with torch.no_grad():
                e = (attn, ehs, attn.to_k(ehs), attn.to_v(ehs))
            self.entries[key] = e
        return e[2], e[3]

    @torch.no_grad()
    def refresh(self):
        for attn, ehs, k, v in self.entries.values():
            k.copy_(attn.to_k(ehs))
            v.copy_(attn.to_v(ehs))


class _ShapeOnly:
    """Stand-in handed to the controller for layers whose maps are not kept (N > 32^2): only `.shape` is meaningful."""

    def __init__(self, *shape):
        self.shape = torch.Size(shape)


def register_attention_control(model, controller):
    """Installs the processor on every attention layer (self and cross) and sets `controller.num_att_layers`
    (reference utils/ptp_utils.py:149-175; 32 for an SD-1.x UNet)."""
    attn_procs = {}
    count = 0
    for name in model.unet.attn_processors.keys():
        if name.startswith("mid_block"):
            place = "mid"
        elif name.startswith("up_blocks"):
            place = "up"
        elif name.startswith("down_blocks"):
            place = "down"
        else:
            continue
        count += 1
        attn_procs[name] = AttendExciteCrossAttnProcessor(attnstore=controller, place_in_unet=place)
    model.unet.set_attn_processor(attn_procs)
    controller.num_att_layers = count
    # The guidance path differentiates with respect to the latents only (pipeline_guided_attention.py:466); a UNet
    # fresh from `from_pretrained` still has requires_grad=True on its weights, which would make every cross layer's
    # K/V "need" gradients: K2 would leave the tcgen05 path to compute dK/dV nobody reads, and the text K/V cache would
    # be bypassed.  Freezing the weights does not change any value the reference computes.
    if hasattr(model.unet, "requires_grad_"):
        model.unet.requires_grad_(False)
    register_fused_norms(model.unet)


def _fused_silu_norm(norm, x):
    if ops.group_norm_supported(x, norm.weight, norm.bias, norm.num_groups):
        return ops.group_norm(x, norm.weight, norm.bias, norm.num_groups, norm.eps, silu=True)
    return F.silu(norm(x))


def _frozen(p) -> bool:
    return p is not None and not p.requires_grad


def _plain_conv(conv) -> bool:
    return (conv.bias is not None and not conv.weight.requires_grad and not conv.bias.requires_grad and conv.groups == 1
            and tuple(conv.dilation) == (1, 1) and conv.padding_mode == "zeros" and not isinstance(conv.padding, str))


def _shift_bias(block):
    """conv1.bias + time_emb_proj.bias of a ResNet block (both frozen), cached until either tensor is modified."""
    tb, cb = block.time_emb_proj.bias, block.conv1.bias
    key = (tb.data_ptr(), tb._version, cb.data_ptr(), cb._version, tb.dtype)
    hit = getattr(block, "_ga_shift_bias", None)
    if hit is None or hit[0] != key:
        fresh = tb.detach() + cb.detach()
        if hit is not None and hit[1].shape == fresh.shape and hit[1].dtype == fresh.dtype and hit[1].device == fresh.device:
            hit[1].copy_(fresh)          # in place: captured CUDA graphs hold this buffer's address
            fresh = hit[1]
        hit = (key, fresh)
        block._ga_shift_bias = hit
    return hit[1]


class _TembShifts:
    """All ResNet blocks of one UNet read the same time embedding: their `time_emb_proj(silu(temb))` projections (22
    launches of a one-row GEMM plus 22 SiLUs per UNet pass in the stock forward) are ONE GEMM against the concatenated
    weights, computed by the first block that sees a new `temb` and cached on that tensor; conv1's bias is folded into
    the concatenated bias, and every block's norm2 kernel reads its column slice as `shift`."""

    def __init__(self, blocks):
        self.blocks = list(blocks)
        self.key = None
        self.weight = self.bias = None
        self.offsets = {}

    def _params(self):
        return [p for b in self.blocks for p in (b.time_emb_proj.weight, b.time_emb_proj.bias, b.conv1.bias)]

    def _build(self):
        key = tuple((p.data_ptr(), p._version, p.dtype) for p in self._params())
        if key != self.key:
            with torch.no_grad():
                weight = torch.cat([b.time_emb_proj.weight for b in self.blocks], dim=0).contiguous()
                bias = torch.cat([b.time_emb_proj.bias + b.conv1.bias for b in self.blocks], dim=0).contiguous()
                # captured CUDA graphs hold the ADDRESSES of these buffers: refresh them in place whenever possible
                if (self.weight is not None and self.weight.shape == weight.shape and self.weight.dtype == weight.dtype
                        and self.weight.device == weight.device):
                    self.weight.copy_(weight)
                    self.bias.copy_(bias)
                else:
                    self.weight, self.bias = weight, bias
            off = 0
            for b in self.blocks:
                self.offsets[id(b)] = (off, b.conv1.out_channels)
                off += b.conv1.out_channels
            self.key = key

    def shift(self, block, temb):
        cached = getattr(temb, "_ga_shifts", None)
        if cached is None or cached[0] is not self:
            self._build()
            cached = (self, F.linear(F.silu(temb), self.weight, self.bias))
            temb._ga_shifts = cached
        off, c = self.offsets[id(block)]
        return cached[1][:, off:off + c]


def _fused_resnet_forward(block, x, temb):
    """ResnetBlock2D with its elementwise work folded away (same mathematics as the stock forward):
        h   = conv1_nobias(silu(norm1(x)))
        h   = conv2_nobias(silu(norm2(h + shift)))       shift[n, c] = conv1.bias + time_emb_proj(silu(temb))
        out = h + conv2.bias + shortcut(x)               one vectorised pass
    i.e. 4 broadcast / residual add launches fewer per block and direction than PyTorch issues."""
    n1, n2, c1, c2 = block.norm1, block.norm2, block.conv1, block.conv2
    if not (ops.group_norm_supported(x, n1.weight, n1.bias, n1.num_groups) and _plain_conv(c1) and _plain_conv(c2)
            and c2.out_channels % 8 == 0 and not temb.requires_grad and not block.time_emb_proj.weight.requires_grad):
        return None
    h = F.conv2d(ops.group_norm(x, n1.weight, n1.bias, n1.num_groups, n1.eps, silu=True), c1.weight, None, c1.stride,
                 c1.padding)
    if not ops.group_norm_supported(h, n2.weight, n2.bias, n2.num_groups):
        return None
    reg = getattr(block, "_ga_temb_shifts", None)
    if reg is not None and temb.dtype == block.time_emb_proj.weight.dtype:
        shift = reg.shift(block, temb)
    else:
        shift = F.linear(F.silu(temb), block.time_emb_proj.weight, _shift_bias(block))
    if shift.shape[0] != h.shape[0]:
        shift = shift.expand(h.shape[0], -1)
    h = F.conv2d(ops.group_norm(h, n2.weight, n2.bias, n2.num_groups, n2.eps, silu=True, shift=shift), c2.weight, None,
                 c2.stride, c2.padding)
    xs = x if block.conv_shortcut is None else block.conv_shortcut(x)
    return ops.add_bias_residual(h, c2.bias, xs)


def _patch_conv(m):
    stock = m.forward

    def forward(x, _m=m, _stock=stock):
        if (x.is_cuda and x.dim() == 4 and x.dtype in (torch.float16, torch.bfloat16) and _plain_conv(_m)
                and _m.weight.dtype == x.dtype):
            if tuple(_m.kernel_size) == (1, 1) and tuple(_m.stride) == (1, 1) and tuple(_m.padding) == (0, 0):
                # a 1x1 convolution on a channels-last tensor IS a linear layer over (pixels, channels): cuBLAS adds
                # the bias in the GEMM epilogue; the result is again a channels-last (n, C, h, w) view
                y = F.linear(x.permute(0, 2, 3, 1), _m.weight.view(_m.out_channels, _m.in_channels), _m.bias)
                return y.permute(0, 3, 1, 2)
            if _m.out_channels % 8 == 0:
                return ops.add_bias_residual(F.conv2d(x, _m.weight, None, _m.stride, _m.padding), _m.bias)
        return _stock(x)
    m.forward = forward


def register_fused_norms(unet) -> int:
    """Routes the elementwise work of the UNet around the attention layers through the fused channels-last kernels of
    `csrc/unet_ops.cu` whenever the activation is a 16-bit CUDA tensor and the parameters are frozen; anything else
    (fp32, CPU, trainable layers) keeps PyTorch's own ops:
      * every `nn.GroupNorm` (ResNet blocks, transformer wrappers, `conv_norm_out`) -> `ops.group_norm`;
      * blocks that expose the `fused_norm_act` / `fused_forward` hooks (the substrate's ResnetBlock2D) get SiLU, the
        conv1 bias, the time-embedding add, the conv2 bias and the residual add folded in (`_fused_resnet_forward`),
        and all their time-embedding projections become one GEMM per UNet pass (`_TembShifts`);
      * every biased `nn.Conv2d`: 1x1 -> `F.linear` on the channels-last view (bias in the GEMM epilogue), others ->
        bias-free cuDNN convolution + one vectorised bias pass instead of PyTorch's broadcast `add_`;
      * every `GEGLU` feed-forward gate -> `ops.geglu` (one vectorised launch per direction);
      * every `nn.LayerNorm` over the channel dimension -> `ops.layer_norm` (warp-per-row forward).
    The guided loop runs the UNet forward and backward ~250 times per image and these ops were the largest non-GEMM
    items of its launch list (csrc/unet_ops.cu, profiles/r02c_profile_ops_*.txt).  Idempotent; `GA_FUSED_NORM=0` leaves
    the UNet untouched, `GA_FUSED_DISABLE=conv,resnet,geglu,layernorm,temb` switches single items off (A/B measurements).  Returns the number of norm layers now routed."""
    if os.environ.get("GA_FUSED_NORM", "1") == "0" or not hasattr(unet, "modules"):
        return 0
    off = set(filter(None, os.environ.get("GA_FUSED_DISABLE", "").split(",")))   # A/B: any of conv, resnet, geglu, layernorm, temb
    n = 0
    for m in unet.modules():
        if isinstance(m, torch.nn.GroupNorm):
            if not getattr(m, "_ga_fused", False):
                stock = m.forward

                def forward(x, _m=m, _stock=stock):
                    if ops.group_norm_supported(x, _m.weight, _m.bias, _m.num_groups):
                        return ops.group_norm(x, _m.weight, _m.bias, _m.num_groups, _m.eps, silu=False)
                    return _stock(x)
                m.forward = forward
                m._ga_fused = True
            n += 1
        elif isinstance(m, torch.nn.Conv2d):
            if not getattr(m, "_ga_fused", False) and "conv" not in off:
                _patch_conv(m)
                m._ga_fused = True
        elif isinstance(m, torch.nn.LayerNorm):
            if not getattr(m, "_ga_fused", False) and "layernorm" not in off and len(m.normalized_shape) == 1:
                stock = m.forward

                def forward(x, _m=m, _stock=stock):
                    if ops.layer_norm_supported(x, _m.weight, _m.bias):
                        return ops.layer_norm(x, _m.weight, _m.bias, _m.eps)
                    return _stock(x)
                m.forward = forward
                m._ga_fused = True
        elif type(m).__name__ == "GEGLU" and isinstance(getattr(m, "proj", None), torch.nn.Linear):
            if not getattr(m, "_ga_fused", False) and "geglu" not in off:
                stock = m.forward

                def forward(x, *a, _m=m, _stock=stock, **kw):
                    if not a and not kw and not _m.proj.weight.requires_grad:
                        proj = _m.proj(x)
                        if ops.geglu_supported(proj):
                            return ops.geglu(proj)
                        h, gate = proj.chunk(2, dim=-1)
                        return h * F.gelu(gate)
                    return _stock(x, *a, **kw)
                m.forward = forward
                m._ga_fused = True
        else:
            if hasattr(m, "fused_norm_act"):
                m.fused_norm_act = _fused_silu_norm
            if hasattr(m, "fused_forward") and "resnet" not in off:
                m.fused_forward = _fused_resnet_forward
    if "temb" not in off:
        blocks = [m for m in unet.modules() if getattr(m, "fused_forward", None) is _fused_resnet_forward
                  and all(hasattr(m, a) for a in ("time_emb_proj", "conv1"))
                  and all(_frozen(q) for q in (m.time_emb_proj.weight, m.time_emb_proj.bias, m.conv1.bias))]
        if len(blocks) > 1 and len({b.time_emb_proj.in_features for b in blocks}) == 1:
            # ONE registry per UNet for its whole life: `register_attention_control` runs again for every prompt / seed
            # (run.py:44-67) while the pipeline's captured CUDA graphs, which hold the addresses of the registry's
            # buffers, are reused -- a fresh registry per call would free the buffers under them
            reg = getattr(unet, "_ga_temb_registry", None)
            if reg is None or [id(b) for b in reg.blocks] != [id(b) for b in blocks]:
                reg = _TembShifts(blocks)
                unet._ga_temb_registry = reg
            for b in blocks:
                b._ga_temb_shifts = reg
    return n


class AttentionControl(abc.ABC):
    """Per-forward layer counter; `between_steps` fires after `num_att_layers` calls (utils/ptp_utils.py:178-210)."""

    def step_callback(self, x_t):
        return x_t

    def between_steps(self):
        return

    @property
    def num_uncond_att_layers(self):
        return 0

    @abc.abstractmethod
    def forward(self, attn, is_cross: bool, place_in_unet: str):
        raise NotImplementedError

    def __call__(self, attn, is_cross: bool, place_in_unet: str):
        if self.cur_att_layer >= self.num_uncond_att_layers:
            self.forward(attn, is_cross, place_in_unet)
        self.cur_att_layer += 1
        if self.cur_att_layer == self.num_att_layers + self.num_uncond_att_layers:
            self.cur_att_layer = 0
            self.cur_step += 1
            self.between_steps()

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0

    def __init__(self):
        self.cur_step = 0
        self.num_att_layers = -1
        self.cur_att_layer = 0


class EmptyControl(AttentionControl):
    def forward(self, attn, is_cross: bool, place_in_unet: str):
        return attn


class AttentionStore(AttentionControl):
    """Keeps, per UNet forward, one `HeadSummedMaps` per cross layer with N <= 32^2 under the reference's six keys
    (utils/ptp_utils.py:219-270)."""

    @staticmethod
    def get_empty_store():
        return {"down_cross": [], "mid_cross": [], "up_cross": [],
                "down_self": [], "mid_self": [], "up_self": []}

    def wants_maps(self, n_query: int) -> bool:
        save_all = bool(getattr(state.config, "save_individual_CA_maps", False)) if state.config is not None else False
        return save_all or n_query <= 32 ** 2

    def forward(self, attn, is_cross: bool, place_in_unet: str):
        key = f"{place_in_unet}_{'cross' if is_cross else 'self'}"
        if not is_cross and not self.save_self_attention:
            return attn
        if self.wants_maps(attn.shape[1]) and not isinstance(attn, _ShapeOnly):
            self.step_store[key].append(attn)
        return attn

    def between_steps(self):
        self.attention_store = self.step_store
        if self.save_global_store:
            with torch.no_grad():
                if len(self.global_store) == 0:
                    self.global_store = {k: [self._as_sum(i) for i in v] for k, v in self.step_store.items()}
                else:
                    for key in self.global_store:
                        for i in range(len(self.global_store[key])):
                            self.global_store[key][i] += self._as_sum(self.step_store[key][i])
        self.step_store = self.get_empty_store()

    @staticmethod
    def _as_sum(item):
        return item.acc.detach().clone() if isinstance(item, HeadSummedMaps) else item.detach().clone()

    def get_average_attention(self):
        return self.attention_store

    def get_average_global_attention(self):
        return {key: [item / self.cur_step for item in self.global_store[key]] for key in self.attention_store}

    def reset(self):
        super(AttentionStore, self).reset()
        self.step_store = self.get_empty_store()
        self.attention_store = {}
        self.global_store = {}

    def __init__(self, save_global_store=False, save_self_attention=False):
        super(AttentionStore, self).__init__()
        self.save_global_store = save_global_store
        self.save_self_attention = save_self_attention
        self.step_store = self.get_empty_store()
        self.attention_store = {}
        self.global_store = {}
        self.curr_step_index = 0


def select_maps(attention_store: AttentionStore, res: int, from_where: List[str], is_cross: bool):
    """The stored items with N == res^2 from the requested places, in the reference's order
    (utils/ptp_utils.py:282-286)."""
    picked = []
    maps = attention_store.get_average_attention()
    for location in from_where:
        for item in maps[f"{location}_{'cross' if is_cross else 'self'}"]:
            if item.shape[1] == res ** 2:
                picked.append(item)
    return picked


def aggregate_attention(attention_store: AttentionStore, res: int, from_where: List[str], is_cross: bool,
                        select: int) -> torch.Tensor:
    """Mean over layers and heads at one resolution -> (res, res, T), differentiable (reference
    utils/ptp_utils.py:273-289).  The guided pipeline does not call this on the hot path: the tail kernel consumes the
    per-layer accumulators directly (`ops.guidance_tail`); this is the general-purpose view for other callers."""
    if select != 0:
        raise IndexError("index %d is out of bounds for dimension 0 with size 1" % select)
    picked = select_maps(attention_store, res, from_where, is_cross)
    if len(picked) == 0:
        raise RuntimeError("torch.cat(): expected a non-empty list of Tensors")
    total, n = None, 0
    for item in picked:
        if isinstance(item, HeadSummedMaps):
            part, cnt = item.acc.sum(0), item.n_maps
        else:
            part, cnt = item.float().sum(0), item.shape[0]
        total = part if total is None else total + part
        n += cnt
    out = total / n
    return out.reshape(res, res, out.shape[-1])
