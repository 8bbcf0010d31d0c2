"""SD-shaped conditional UNet substrate (random-init), plain PyTorch.

diffusers is not installed in this image (nor on the GPU box) and no checkpoints exist offline, so the UNet
that the guidance path hooks into is *defined here*.  It is a shape-faithful restatement of the
`UNet2DConditionModel` that the reference drives (reference `pipeline_guided_attention.py:583-743` shows every
member it touches; `utils/ptp_utils.py:149-175` shows the processor registry it expects).  The same definition
is shared verbatim by the CPU oracle and by the CUDA product path so that parity is well-posed.

Assumptions about diffusers 0.12.1 encoded here ("[memory]" rows) -- each is a *definition*
in this repo, not a verified fact about diffusers:
  * attention heads per level = `attention_head_dim` (8 everywhere for SD-1.x; 5/10/20/20 for SD-2.x),
    head width d = channels / heads, scale = d ** -0.5
  * to_q / to_k / to_v have no bias, to_out[0] has a bias, to_out[1] is Dropout(0)
  * head_to_batch_dim: (B, S, H*d) -> (B*H, S, d), batch-major (row index b*H + h)
  * SD-1.x uses 1x1-conv proj_in/proj_out in the transformer wrapper, SD-2.x uses Linear
  * processor names look like `down_blocks.1.attentions.0.transformer_blocks.0.attn2.processor`

UNet convolutions / linears stay in PyTorch (cuDNN / cuBLAS): they are outside the hot path (north_star).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch
from torch import nn
from torch.nn import functional as F


# ----------------------------------------------------------------------------------------------- config
@dataclass
class UNetConfig:
    sample_size: int = 64
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    attention_head_dim: Tuple[int, ...] = (8, 8, 8, 8)  # = number of heads per level (diffusers 0.12 quirk)
    norm_num_groups: int = 32
    use_linear_projection: bool = False
    center_input_sample: bool = False
    class_embed_type: Optional[str] = None
    # which down levels carry attention (SD: first three)
    down_has_attention: Tuple[bool, ...] = (True, True, True, False)

    @staticmethod
    def sd14() -> "UNetConfig":
        return UNetConfig()

    @staticmethod
    def sd21_base(sample_size: int = 96) -> "UNetConfig":
        return UNetConfig(sample_size=sample_size, cross_attention_dim=1024, attention_head_dim=(5, 10, 20, 20),
                          use_linear_projection=True)

    @staticmethod
    def tiny(sample_size: int = 64, cross_attention_dim: int = 64) -> "UNetConfig":
        """Same topology (4 levels, 16 transformer blocks, 32 processors), narrow channels; used by CPU parity tests."""
        return UNetConfig(sample_size=sample_size, block_out_channels=(32, 64, 128, 128),
                          cross_attention_dim=cross_attention_dim, attention_head_dim=(2, 2, 4, 4),
                          norm_num_groups=8)


@dataclass
class UNet2DConditionOutput:
    sample: torch.Tensor


# ------------------------------------------------------------------------------------------- attention
class DefaultAttnProcessor:
    """Plain exact attention (what diffusers' stock `CrossAttnProcessor` computes)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None):
        q = attn.to_q(hidden_states)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k = attn.to_k(ctx)
        v = attn.to_v(ctx)
        q, k, v = attn.head_to_batch_dim(q), attn.head_to_batch_dim(k), attn.head_to_batch_dim(v)
        probs = attn.get_attention_scores(q, k, attention_mask)
        out = attn.batch_to_head_dim(torch.bmm(probs, v))
        return attn.to_out[1](attn.to_out[0](out))


class CrossAttention(nn.Module):
    """Duck-compatible with the members the reference processor reads (`utils/ptp_utils.py:66-146`)."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        ctx_dim = query_dim if cross_attention_dim is None else cross_attention_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.upcast_attention = False
        self.upcast_softmax = False
        self.is_cross = cross_attention_dim is not None
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(ctx_dim, inner, bias=False)
        self.to_v = nn.Linear(ctx_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.processor = DefaultAttnProcessor()

    def set_processor(self, processor):
        self.processor = processor

    def prepare_attention_mask(self, attention_mask, target_length):
        if attention_mask is None:
            return None
        if attention_mask.shape[-1] != target_length:
            attention_mask = F.pad(attention_mask, (0, target_length), value=0.0)
            attention_mask = attention_mask.repeat_interleave(self.heads, dim=0)
        return attention_mask

    def head_to_batch_dim(self, t):
        b, s, c = t.shape
        h = self.heads
        return t.reshape(b, s, h, c // h).permute(0, 2, 1, 3).reshape(b * h, s, c // h)

    def batch_to_head_dim(self, t):
        bh, s, d = t.shape
        h = self.heads
        return t.reshape(bh // h, h, s, d).permute(0, 2, 1, 3).reshape(bh // h, s, d * h)

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        if self.upcast_attention:
            query, key = query.float(), key.float()
        scores = torch.baddbmm(
            torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype, device=query.device),
            query, key.transpose(-1, -2), beta=0, alpha=self.scale)
        if attention_mask is not None:
            scores = scores + attention_mask
        if self.upcast_softmax:
            scores = scores.float()
        return scores.softmax(dim=-1).to(dtype)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = CrossAttention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = CrossAttention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, encoder_hidden_states=None, attention_mask=None, cross_attention_kwargs=None):
        kw = cross_attention_kwargs or {}
        x = self.attn1(self.norm1(x), attention_mask=attention_mask, **kw) + x
        x = self.attn2(self.norm2(x), encoder_hidden_states=encoder_hidden_states, **kw) + x
        return self.ff(self.norm3(x)) + x


class Transformer2DModel(nn.Module):
    def __init__(self, channels, heads, cross_attention_dim, groups, use_linear_projection):
        super().__init__()
        self.use_linear_projection = use_linear_projection
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6, affine=True)
        if use_linear_projection:
            self.proj_in = nn.Linear(channels, channels)
            self.proj_out = nn.Linear(channels, channels)
        else:
            self.proj_in = nn.Conv2d(channels, channels, 1)
            self.proj_out = nn.Conv2d(channels, channels, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(channels, heads, channels // heads, cross_attention_dim)])

    def forward(self, x, encoder_hidden_states=None, attention_mask=None, cross_attention_kwargs=None):
        b, c, h, w = x.shape
        res = x
        x = self.norm(x)
        if self.use_linear_projection:
            x = self.proj_in(x.permute(0, 2, 3, 1).reshape(b, h * w, c))
        else:
            x = self.proj_in(x).permute(0, 2, 3, 1).reshape(b, h * w, c)
        for blk in self.transformer_blocks:
            x = blk(x, encoder_hidden_states=encoder_hidden_states, attention_mask=attention_mask,
                    cross_attention_kwargs=cross_attention_kwargs)
        if self.use_linear_projection:
            x = self.proj_out(x).reshape(b, h, w, c).permute(0, 3, 1, 2)
        else:
            x = self.proj_out(x.reshape(b, h, w, c).permute(0, 3, 1, 2))
        return x + res


# ---------------------------------------------------------------------------------------------- resnet
class ResnetBlock2D(nn.Module):
    def __init__(self, in_ch, out_ch, temb_ch, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_ch, eps=1e-5)
        self.conv1 = nn.Conv2d(in_ch, out_ch, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_ch, out_ch)
        self.norm2 = nn.GroupNorm(groups, out_ch, eps=1e-5)
        self.conv2 = nn.Conv2d(out_ch, out_ch, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_ch, out_ch, 1) if in_ch != out_ch else None
        # hook for a fused `silu(norm(x))` (callable(norm_module, x) -> tensor); None = the two stock ops.  Installed by
        # the product path (`ptp_utils.register_fused_norms`), never by the oracle.
        self.fused_norm_act = None
        # hook for the whole block (callable(block, x, temb) -> tensor, or None to decline): same mathematics with the
        # conv biases / time-embedding add / residual add folded into neighbouring kernels
        self.fused_forward = None

    def _norm_act(self, norm, x):
        if self.fused_norm_act is not None:
            return self.fused_norm_act(norm, x)
        return F.silu(norm(x))

    def forward(self, x, temb):
        if self.fused_forward is not None:
            out = self.fused_forward(self, x, temb)
            if out is not None:
                return out
        h = self.conv1(self._norm_act(self.norm1, x))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(self._norm_act(self.norm2, h))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Downsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x, output_size=None):
        if output_size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=output_size, mode="nearest")
        return self.conv(x)


# ---------------------------------------------------------------------------------------------- blocks
class DownBlock(nn.Module):
    def __init__(self, in_ch, out_ch, temb_ch, n_layers, groups, heads, cross_dim, has_attn, add_down, linear_proj):
        super().__init__()
        self.has_cross_attention = has_attn
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(in_ch if i == 0 else out_ch, out_ch, temb_ch, groups) for i in range(n_layers)])
        self.attentions = nn.ModuleList(
            [Transformer2DModel(out_ch, heads, cross_dim, groups, linear_proj) for _ in range(n_layers)]
        ) if has_attn else None
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch)]) if add_down else None

    def forward(self, hidden_states, temb, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs=None):
        outs = ()
        for i, res in enumerate(self.resnets):
            hidden_states = res(hidden_states, temb)
            if self.attentions is not None:
                hidden_states = self.attentions[i](hidden_states, encoder_hidden_states=encoder_hidden_states,
                                                   attention_mask=attention_mask,
                                                   cross_attention_kwargs=cross_attention_kwargs)
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class MidBlock(nn.Module):
    def __init__(self, ch, temb_ch, groups, heads, cross_dim, linear_proj):
        super().__init__()
        self.has_cross_attention = True
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch, groups), ResnetBlock2D(ch, ch, temb_ch, groups)])
        self.attentions = nn.ModuleList([Transformer2DModel(ch, heads, cross_dim, groups, linear_proj)])

    def forward(self, hidden_states, temb, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        hidden_states = self.attentions[0](hidden_states, encoder_hidden_states=encoder_hidden_states,
                                           attention_mask=attention_mask,
                                           cross_attention_kwargs=cross_attention_kwargs)
        return self.resnets[1](hidden_states, temb)


class UpBlock(nn.Module):
    def __init__(self, in_ch, prev_ch, out_ch, temb_ch, n_layers, groups, heads, cross_dim, has_attn, add_up,
                 linear_proj):
        super().__init__()
        self.has_cross_attention = has_attn
        resnets = []
        for i in range(n_layers):
            skip_ch = in_ch if i == n_layers - 1 else out_ch
            res_in = prev_ch if i == 0 else out_ch
            resnets.append(ResnetBlock2D(res_in + skip_ch, out_ch, temb_ch, groups))
        self.resnets = nn.ModuleList(resnets)
        self.attentions = nn.ModuleList(
            [Transformer2DModel(out_ch, heads, cross_dim, groups, linear_proj) for _ in range(n_layers)]
        ) if has_attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_up else None

    def forward(self, hidden_states, temb, res_hidden_states_tuple, encoder_hidden_states=None,
                cross_attention_kwargs=None, upsample_size=None, attention_mask=None):
        for i, res in enumerate(self.resnets):
            skip = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = res(torch.cat([hidden_states, skip], dim=1), temb)
            if self.attentions is not None:
                hidden_states = self.attentions[i](hidden_states, encoder_hidden_states=encoder_hidden_states,
                                                   attention_mask=attention_mask,
                                                   cross_attention_kwargs=cross_attention_kwargs)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


# ------------------------------------------------------------------------------------------ time embed
class Timesteps(nn.Module):
    """Sinusoidal embedding, flip_sin_to_cos=True, freq_shift=0 (SD setting)."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, t):
        half = self.dim // 2
        exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
        emb = t[:, None].float() * torch.exp(exponent)[None, :]
        return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim, dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


# ------------------------------------------------------------------------------------------------ unet
class UNet2DConditionModel(nn.Module):
    def __init__(self, config: UNetConfig):
        super().__init__()
        self.config = config
        ch = config.block_out_channels
        temb_ch = ch[0] * 4
        g = config.norm_num_groups
        heads = config.attention_head_dim
        lp = config.use_linear_projection
        self.in_channels = config.in_channels
        self.class_embedding = None

        self.conv_in = nn.Conv2d(config.in_channels, ch[0], 3, padding=1)
        self.time_proj = Timesteps(ch[0])
        self.time_embedding = TimestepEmbedding(ch[0], temb_ch)

        self.down_blocks = nn.ModuleList()
        out = ch[0]
        for i, c in enumerate(ch):
            inp, out = out, c
            last = i == len(ch) - 1
            self.down_blocks.append(DownBlock(inp, out, temb_ch, config.layers_per_block, g, heads[i],
                                              config.cross_attention_dim, config.down_has_attention[i],
                                              not last, lp))
        self.mid_block = MidBlock(ch[-1], temb_ch, g, heads[-1], config.cross_attention_dim, lp)

        self.up_blocks = nn.ModuleList()
        rch = list(reversed(ch))
        rheads = list(reversed(heads))
        rattn = list(reversed(config.down_has_attention))
        out = rch[0]
        self.num_upsamplers = 0
        for i, c in enumerate(rch):
            prev, out = out, c
            inp = rch[min(i + 1, len(ch) - 1)]
            last = i == len(ch) - 1
            if not last:
                self.num_upsamplers += 1
            self.up_blocks.append(UpBlock(inp, prev, out, temb_ch, config.layers_per_block + 1, g, rheads[i],
                                          config.cross_attention_dim, rattn[i], not last, lp))

        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[0], config.out_channels, 3, padding=1)

    # -- processor registry (the plug-in point used by `register_attention_control`, ptp_utils.py:153,174)
    def _attn_modules(self):
        for name, mod in self.named_modules():
            if isinstance(mod, CrossAttention):
                yield name, mod

    @property
    def attn_processors(self) -> Dict[str, object]:
        return {f"{name}.processor": mod.processor for name, mod in self._attn_modules()}

    def set_attn_processor(self, processor):
        for name, mod in self._attn_modules():
            if isinstance(processor, dict):
                key = f"{name}.processor"
                if key in processor:
                    mod.set_processor(processor[key])
            else:
                mod.set_processor(processor)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, attention_mask=None,
                cross_attention_kwargs=None, return_dict=True):
        """Stock forward (equal to the reference's patched copy when `optimizeDeepLatent` is off,
        reference `pipeline_guided_attention.py:583-743`)."""
        up_factor = 2 ** self.num_upsamplers
        forward_upsample_size = any(s % up_factor != 0 for s in sample.shape[-2:])
        upsample_size = None
        if attention_mask is not None:
            attention_mask = ((1 - attention_mask.to(sample.dtype)) * -10000.0).unsqueeze(1)
        if self.config.center_input_sample:
            sample = 2 * sample - 1.0
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            dtype = torch.float64 if isinstance(timestep, float) else torch.int64
            timesteps = torch.tensor([timesteps], dtype=dtype, device=sample.device)
        elif timesteps.dim() == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        emb = self.time_embedding(self.time_proj(timesteps).to(dtype=self.dtype))

        if getattr(self, "channels_last", False):
            sample = sample.contiguous(memory_format=torch.channels_last)
        sample = self.conv_in(sample)
        skips = (sample,)
        for blk in self.down_blocks:
            if blk.has_cross_attention:
                sample, outs = blk(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states,
                                   attention_mask=attention_mask, cross_attention_kwargs=cross_attention_kwargs)
            else:
                sample, outs = blk(hidden_states=sample, temb=emb)
            skips += outs
        sample = self.mid_block(sample, emb, encoder_hidden_states=encoder_hidden_states,
                                attention_mask=attention_mask, cross_attention_kwargs=cross_attention_kwargs)
        for i, blk in enumerate(self.up_blocks):
            final = i == len(self.up_blocks) - 1
            n = len(blk.resnets)
            res, skips = skips[-n:], skips[:-n]
            if not final and forward_upsample_size:
                upsample_size = skips[-1].shape[2:]
            if blk.has_cross_attention:
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res,
                             encoder_hidden_states=encoder_hidden_states,
                             cross_attention_kwargs=cross_attention_kwargs, upsample_size=upsample_size,
                             attention_mask=attention_mask)
            else:
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res,
                             upsample_size=upsample_size)
        sample = self.conv_out(self.conv_act(self.conv_norm_out(sample)))
        if not return_dict:
            return (sample,)
        return UNet2DConditionOutput(sample=sample)


def build_unet(config: UNetConfig, seed: int = 0, dtype=torch.float32, device="cpu",
               channels_last=None) -> UNet2DConditionModel:
    """Random-init UNet, deterministic: built on CPU in fp32 under `torch.manual_seed(seed)` (default nn init),
    then cast / moved.  Every rank of a seed sweep builds the identical model (no broadcast needed).
    On CUDA the model is kept in channels-last memory format by default: cuDNN's fp16 tensor-core convolutions are NHWC
    kernels, and with NCHW activations a fifth of the UNet's GPU time goes into nchw<->nhwc transposes (profiles/)."""
    gen_state = torch.random.get_rng_state()
    try:
        torch.manual_seed(seed)
        model = UNet2DConditionModel(config)
    finally:
        torch.random.set_rng_state(gen_state)
    model = model.to(dtype=dtype, device=device)
    if channels_last is None:
        channels_last = torch.device(device).type == "cuda"
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    model.channels_last = bool(channels_last)
    model.eval()
    for p in model.parameters():
        p.requires_grad_(False)
    return model
