"""Guided-attention pipeline -- drop-in mirror of the reference's `pipeline_guided_attention.py` for the per-step guidance
path (same class name, method names, argument meaning and return shapes), with the loss evaluation and its gradient
running as fused sm_100a launches instead of ~30k tiny ATen ops:

  reference                                              here
  -----------------------------------------------------  ---------------------------------------------------------------
  aggregate_attention (ptp_utils.py:273-289)              }
  _compute_max_attention_per_index (:201-296)             }  ONE launch: ops.guidance_tail  (csrc/guidance_tail.cu)
  helpers.calculate_bounding_box_losses (helpers:215-277) }
  _compute_loss (:398-451)                                }
  autograd of all of the above (~11k backward ops)           ONE launch: its backward
  18 `.item()` syncs + PNG writes per evaluation             one small D2H read, only on steps that test a threshold

The UNet, scheduler and prompt encoding are out of scope: they come from
`guided_attention_b200.substrate` (or any object with the same duck interface, e.g. a diffusers UNet).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import _cabi as abi
from . import helpers, ops
from . import shared_state as state
from .ptp_utils import (AttentionStore, HeadSummedMaps, PaintWithWords, TextKVCache, aggregate_attention,
                        select_maps)
from .substrate import DDIMScheduler, StableDiffusionPipelineBase, StableDiffusionPipelineOutput

AT = helpers.AnnotationType
_KIND = {AT.COOR: abi.GA_TOKEN_COOR, AT.BOX: abi.GA_TOKEN_BOX, AT.KEYWORD: abi.GA_TOKEN_KEYWORD}



def _token_table_key(token_dict):
    """Hashable snapshot of `config.token_dict` (run.py:81-91): everything the tail spec and the captured graphs bake in."""
    def payload(v):
        if hasattr(v, 'as_tuple'):
            return tuple(v.as_tuple()) + (getattr(v, 'size', None),)
        return tuple(v) if isinstance(v, (list, tuple)) else v
    return tuple((i, v['loss_type'], v['subprompt'], payload(v['loss'])) for i, v in token_dict.items())


class _StepGraphs:
    """CUDA-graph execution of the three device programs the guided loop keeps re-issuing (DESIGN.md "CUDA graphs"):

      eval    text-conditioned UNet forward (hooks -> K1 accumulators) + guidance tail          -> per-token stats, loss
      update  the same forward with the autograd graph + tail + backward (tail bwd, K2, UNet bwd)
              + latents <- latents - step * grad                                               -> stats, new latents
      cfg     CFG UNet forward (batch 2) + guidance combine + DDIM step                        -> next latents

    The reference re-launches ~10^3 kernels from Python for each of these; a replay is one graph launch.  Inputs are
    copied into static buffers (latents, timestep, step size, DDIM coefficients, text embeddings); control flow
    (thresholds, refinement loop) stays on the host and reads only the small `stats` tensor."""

    def __init__(self, pipe, store, loss_kw, prompt_embeds, guidance_scale, latents_like):
        self.pipe, self.store, self.loss_kw, self.gs = pipe, store, loss_kw, float(guidance_scale)
        dev, dt = latents_like.device, latents_like.dtype
        self.lat = torch.zeros_like(latents_like)
        self.lat_out = torch.zeros_like(latents_like)
        self.t = torch.zeros(1, dtype=torch.int64, device=dev)
        self.step = torch.zeros((), dtype=torch.float32, device=dev)
        self.coef = torch.zeros(4, dtype=torch.float32, device=dev)
        self.embeds = prompt_embeds.detach().clone()
        self.graphs, self.outputs, self.launches, self.replays = {}, {}, {}, {}
        self.pool = None
        # text K/V projections are computed once per embedding buffer instead of once per UNet pass (ptp_utils.TextKVCache)
        self.text_kv = TextKVCache()
        # device-side step driver (csrc/step_driver.cu): re-noise coefficients, the pre-drawn re-noise
        # tensors of one image with their draw counter, the control block, and the driver handle
        hp = state.curHyperParams or {}
        n_steps = int(pipe.scheduler.num_inference_steps or 0)
        self.max_draws = max(1, n_steps * max(int(hp.get("recurse_steps", 1)) - 1, 0))
        self.bt = torch.zeros(2, dtype=torch.float32, device=dev)
        self.noise = torch.zeros((self.max_draws,) + tuple(latents_like.shape), dtype=torch.float32, device=dev)
        self.n_draws = torch.zeros(1, dtype=torch.int64, device=dev)
        self.ctl = torch.zeros(abi.GA_STEP_CTL_BYTES // 4, dtype=torch.int32, device=dev)
        self._driver = None
        # `use_optimizer` (reference :495-497, :545-547): the refinement loop steps the latents with SGD + momentum 0.8;
        # `mom` is the momentum buffer, `first` = 1 on the first iteration of a refinement (fresh optimizer state)
        self.use_optimizer = bool(hp.get("use_optimizer", False))
        self.mom = torch.zeros_like(latents_like)
        self.first = torch.ones((), dtype=torch.float32, device=dev)

    def set_embeds(self, prompt_embeds):
        self.embeds.copy_(prompt_embeds)
        self.text_kv.refresh()

    # -- the three programs ---------------------------------------------------------------------------------------
    def _prog_eval(self):
        with torch.no_grad():
            self.pipe.unet(self.lat, self.t, encoder_hidden_states=self.embeds[1:2])
            ld = self.pipe._aggregate_and_get_max_attention_per_token(**self.loss_kw)
            loss, losses, unscaled = self.pipe._compute_loss(ld)
        return {"loss": loss, "losses": losses, "unscaled": unscaled, "stats": ld["_stats"]}

    def _prog_update(self):
        with torch.enable_grad():
            lat = self.lat.detach().clone().requires_grad_(True)
            self.pipe.unet(lat, self.t, encoder_hidden_states=self.embeds[1:2])
            ld = self.pipe._aggregate_and_get_max_attention_per_token(**self.loss_kw)
            loss, losses, unscaled = self.pipe._compute_loss(ld)
            (g,) = torch.autograd.grad(loss, [lat])
        with torch.no_grad():
            self.lat_out.copy_((lat.detach().float() - self.step * g.float()).to(lat.dtype))
        return {"loss": loss.detach(), "losses": [(i, v.detach()) for i, v in losses],
                "unscaled": [(i, v.detach()) for i, v in unscaled], "stats": ld["_stats"].detach()}

    def _prog_update_mom(self):
        """One refinement iteration under `use_optimizer`: torch.optim.SGD(lr = step_size / 2.5, momentum = 0.8) on the
        latents (reference :495-497: a fresh optimizer per refinement; :545-547: loss.backward(); optim.step()), i.e.
        buf <- 0.8 buf + grad (buf = grad on the first iteration), latents <- latents - lr buf, in the latents' dtype."""
        with torch.enable_grad():
            lat = self.lat.detach().clone().requires_grad_(True)
            self.pipe.unet(lat, self.t, encoder_hidden_states=self.embeds[1:2])
            ld = self.pipe._aggregate_and_get_max_attention_per_token(**self.loss_kw)
            loss, losses, unscaled = self.pipe._compute_loss(ld)
            (g,) = torch.autograd.grad(loss, [lat])
        with torch.no_grad():
            keep = (0.8 * (1.0 - self.first)).to(self.mom.dtype)
            self.mom.mul_(keep).add_(g)
            lr = self.step / 2.5
            self.lat_out.copy_((lat.detach().float() - lr * self.mom.float()).to(lat.dtype))
        return {"loss": loss.detach(), "losses": [(i, v.detach()) for i, v in losses],
                "unscaled": [(i, v.detach()) for i, v in unscaled], "stats": ld["_stats"].detach()}

    def _prog_cfg(self):
        with torch.no_grad():
            x2 = torch.cat([self.lat] * 2)
            noise = self.pipe.unet(x2, self.t, encoder_hidden_states=self.embeds).sample
            n_u, n_t = noise.chunk(2)
            noise = n_u + self.gs * (n_t - n_u)
            out = self.pipe.scheduler.step(noise, None, self.lat, coeffs=(self.coef[0], self.coef[1], self.coef[2],
                                                                          self.coef[3]))
            self.lat_out.copy_(out.prev_sample)
        return {}

    def _prog_advance(self):
        with torch.no_grad():
            self.lat.copy_(self.lat_out)
        return {}

    def _prog_renoise(self):
        """Back to the previous noise level before a recursion round (reference :1046-1050), with the next of the
        image's pre-drawn noise tensors: same fp32 arithmetic as `GuidedAttention._renoise`."""
        with torch.no_grad():
            # (clamped: the warm-up passes before capture advance the counter past a one-entry buffer)
            noise = self.noise.index_select(0, self.n_draws.clamp(max=self.max_draws - 1))[0]
            self.lat.copy_((self.bt[0] * self.lat.float() + self.bt[1] * noise).to(self.lat.dtype))
            self.n_draws.add_(1)
        return {}

    def driver(self):
        """The step driver for these programs (`ga_step_driver_create`), built on first use: captures all five
        programs, hands their raw cudaGraph_t handles plus the static `stats` buffers to the library."""
        if self._driver is not None:
            return self._driver
        self.store.text_kv = self.text_kv
        names = ("eval", "update", "cfg", "advance", "renoise") + (("update_mom",) if self.use_optimizer else ())
        for name in names:
            self._graph(name)
        lib = abi.load()
        progs = abi.GaStepPrograms(*[int(self.graphs[n].raw_cuda_graph()) for n in names])

        def custom_of(out):
            if out["losses"] and out["losses"][-1][0] is None and torch.is_tensor(out["losses"][-1][1]):
                c = out["losses"][-1][1]
                if c.dtype != torch.float32 or not c.is_cuda:
                    raise RuntimeError("custom loss term must be a float32 CUDA tensor")
                return c
            return None
        spec = self.pipe._tail_spec_cache[1]
        refine = self.outputs["update_mom"] if self.use_optimizer else None
        self._custom = (custom_of(self.outputs["eval"]), custom_of(self.outputs["update"]),
                        custom_of(refine) if refine is not None else None)

        def ptr(t):
            return C.c_void_p(t.data_ptr() if t is not None else 0)
        handle = C.c_void_p()
        abi.check(lib.ga_step_driver_create(
            C.byref(handle), C.byref(progs), ptr(self.ctl), ptr(self.outputs["eval"]["stats"]),
            ptr(self.outputs["update"]["stats"]), ptr(self._custom[0]), ptr(self._custom[1]),
            ptr(refine["stats"] if refine is not None else None), ptr(self._custom[2]), ptr(self.first),
            spec.tokens, len(spec.token_indices), int(spec.params.n_groups),
            int(bool(getattr(state.config, "sub_prompt_avg_within", False))),
            C.c_void_p(self.t.data_ptr()), C.c_void_p(self.step.data_ptr()), C.c_void_p(self.coef.data_ptr()),
            C.c_void_p(self.bt.data_ptr())), "ga_step_driver_create")
        self._driver = handle
        return handle

    def __del__(self):
        try:
            if getattr(self, "_driver", None) is not None:
                abi.load().ga_step_driver_destroy(self._driver)
        except Exception:
            pass

    def _graph(self, name):
        if name in self.graphs:
            return self.graphs[name]
        prog = getattr(self, "_prog_" + name)
        # warm-up on a side stream (allocator, cuDNN heuristics, lazy kernel attributes), then capture
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                prog()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        snap = dict(ops.launch_counts)
        # keep_graph: the raw cudaGraph_t stays available for the step driver (csrc/step_driver.cu clones it into its
        # conditional graph); `replay()` instantiates on first use
        g = torch.cuda.CUDAGraph(keep_graph=True)
        with torch.cuda.graph(g, pool=self.pool):
            self.outputs[name] = prog()
        if self.pool is None:
            self.pool = g.pool()
        # launches recorded by the capture pass = launches every replay performs; the capture itself ran nothing
        self.launches[name] = {k: v - snap.get(k, 0) for k, v in ops.launch_counts.items() if v - snap.get(k, 0) > 0}
        for k, v in self.launches[name].items():
            ops.launch_counts[k] -= v
        self.graphs[name] = g
        return g

    def run(self, name, latents, t, step_size=None, coeffs=None):
        """Copies the inputs into the static buffers and replays `name`; returns (static outputs, new latents | None).
        The static outputs are overwritten by the next replay of the same program."""
        self.store.text_kv = self.text_kv
        pww = PaintWithWords.of(self.store)
        if pww is not None:          # the step-dependent paint-with-words factor is data of the captured programs
            pww.update(self.lat.device)
        g = self._graph(name)
        self.lat.copy_(latents)
        self.t.fill_(int(t))
        if step_size is not None:
            self.step.fill_(float(step_size))
        if coeffs is not None:
            for n, c in enumerate(coeffs):
                self.coef[n].fill_(float(c))
        g.replay()
        self.replays[name] = self.replays.get(name, 0) + 1
        self.pipe._count_pass("update" if name == "update_mom" else name)
        for k, v in self.launches[name].items():
            ops._count(k, v)
        return self.outputs[name], (self.lat_out.clone() if name != "eval" else None)


class _BatchGraphs:
    """Seed-batched versions of the three device programs (extension): S independent seeds advance
    through the same UNet pass, every sample keeps its own attention accumulators (K1 writes `acc[b]` per batch element),
    the tail kernel evaluates the S losses in one launch (`n_samples = S`) and the latent step takes a per-sample step
    size, so seeds that must not move in a given pass simply get step 0.  Nothing couples the samples (eval-mode UNet:
    GroupNorm / LayerNorm / attention are per sample), which the tests check against the one-seed path."""

    def __init__(self, pipe, store, loss_kw, prompt_embeds, guidance_scale, latents_like, eager=False):
        self.pipe, self.store, self.loss_kw, self.gs, self.eager = pipe, store, loss_kw, float(guidance_scale), eager
        S = latents_like.shape[0]
        dev = latents_like.device
        self.S = S
        self.lat = torch.zeros_like(latents_like)
        self.lat_out = torch.zeros_like(latents_like)
        self.t = torch.zeros(1, dtype=torch.int64, device=dev)
        self.step = torch.zeros(S, 1, 1, 1, dtype=torch.float32, device=dev)
        self.coef = torch.zeros(4, dtype=torch.float32, device=dev)
        self.text_kv = TextKVCache()
        self.set_embeds(prompt_embeds)
        self.graphs, self.outputs, self.launches, self.replays = {}, {}, {}, {}
        self.pool = None

    def set_embeds(self, prompt_embeds):
        S = self.S
        cond = prompt_embeds[1:2].detach().expand(S, -1, -1).contiguous()
        unc = prompt_embeds[0:1].detach().expand(S, -1, -1).contiguous()
        if hasattr(self, "cond"):
            self.cond.copy_(cond)
            self.both.copy_(torch.cat([unc, cond]))
        else:
            self.cond, self.both = cond, torch.cat([unc, cond])
        self.text_kv.refresh()

    def _loss(self):
        kw = self.loss_kw
        picked = select_maps(self.store, kw["attention_res"], ("up", "down", "mid"), True)
        if len(picked) == 0:
            raise RuntimeError("torch.cat(): expected a non-empty list of Tensors")
        accs = [m.acc for m in picked]                       # each (S, N, T): one slice per sample
        res, T = kw["attention_res"], accs[0].shape[2]
        spec = self.pipe._tail_spec(res, T, kw["smooth_attentions"], kw["sigma"], kw["kernel_size"],
                                    kw["normalize_eot"], accs[0].device)
        n_maps = sum(m.heads for m in picked)                # maps per sample
        _, _, stats, _, total = ops.guidance_tail(spec, accs, n_maps, n_samples=self.S)
        if self.S == 1:
            stats, total = stats[None], total
        custom = None
        if getattr(state.config, "custom_loss", None):
            # built-in keyword losses per sample, from the tail's raw-map statistics (scalars; differentiable via g_stats)
            terms = []
            for s_ in range(self.S):
                c = torch.zeros(1, dtype=torch.float32, device=stats.device)
                for _, (loss_obj, args) in state.config.custom_loss.items():
                    c = c + loss_obj.calc_loss_from_stats(stats[s_], spec.token_indices, res, args)
                terms.append(c)
            custom = torch.cat(terms)
            total = total + custom
        return stats, total, custom

    def _prog_eval(self):
        with torch.no_grad():
            self.pipe.unet(self.lat, self.t, encoder_hidden_states=self.cond)
            stats, total, custom = self._loss()
        return {"stats": stats, "total": total, "custom": custom}

    def _prog_update(self):
        with torch.enable_grad():
            lat = self.lat.detach().clone().requires_grad_(True)
            self.pipe.unet(lat, self.t, encoder_hidden_states=self.cond)
            stats, total, custom = self._loss()
            (g,) = torch.autograd.grad(total.sum(), [lat])   # samples are independent: row s of g is d total[s] / d lat[s]
        with torch.no_grad():
            self.lat_out.copy_((lat.detach().float() - self.step * g.float()).to(lat.dtype))
        return {"stats": stats.detach(), "total": total.detach(), "custom": None if custom is None else custom.detach()}

    def _prog_cfg(self):
        with torch.no_grad():
            x2 = torch.cat([self.lat] * 2)
            noise = self.pipe.unet(x2, self.t, encoder_hidden_states=self.both).sample
            n_u, n_t = noise.chunk(2)
            noise = n_u + self.gs * (n_t - n_u)
            out = self.pipe.scheduler.step(noise, None, self.lat, coeffs=(self.coef[0], self.coef[1], self.coef[2],
                                                                          self.coef[3]))
            self.lat_out.copy_(out.prev_sample)
        return {}

    _graph = _StepGraphs._graph

    def run(self, name, latents, t, step_sizes=None, coeffs=None):
        self.store.text_kv = self.text_kv
        self.lat.copy_(latents)
        self.t.fill_(int(t))
        if step_sizes is not None:
            self.step.copy_(torch.tensor(step_sizes, dtype=torch.float32).reshape(-1, 1, 1, 1), non_blocking=True)
        if coeffs is not None:
            for n, c in enumerate(coeffs):
                self.coef[n].fill_(float(c))
        if self.eager:
            out = getattr(self, "_prog_" + name)()
        else:
            g = self._graph(name)
            g.replay()
            for k, v in self.launches[name].items():
                ops._count(k, v)
            out = self.outputs[name]
        self.replays[name] = self.replays.get(name, 0) + 1
        self.pipe._count_pass(name)
        return out, (self.lat_out.clone() if name != "eval" else None)


class GuidedAttention(StableDiffusionPipelineBase):
    """Pipeline for text-to-image generation with cross-attention guidance (boxes, crosshairs, keyword losses)."""

    _optional_components = ["safety_checker", "feature_extractor"]
    update_calls = 0          # number of eager `_update_latent` backward passes (class-level: the method is static)
    inside_iterative_refinement = False
    optim = None

    # ------------------------------------------------------------------------------------------ prompt encoding
    def _encode_prompt(self, prompt, device, num_images_per_prompt, do_classifier_free_guidance, negative_prompt=None,
                       prompt_embeds: Optional[torch.Tensor] = None,
                       negative_prompt_embeds: Optional[torch.Tensor] = None):
        """Returns (text_inputs, cat([negative, positive])) like the reference (:64-199).  No CLIP weights exist offline:
        `prompt_embeds` (and `negative_prompt_embeds` for CFG) must be supplied unless a text encoder was attached."""
        text_inputs = None
        if prompt_embeds is None:
            if not callable(getattr(self, "text_encoder", None)):
                raise ValueError("no text encoder is available offline: pass `prompt_embeds` / `negative_prompt_embeds`")
            text_inputs = self.tokenizer(prompt, padding="max_length", max_length=self.tokenizer.model_max_length,
                                         truncation=True, return_tensors="pt")
            prompt_embeds = self.text_encoder(text_inputs.input_ids.to(device), attention_mask=None)[0]
        elif prompt is not None:
            text_inputs = self.tokenizer(prompt, padding="max_length", max_length=self.tokenizer.model_max_length,
                                         truncation=True, return_tensors="pt")
        dtype = self.unet.dtype
        prompt_embeds = prompt_embeds.to(dtype=dtype, device=device)
        bs, seq, _ = prompt_embeds.shape
        prompt_embeds = prompt_embeds.repeat(1, num_images_per_prompt, 1).view(bs * num_images_per_prompt, seq, -1)
        if do_classifier_free_guidance:
            if negative_prompt_embeds is None:
                if not callable(getattr(self, "text_encoder", None)):
                    raise ValueError("classifier-free guidance needs `negative_prompt_embeds` (no text encoder offline)")
                tokens = [negative_prompt or ""] * bs if not isinstance(negative_prompt, list) else negative_prompt
                uncond = self.tokenizer(tokens, padding="max_length", max_length=seq, truncation=True,
                                        return_tensors="pt")
                negative_prompt_embeds = self.text_encoder(uncond.input_ids.to(device), attention_mask=None)[0]
            negative_prompt_embeds = negative_prompt_embeds.to(dtype=dtype, device=device)
            negative_prompt_embeds = negative_prompt_embeds.repeat(1, num_images_per_prompt, 1).view(
                bs * num_images_per_prompt, negative_prompt_embeds.shape[1], -1)
            prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds])
        return text_inputs, prompt_embeds

    # ------------------------------------------------------------------------------------------ tail specification
    def _tail_spec(self, attention_res: int, n_ctx: int, smooth_attentions: bool, sigma: float, kernel_size: int,
                   normalize_eot: bool, device) -> ops.TailSpec:
        """Host-side, once per (prompt, hyper-parameters, resolution): token table, rasterised masks, strict weights.
        Mirrors what the reference re-derives inside every loss evaluation (pipeline :209-228, :277-279;
        helpers.py:216-246; :405-430)."""
        cfg, hp = state.config, state.curHyperParams
        last_idx = n_ctx - 1
        if normalize_eot:
            prompt = self.prompt[0] if isinstance(self.prompt, list) else self.prompt
            last_idx = len(self.tokenizer(prompt)['input_ids']) - 1
        token_dict = cfg.token_dict
        # keyed on the CONTENTS of the token table (the reference re-derives all of this on every evaluation): a dict
        # mutated in place, or a new dict at a recycled address, must not hit a stale spec
        key = (_token_table_key(token_dict), attention_res, n_ctx, last_idx, bool(smooth_attentions),
               float(sigma), int(kernel_size), bool(hp["strict"]), float(hp["shrink_factor"]),
               float(hp["inside_loss_scale"]), float(hp["outside_loss_scale"]), float(hp.get("bb_center_weight", .05)),
               bool(getattr(cfg, "sub_prompt_avg_within", False)), str(device))
        cached = getattr(self, "_tail_spec_cache", None)
        if cached is not None and cached[0] == key:
            return cached[1]

        res = attention_res
        indices = list(token_dict.keys())
        if len(indices) > abi.GA_MAX_TOKENS:
            raise ValueError(f"at most {abi.GA_MAX_TOKENS} tracked tokens are supported, got {len(indices)}")
        subprompts = [token_dict[i]['subprompt'] for i in indices]
        group_ids = {s: g for g, s in enumerate(dict.fromkeys(subprompts))}
        # `_compute_loss` only emits a loss for COOR and BOX tokens; the sub-prompt mean divides by those (:376-379)
        emitting = [token_dict[i]['loss_type'] in (AT.COOR, AT.BOX) for i in indices]
        per_group = {s: sum(1 for s2, e in zip(subprompts, emitting) if e and s2 == s) for s in group_ids}
        avg_within = bool(getattr(cfg, "sub_prompt_avg_within", False))
        center_w = float(hp.get("bb_center_weight", .05))

        boxes, box_of_token = [], {}
        for i in indices:
            info = token_dict[i]
            if info['loss_type'] == AT.BOX:
                r = info['loss']
                box_of_token[i] = len(boxes)
                boxes.append((r.x / r.size, r.y / r.size, r.width / r.size, r.height / r.size)
                             if r.size != 1 else (r.x, r.y, r.width, r.height))
        masks = ops.rasterize_boxes(boxes, res, float(hp["shrink_factor"]), device) if boxes else None
        n_inside, weights = [], None
        if boxes:
            n_inside = [int(v) for v in masks.reshape(len(boxes), -1).sum(1).tolist()]   # one sync per prompt
            if any(n == 0 for n in n_inside):
                raise ZeroDivisionError("float division by zero")   # reference: at_most = 1.0 / num_inside
            if hp["strict"]:
                scaled = [token_dict[i]['loss'].of_size(float(res)) for i in indices if i in box_of_token]
                saved = state.curHyperParams
                weights = torch.from_numpy(np.stack([helpers.strict_weights_host(r, res) for r in scaled])).to(device)
                state.curHyperParams = saved

        toks = (abi.GaToken * max(len(indices), 1))()
        for n, i in enumerate(indices):
            info = token_dict[i]
            kind = info['loss_type']
            t = toks[n]
            t.column, t.kind, t.box = i - 1, _KIND[kind], box_of_token.get(i, -1)
            t.group = group_ids[info['subprompt']]
            t.group_weight = (1.0 / per_group[info['subprompt']]) if (avg_within and emitting[n]) else 1.0
            if kind == AT.BOX:
                cx, cy = info['loss'].center()
                t.center_weight = center_w
            elif kind == AT.COOR:
                cx, cy = info['loss']
                t.center_weight = 1.0
            else:
                cx, cy, t.center_weight = 0.0, 0.0, 0.0
            # reference: `col - center[0]*16` with a float32 tensor -> the python double is rounded to float32
            t.target_x, t.target_y = float(np.float32(cx * res)), float(np.float32(cy * res))
        p = abi.GaTailParams()
        p.res, p.n_ctx, p.first, p.last = res, n_ctx, 1, last_idx
        p.n_tokens, p.n_groups = len(indices), len(group_ids)
        p.strict, p.smooth = int(bool(hp["strict"])), int(bool(smooth_attentions))
        taps = ops.gaussian_taps(kernel_size, sigma) if smooth_attentions else [0.0, 1.0, 0.0]
        for a in range(3):
            p.w1d[a] = taps[a]
        p.temperature, p.inv_count = 100.0, 1.0
        p.inside_scale = float(hp["inside_loss_scale"])
        p.outside_scale = float(hp["outside_loss_scale"] * 3)
        p.n_samples = 1
        spec = ops.TailSpec(res=res, n_ctx=n_ctx, first=1, last=last_idx, token_indices=indices,
                            kinds=[_KIND[token_dict[i]['loss_type']] for i in indices], groups=subprompts, tokens=toks,
                            params=p, masks=masks, weights=weights, n_inside=n_inside)
        self._tail_spec_cache = (key, spec)
        return spec

    # ------------------------------------------------------------------------------------------- loss evaluation
    def _compute_max_attention_per_index(self, attention_maps, smooth_attentions: bool = False, sigma: float = 0.5,
                                         kernel_size: int = 3, normalize_eot: bool = False):
        """Reference signature (:201-206) takes the aggregated (res, res, T) map.  Accepts either that tensor (it is then
        treated as a single accumulator holding one map) or a list of `HeadSummedMaps`; returns the reference's
        `losses_dict` (lists of 1-element tensors, token order = `config.token_dict` order) plus the fused extras."""
        if isinstance(attention_maps, torch.Tensor):
            res, T = attention_maps.shape[0], attention_maps.shape[-1]
            accs, n_maps = [attention_maps.reshape(1, res * res, T).float()], 1
        else:
            accs = [m.acc for m in attention_maps]
            n_maps = sum(m.n_maps for m in attention_maps)
            res, T = int(round(accs[0].shape[1] ** 0.5)), accs[0].shape[2]
        spec = self._tail_spec(res, T, smooth_attentions, sigma, kernel_size, normalize_eot, accs[0].device)
        attn_text, smoothed, stats, argmax, total = ops.guidance_tail(spec, accs, n_maps)

        n = len(spec.token_indices)
        is_box = [k == abi.GA_TOKEN_BOX for k in spec.kinds]
        losses_dict = {
            "max_loss": [stats[i, abi.GA_STAT_MAX] for i in range(n)],
            "col": [stats[i, abi.GA_STAT_COL:abi.GA_STAT_COL + 1] for i in range(n)],
            "row": [stats[i, abi.GA_STAT_ROW:abi.GA_STAT_ROW + 1] for i in range(n)],
            "inside_loss": [stats[i, abi.GA_STAT_INSIDE:abi.GA_STAT_INSIDE + 1] if is_box[i] else 0 for i in range(n)],
            "outside_loss": [stats[i, abi.GA_STAT_OUTSIDE:abi.GA_STAT_OUTSIDE + 1] if is_box[i] else 0
                             for i in range(n)],
            # fused extras (not in the reference dict)
            "_stats": stats, "_total": total, "_argmax": argmax, "_smoothed": smoothed, "_spec": spec,
            "attention_for_text": attn_text,
        }
        if hasattr(state.config, "custom_loss"):
            custom = torch.zeros(1, dtype=torch.float32, device=stats.device)
            for _, (loss_obj, args) in state.config.custom_loss.items():
                term = None
                if hasattr(loss_obj, "calc_loss_from_stats"):     # built-in losses: scalars from the tail's statistics
                    term = loss_obj.calc_loss_from_stats(stats, spec.token_indices, res, args)
                if term is None:                                  # plug-in contract of the reference (run.py:148-232)
                    term = loss_obj.calc_loss(attn_text, args)
                custom = custom + term
            losses_dict["custom_loss"] = custom
        if state.config.diagnostic_level > 0 or state.config.save_all_maps:
            self._log_maps(losses_dict)
        return losses_dict

    def _aggregate_and_get_max_attention_per_token(self, attention_store: AttentionStore, attention_res: int = 16,
                                                   smooth_attentions: bool = False, sigma: float = 0.5,
                                                   kernel_size: int = 3, normalize_eot: bool = False):
        """Aggregation + per-token statistics (reference :298-354): the per-layer accumulators go straight into the
        tail kernel, the (res, res, 77) mean is never materialised."""
        from_where = ("up", "down", "mid")
        picked = select_maps(attention_store, attention_res, from_where, True)
        if len(picked) == 0:
            raise RuntimeError("torch.cat(): expected a non-empty list of Tensors")   # reference ptp_utils.py:287
        return self._compute_max_attention_per_index(picked, smooth_attentions=smooth_attentions, sigma=sigma,
                                                     kernel_size=kernel_size, normalize_eot=normalize_eot)

    def _log_maps(self, losses_dict):
        """Opt-in diagnostics (the reference does this unconditionally inside the loss, :243-246, :273-274)."""
        stats = losses_dict["_stats"].detach().cpu()
        for n, idx in enumerate(losses_dict["_spec"].token_indices):
            helpers.log(f"{self.get_token(idx)}: weighted center col: {stats[n, abi.GA_STAT_COL].item()} "
                        f"row: {stats[n, abi.GA_STAT_ROW].item()}")

    @staticmethod
    def group_losses_by_sumprompt(losses):
        """(total, {sub-prompt: sum or mean}) -- reference :358-387."""
        groups: Dict[Any, list] = {}
        for idx, val in losses:
            sub = None if idx is None else state.config.token_dict[idx]['subprompt']
            groups.setdefault(sub, []).append(val)
        total, final = 0., {}
        for sub, vals in groups.items():
            tot = 0.
            for v in vals:
                tot = tot + (v / len(vals) if state.config.sub_prompt_avg_within else v)
            final[sub] = tot
            total = total + tot
        return total, final

    @staticmethod
    def get_centering_loss(center, losses_dict, i):
        """reference :390-395 with 16 -> res, 15 -> res - 1."""
        res = losses_dict["_spec"].res if "_spec" in losses_dict else 16
        part1 = 1. * (losses_dict["col"][i] - center[0] * res).abs() / (res - 1.)
        part2 = 4. * (losses_dict["row"][i] - center[1] * res).abs() / (res - 1.)
        return part1 + part2

    @staticmethod
    def _compute_loss(losses_dict: dict, return_losses: bool = False):
        """(loss, losses, unscaled_losses) -- reference :398-451.  The per-token terms and their weighted sum were
        already produced by the tail kernel; this only packages them in the reference's list-of-tuples form."""
        stats, spec = losses_dict["_stats"], losses_dict["_spec"]
        losses, unscaled = [], []
        for n, idx in enumerate(spec.token_indices):
            if spec.kinds[n] == abi.GA_TOKEN_KEYWORD:
                continue
            losses.append((idx, stats[n, abi.GA_STAT_SCALED:abi.GA_STAT_SCALED + 1]))
            unscaled.append((idx, stats[n, abi.GA_STAT_UNSCALED:abi.GA_STAT_UNSCALED + 1]))
        loss = losses_dict["_total"]
        if "custom_loss" in losses_dict:
            losses.append((None, losses_dict["custom_loss"]))
            unscaled.append((None, losses_dict["custom_loss"]))
            loss = loss + losses_dict["custom_loss"]
        return loss, losses, unscaled

    @staticmethod
    def _update_latent(latents: torch.Tensor, loss: torch.Tensor, step_size: float) -> torch.Tensor:
        """latents - step_size * d loss / d latents (reference :455-470); the backward runs K2 in every cross layer and
        the tail backward kernel once."""
        grad_cond = torch.autograd.grad(loss.requires_grad_(True), [latents], retain_graph=True)[0]
        GuidedAttention.update_calls += 1
        # fp32 arithmetic, one rounding: identical whether `step_size` is a python float or a device scalar (graphs)
        return (latents.float() - step_size * grad_cond.float()).to(latents.dtype)

    # --------------------------------------------------------------------------------------- threshold (host side)
    def meets_threshold(self, i, thresholds, losses):
        """True when every sub-prompt's summed unscaled loss is <= the threshold (reference :1074-1088).  This is the one
        place the loss values are read back to the host, and only on steps that own a threshold."""
        if (i not in thresholds and i != -1) or len(thresholds) == 0:
            return True
        thresh = list(thresholds.values())[-1] if i == -1 else thresholds[i]
        _, per_sub = GuidedAttention.group_losses_by_sumprompt(
            [(idx, float(v)) if not isinstance(v, float) else (idx, v) for idx, v in self._host_values(losses)])
        return all(not (v > thresh) for v in per_sub.values())

    @staticmethod
    def _host_values(losses):
        """One D2H copy for the whole list."""
        tensors = [v for _, v in losses if torch.is_tensor(v)]
        if not tensors:
            return list(losses)
        flat = torch.cat([t.detach().reshape(-1)[:1].float() for t in tensors]).cpu().tolist()
        it = iter(flat)
        return [(idx, next(it) if torch.is_tensor(v) else float(v)) for idx, v in losses]

    # -------------------------------------------------------------------------------------- iterative refinement
    def _perform_iterative_refinement_step(self, latents: torch.Tensor, loss: torch.Tensor, threshold: float,
                                           text_embeddings: torch.Tensor, text_input, attention_store: AttentionStore,
                                           step_size: float, t: int, attention_res: int = 16,
                                           smooth_attentions: bool = True, sigma: float = 0.5, kernel_size: int = 3,
                                           max_refinement_steps: int = 5, normalize_eot: bool = False):
        """Repeat [UNet forward (grad) -> loss -> latent update] until every sub-prompt meets the step's threshold or
        `max_refinement_steps` is hit, then one more forward for the loss at the refined latent (reference :475-581)."""
        self.inside_iterative_refinement = True
        use_optimizer = state.curHyperParams.get("use_optimizer", False)
        if use_optimizer:
            self.optim = torch.optim.SGD([latents], lr=step_size / 2.5, momentum=0.8)
        iteration = 0
        state.sub_iteration = iteration
        losses = unscaled_losses = None
        kw = dict(attention_store=attention_store, attention_res=attention_res, smooth_attentions=smooth_attentions,
                  sigma=sigma, kernel_size=kernel_size, normalize_eot=normalize_eot)
        while losses is None or not self.meets_threshold(state.cur_time_step_iter, state.config.thresholds,
                                                         unscaled_losses):
            helpers.log(f"subiteration: {iteration}")
            if use_optimizer:
                self.optim.zero_grad()
            iteration += 1
            state.sub_iteration = iteration
            if not use_optimizer:
                latents = latents.clone().detach().requires_grad_(True)
            self.unet(latents, t, encoder_hidden_states=text_embeddings[1].unsqueeze(0))
            self._count_pass("eval")
            losses_dict = self._aggregate_and_get_max_attention_per_token(**kw)
            loss, losses, unscaled_losses = self._compute_loss(losses_dict, return_losses=True)
            if use_optimizer:
                loss.backward()
                self.optim.step()
            elif self._nonzero(loss):
                latents = self._update_latent(latents, loss, step_size)
            if iteration >= max_refinement_steps:
                helpers.log(f'\t Exceeded max number of iterations ({max_refinement_steps})! ', True)
                break
        latents = latents.clone().detach().requires_grad_(True)
        self.unet(latents, t, encoder_hidden_states=text_embeddings[1].unsqueeze(0))
        self._count_pass("eval")
        max_attention_per_index = self._aggregate_and_get_max_attention_per_token(**kw)
        loss, losses, unscaled_losses = self._compute_loss(max_attention_per_index, return_losses=True)
        helpers.log(f"\t Finished with loss iter: {iteration}", True)
        state.sub_iteration = 0
        self.inside_iterative_refinement = False
        return loss, latents, max_attention_per_index

    @staticmethod
    def _nonzero(loss) -> bool:
        """`loss != 0` of the reference (:552, :1002).  With at least one BOX/COOR token the loss is a sum of
        non-negative terms that is zero only in degenerate cases; the exact test costs a host sync, so it is only
        evaluated when no token can contribute."""
        spec_tokens = state.config.token_dict
        if any(v['loss_type'] in (AT.COOR, AT.BOX) for v in spec_tokens.values()):
            return True
        return bool(loss != 0)

    # UNet-pass bookkeeping (both execution modes): eval = text-cond forward + loss; update = the backward to the
    # latents (+ its forward when replayed from a graph); cfg = the CFG forward + scheduler step
    def _count_pass(self, name, n=1):
        if not hasattr(self, "pass_counts"):
            self.pass_counts = {"eval": 0, "update": 0, "cfg": 0}
        self.pass_counts[name] += n

    # ------------------------------------------------------------------------------------- CUDA-graph execution
    use_cuda_graphs = False   # `run.load_model` (the `run.py --meta_prompt` entry point) and bench.py turn it on for CUDA

    def _step_graphs(self, attention_store, loss_kw, prompt_embeds, guidance_scale, latents) -> "_StepGraphs":
        """Graphs are cached on the pipeline and reused across calls (seeds) while everything baked into them is
        unchanged: UNet, shapes, dtype, tracked tokens, boxes and hyper-parameters, guidance scale, number of steps."""
        hp, cfg = state.curHyperParams, state.config
        key = (id(self.unet), id(attention_store), tuple(prompt_embeds.shape), prompt_embeds.dtype, tuple(latents.shape),
               float(guidance_scale), (loss_kw["attention_res"], loss_kw["smooth_attentions"], loss_kw["sigma"],
                                       loss_kw["kernel_size"], loss_kw["normalize_eot"]),
               _token_table_key(cfg.token_dict),
               tuple(sorted((k, str(v)) for k, v in hp.items())), bool(cfg.sub_prompt_avg_within),
               tuple(getattr(cfg, "custom_loss", {}).keys()), self.scheduler.num_inference_steps, str(latents.device))
        cached = getattr(self, "_graphs_cache", None)
        if cached is None or cached[0] != key:
            self._graphs_cache = (key, _StepGraphs(self, attention_store, loss_kw, prompt_embeds, guidance_scale,
                                                   latents))
        G = self._graphs_cache[1]
        G.set_embeds(prompt_embeds)
        return G

    def _denoise_graphed(self, G, latents, timesteps, thresholds, scale_range, scale_factor, recurse_steps,
                         recurse_until, max_iter_to_alter, run_standard_sd, renoise_gen, callback, callback_steps):
        """The guided loop of `__call__` (reference :925-1053) with every device program replayed from a CUDA graph.
        Host control flow, thresholds and random streams are exactly those of the eager loop below."""
        cfg = state.config
        for i, t in enumerate(timesteps):
            t = int(t)
            if i:
                ops.nvtx_pop()
            ops.nvtx_push("denoise_step %d (host-driven graphs)" % i)
            for recurse_step in range(0, recurse_steps):
                did_we_update = False
                state.cur_time_step_iter = i
                step_size = scale_factor * np.sqrt(scale_range[i])
                out, _ = G.run("eval", latents, t)
                if not run_standard_sd:
                    update_cond = (not cfg.only_update_on_threshold_steps and i < max_iter_to_alter) or \
                        (i in cfg.thresholds)
                    met, do_update = True, False
                    if i in thresholds or update_cond:
                        u0 = self._host_values(out["unscaled"])          # the one D2H read of this evaluation
                        met = self.meets_threshold(i, thresholds, u0)
                        do_update = update_cond and not self.meets_threshold(-1, cfg.thresholds, u0)
                    if not met:
                        did_we_update = True
                        iteration, u = 0, None
                        state.sub_iteration = 0
                        refine_prog = "update_mom" if G.use_optimizer else "update"
                        G.first.fill_(1.0)
                        while u is None or not self.meets_threshold(i, cfg.thresholds, u):
                            iteration += 1
                            state.sub_iteration = iteration
                            o, latents = G.run(refine_prog, latents, t, step_size=step_size)
                            G.first.fill_(0.0)
                            u = self._host_values(o["unscaled"])
                            if iteration >= 10:
                                helpers.log('\t Exceeded max number of iterations (10)! ', True)
                                break
                        state.sub_iteration = 0
                        # final evaluation at the refined latents; on a threshold step it carries the update of :1003
                        if do_update:
                            _, latents = G.run("update", latents, t, step_size=step_size)
                        else:
                            G.run("eval", latents, t)
                    elif do_update:
                        _, latents = G.run("update", latents, t, step_size=step_size)
                    did_we_update = did_we_update or do_update
                _, latents = G.run("cfg", latents, t, coeffs=self.scheduler.coefficients(t))
                if callback is not None and i % callback_steps == 0:
                    callback(i, t, latents)
                if i > recurse_until or not did_we_update:
                    break
                if recurse_step != (recurse_steps - 1):
                    latents = self._renoise(latents, t, renoise_gen)
        if len(timesteps):
            ops.nvtx_pop()
        return latents

    # ------------------------------------------------------------------------------- device-side control (f3)
    device_side_control = True    # with CUDA graphs: refinement / recursion / threshold tests run inside one graph launch
    control_mode = None           # what the last `__call__` used: "device", "host-graphs" or "eager"

    def _denoise_device(self, G, latents, timesteps, thresholds, scale_range, scale_factor, recurse_steps,
                        recurse_until, max_iter_to_alter, renoise_gen):
        """The guided loop of `__call__` (reference :925-1053) with the per-step control flow ON THE DEVICE
        (`csrc/step_driver.cu`): one graph launch per denoising step, no loss is read back, the host only feeds the
        step's scalars (timestep, step size, DDIM / re-noise coefficients, thresholds).  Same UNet passes in the same
        order and the same latents as `_denoise_graphed`.  The image's re-noise tensors are drawn up front from the
        same CPU generator stream, in the same order the host loop would draw them."""
        cfg = state.config
        lib = abi.load()
        drv = G.driver()
        dev = G.lat.device
        n_train = self.scheduler.config.num_train_timesteps
        ratio = n_train // self.scheduler.num_inference_steps

        def update_cond(i):
            return (not cfg.only_update_on_threshold_steps and i < max_iter_to_alter) or (i in cfg.thresholds)
        n_draws = 0
        for i, t in enumerate(timesteps):
            if ((i in thresholds) or update_cond(i)) and not (i > recurse_until) and int(t) - ratio > 0:
                n_draws += recurse_steps - 1
        if n_draws > G.max_draws:
            raise RuntimeError(f"{n_draws} re-noise draws exceed the captured capacity {G.max_draws}")
        if n_draws:
            noise = torch.stack([torch.randn(latents.shape, generator=renoise_gen, dtype=torch.float32)
                                 for _ in range(n_draws)]).pin_memory()
            G.noise[:n_draws].copy_(noise, non_blocking=True)
        G.n_draws.zero_()
        G.lat.copy_(latents)
        before = G.ctl.clone()
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        last_cfg = list(cfg.thresholds.values())[-1] if len(cfg.thresholds) else 0.0
        pww = PaintWithWords.of(G.store)
        plain_steps = 0
        for i, t in enumerate(timesteps):
            t = int(t)
            state.cur_time_step_iter = i
            if pww is not None:
                pww.update(dev)
            p = abi.GaStepParams()
            p.has_thr_call = int(i in thresholds and len(thresholds) > 0)
            p.thr_call = float(thresholds[i]) if p.has_thr_call else 0.0
            p.has_thr_cfg = int(i in cfg.thresholds and len(cfg.thresholds) > 0)
            p.thr_cfg = float(cfg.thresholds[i]) if p.has_thr_cfg else 0.0
            p.has_thr_last, p.thr_last = int(len(cfg.thresholds) > 0), float(last_cfg)
            p.update_cond = int(update_cond(i))
            p.check = int((i in thresholds) or bool(p.update_cond))
            p.recurse_ok = int(not (i > recurse_until))
            prev_t = t - ratio
            p.renoise_ok = int(prev_t > 0)
            p.recurse_steps, p.max_refine = int(recurse_steps), 10
            p.timestep = t
            p.step_size = float(scale_factor * np.sqrt(scale_range[i]))
            for n, c in enumerate(self.scheduler.coefficients(t)):
                p.ddim[n] = float(c)
            if prev_t > 0:
                Bt = float(self.scheduler.alphas_cumprod[t] / self.scheduler.alphas_cumprod[prev_t])
                p.renoise[0], p.renoise[1] = Bt ** 0.5, (1 - Bt) ** 0.5
            with torch.cuda.device(dev), ops.nvtx_range("denoise_step %d (device control)" % i):
                if p.check:
                    abi.check(lib.ga_step_driver_run(drv, C.byref(p), stream), "ga_step_driver_run")
                else:
                    # nothing to decide on this step (no threshold, no update): the reference still runs its
                    # text-conditioned forward and the CFG step (:946-947, :1008-1029) -- replay them directly, still
                    # without any read-back (a conditional node costs ~50 us per executed body)
                    abi.check(lib.ga_step_driver_set_params(drv, C.byref(p), stream), "ga_step_driver_set_params")
                    G.graphs["eval"].replay()
                    G.graphs["cfg"].replay()
                    G.graphs["advance"].replay()
                    plain_steps += 1
        out = G.lat.clone()
        # bookkeeping after the fact: one small read together with the image (the caller reads the latents anyway)
        base = abi.GA_STEP_COUNTER_BASE
        d = (G.ctl - before)[base:base + 6].cpu().tolist()
        d[abi.GA_STEP_N_EVAL] += plain_steps
        d[abi.GA_STEP_N_CFG] += plain_steps
        d[abi.GA_STEP_N_ROUNDS] += plain_steps
        n_prog = {"eval": d[abi.GA_STEP_N_EVAL], "update": d[abi.GA_STEP_N_UPDATE], "cfg": d[abi.GA_STEP_N_CFG]}
        self.last_step_counters = {"eval": d[0], "update": d[1], "cfg": d[2], "refine_iterations": d[3],
                                   "rounds": d[4], "renoise": d[5], "steps_without_decisions": plain_steps}
        for name, n in n_prog.items():
            self._count_pass(name, n)
        if G.use_optimizer:      # the refinement iterations ran the momentum program
            n_prog["update_mom"] = d[abi.GA_STEP_N_REFINE]
            n_prog["update"] -= d[abi.GA_STEP_N_REFINE]
        for name, n in n_prog.items():
            G.replays[name] = G.replays.get(name, 0) + n
            for k, v in G.launches[name].items():
                ops._count(k, v * n)
        return out

    def _renoise(self, latents, t, renoise_gen):
        """Back to the previous noise level before a recursion (reference :1046-1050)."""
        prev_timestep = t - self.scheduler.config.num_train_timesteps // self.scheduler.num_inference_steps
        if prev_timestep > 0:
            Bt = float(self.scheduler.alphas_cumprod[t] / self.scheduler.alphas_cumprod[prev_timestep])
            noise = torch.randn(latents.shape, generator=renoise_gen, dtype=torch.float32)
            latents = ((Bt ** 0.5) * latents.float()
                       + ((1 - Bt) ** 0.5) * noise.to(latents.device)).to(latents.dtype)
        return latents

    # ------------------------------------------------------------------------------------------ seed batching
    def _unscaled_groups(self, stats_row, spec):
        """{sub-prompt: summed (or averaged) unscaled loss} for one sample from a host copy of its stats rows."""
        groups = {}
        for n, sub in enumerate(spec.groups):
            if spec.kinds[n] == abi.GA_TOKEN_KEYWORD:
                continue
            groups.setdefault(sub, []).append(float(stats_row[n][abi.GA_STAT_UNSCALED]))
        avg = bool(state.config.sub_prompt_avg_within)
        return {k: (sum(v) / len(v) if avg else sum(v)) for k, v in groups.items()}

    @staticmethod
    def _met(i, thresholds, groups):
        if (i not in thresholds and i != -1) or len(thresholds) == 0:
            return True
        thresh = list(thresholds.values())[-1] if i == -1 else thresholds[i]
        return all(not (v > thresh) for v in groups.values())

    @torch.no_grad()
    def generate_batch(self, prompt, attention_store, seeds, prompt_embeds, negative_prompt_embeds,
                       attention_res: int = 16, num_inference_steps: int = 50, guidance_scale: float = 7.5,
                       max_iter_to_alter: int = 25, run_standard_sd: bool = False, thresholds: Optional[dict] = None,
                       scale_factor: int = 20, scale_range=(1., 0.5), smooth_attentions: bool = True,
                       sigma: float = 0.5, kernel_size: int = 3, sd_2_1: bool = False, latents=None,
                       max_refinement_steps: int = 10):
        """Extension: the guided loop of `__call__` for S seeds at once (one UNet pass serves all of them).  Per-seed
        semantics are those of S separate `__call__`s: each seed has its own initial noise and re-noise stream, its own
        threshold tests, refinement iteration count and recursion decisions; seeds that are done with a stage ride along
        with step size 0 / are masked out of the result.  Returns the final latents (S, C, h, w).  Keyword losses are
        supported when they can be computed from the tail's statistics (the built-in toLeftOf); Python plug-ins that
        need the materialised maps are not."""
        if (state.curHyperParams or {}).get("paint_with_words_stop", 0):
            raise NotImplementedError("generate_batch does not support paint-with-words: the reference's bias uses the "
                                      "max over the whole launch (utils/ptp_utils.py:138), which would couple the seeds")
        for _, (loss_obj, _args) in (getattr(state.config, "custom_loss", None) or {}).items():
            if not hasattr(loss_obj, "calc_loss_from_stats"):
                raise NotImplementedError("generate_batch supports [CustomLoss:...] only for losses with a fused "
                                          "calc_loss_from_stats (the built-in toLeftOf)")
        thresholds = dict(thresholds if thresholds is not None else state.config.thresholds)
        if len(thresholds) == 0:
            thresholds = {0: float("inf")}
        cfg = state.config
        device, dtype = self._execution_device, self.unet.dtype
        S = len(seeds)
        self.prompt = prompt
        embeds = torch.cat([negative_prompt_embeds, prompt_embeds]).to(device=device, dtype=dtype)
        self.scheduler = DDIMScheduler.from_config(self.scheduler.config)
        self.scheduler.set_timesteps(num_inference_steps)
        timesteps = self.scheduler.timesteps
        hw = self.unet.config.sample_size
        if latents is None:
            latents = torch.cat([torch.randn(1, self.unet.in_channels, hw, hw,
                                             generator=torch.Generator("cpu").manual_seed(int(sd))) for sd in seeds])
        latents = latents.to(device=device, dtype=dtype) * self.scheduler.init_noise_sigma
        scale = np.linspace(scale_range[0], scale_range[1], len(timesteps))
        recurse_steps = max(state.curHyperParams.get("recurse_steps", 1), 1)
        recurse_until = state.curHyperParams.get("recurse_until", 20)
        renoise = [torch.Generator("cpu").manual_seed(int(sd)) for sd in seeds] if recurse_steps > 1 else None

        loss_kw = dict(attention_store=attention_store, attention_res=attention_res,
                       smooth_attentions=smooth_attentions, sigma=sigma, kernel_size=kernel_size, normalize_eot=sd_2_1)
        key = ("batch", S, id(self.unet), id(attention_store), tuple(embeds.shape), embeds.dtype, float(guidance_scale),
               attention_res, smooth_attentions, sigma, kernel_size, sd_2_1, num_inference_steps,
               _token_table_key(cfg.token_dict),
               tuple(sorted((k, str(v)) for k, v in state.curHyperParams.items())), bool(cfg.sub_prompt_avg_within),
               bool(self.use_cuda_graphs), str(device))
        cached = getattr(self, "_batch_graphs_cache", None)
        if cached is None or cached[0] != key:
            self._batch_graphs_cache = (key, _BatchGraphs(self, attention_store, loss_kw, embeds, guidance_scale,
                                                          latents, eager=not self.use_cuda_graphs))
        G = self._batch_graphs_cache[1]
        G.set_embeds(embeds)

        def groups_of(out):
            rows = out["stats"].detach().cpu().tolist()       # the one D2H read of this evaluation
            spec = self._tail_spec_cache[1]
            groups = [self._unscaled_groups(rows[s_], spec) for s_ in range(S)]
            if hasattr(state.config, "custom_loss"):          # the reference appends (None, custom) even when it is 0
                cust = out["custom"].detach().cpu().tolist() if out.get("custom") is not None else [0.0] * S
                for s_ in range(S):
                    groups[s_][None] = groups[s_].get(None, 0.0) + float(cust[s_])
            return groups

        for i, t in enumerate(timesteps):
            t = int(t)
            state.cur_time_step_iter = i
            step_size = float(scale_factor * np.sqrt(scale[i]))
            live = [True] * S                        # seeds taking part in this recursion round
            for recurse_step in range(recurse_steps):
                if not any(live):
                    break
                did_update = [False] * S
                out, _ = G.run("eval", latents, t)
                if not run_standard_sd:
                    update_cond = (not cfg.only_update_on_threshold_steps and i < max_iter_to_alter) or \
                        (i in cfg.thresholds)
                    met, do_update = [True] * S, [False] * S
                    if i in thresholds or update_cond:
                        g0 = groups_of(out)
                        met = [self._met(i, thresholds, g0[s_]) or not live[s_] for s_ in range(S)]
                        do_update = [live[s_] and update_cond and not self._met(-1, cfg.thresholds, g0[s_])
                                     for s_ in range(S)]
                    refine = [not m for m in met]
                    need = list(refine)
                    iteration = 0
                    while any(need):
                        iteration += 1
                        o, latents = G.run("update", latents, t, step_sizes=[step_size if n_ else 0.0 for n_ in need])
                        gi = groups_of(o)
                        need = [need[s_] and not self._met(i, cfg.thresholds, gi[s_]) for s_ in range(S)]
                        if iteration >= max_refinement_steps:
                            break
                    if any(do_update):
                        # the final evaluation at the (refined) latents carries the threshold-step update (:998-1004)
                        _, latents = G.run("update", latents, t,
                                           step_sizes=[step_size if u else 0.0 for u in do_update])
                    elif any(refine):
                        G.run("eval", latents, t)
                    did_update = [refine[s_] or do_update[s_] for s_ in range(S)]
                _, stepped = G.run("cfg", latents, t, coeffs=self.scheduler.coefficients(t))
                if all(live):
                    latents = stepped
                else:
                    mask = torch.tensor(live, device=device).reshape(S, 1, 1, 1)
                    latents = torch.where(mask, stepped, latents)
                # seeds that updated (and are early enough) are re-noised and repeat this timestep; the others are done
                again = [live[s_] and did_update[s_] and i <= recurse_until and recurse_step != recurse_steps - 1
                         for s_ in range(S)]
                if any(again):
                    prev_timestep = t - self.scheduler.config.num_train_timesteps // self.scheduler.num_inference_steps
                    if prev_timestep > 0:
                        Bt = float(self.scheduler.alphas_cumprod[t] / self.scheduler.alphas_cumprod[prev_timestep])
                        noise = torch.zeros(latents.shape, dtype=torch.float32)
                        for s_ in range(S):
                            if again[s_]:
                                noise[s_] = torch.randn(latents.shape[1:], generator=renoise[s_], dtype=torch.float32)
                        mixed = ((Bt ** 0.5) * latents.float() + ((1 - Bt) ** 0.5) * noise.to(device)).to(dtype)
                        mask = torch.tensor(again, device=device).reshape(S, 1, 1, 1)
                        latents = torch.where(mask, mixed, latents)
                live = again
        return latents.detach()

    # ---------------------------------------------------------------------------------------------------- call
    @torch.no_grad()
    def __call__(self, prompt: Union[str, List[str]], attention_store: AttentionStore, attention_res: int = 16,
                 height: Optional[int] = None, width: Optional[int] = None, num_inference_steps: int = 50,
                 guidance_scale: float = 7.5, negative_prompt: Optional[Union[str, List[str]]] = None,
                 num_images_per_prompt: Optional[int] = 1, eta: float = 0.0,
                 generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
                 latents: Optional[torch.Tensor] = None, prompt_embeds: Optional[torch.Tensor] = None,
                 negative_prompt_embeds: Optional[torch.Tensor] = None, output_type: Optional[str] = "pil",
                 return_dict: bool = True, callback: Optional[Callable[[int, int, torch.Tensor], None]] = None,
                 callback_steps: Optional[int] = 1, cross_attention_kwargs: Optional[Dict[str, Any]] = None,
                 max_iter_to_alter: Optional[int] = 25, run_standard_sd: bool = False,
                 thresholds: Optional[dict] = {0: 0.05, 10: 0.5, 20: 0.8}, scale_factor: int = 20,
                 scale_range: Tuple[float, float] = (1., 0.5), smooth_attentions: bool = True, sigma: float = 0.5,
                 kernel_size: int = 3, sd_2_1: bool = False):
        """Same arguments as the reference `__call__` (:746-777).  `output_type="latent"` returns the final latents
        tensor in `.images` (no VAE exists offline; "pil"/"np" go through the placeholder decode)."""
        height = height or self.unet.config.sample_size * self.vae_scale_factor
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        self.check_inputs(prompt, height, width, callback_steps, negative_prompt, prompt_embeds, negative_prompt_embeds)

        self.prompt = prompt
        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None and isinstance(prompt, list):
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        device = self._execution_device
        do_classifier_free_guidance = guidance_scale > 1.0
        text_inputs, prompt_embeds = self._encode_prompt(prompt, device, num_images_per_prompt,
                                                         do_classifier_free_guidance, negative_prompt,
                                                         prompt_embeds=prompt_embeds,
                                                         negative_prompt_embeds=negative_prompt_embeds)
        state.always_save_iter = [0, 1, 2]
        self.scheduler = DDIMScheduler.from_config(self.scheduler.config)
        self.scheduler.set_timesteps(num_inference_steps, device=device)
        state.sigmas = (((1 - self.scheduler.alphas_cumprod) / self.scheduler.alphas_cumprod) ** 0.5).numpy()
        timesteps = self.scheduler.timesteps
        state.timesteps = timesteps

        latents = self.prepare_latents(batch_size * num_images_per_prompt, self.unet.in_channels, height, width,
                                       prompt_embeds.dtype, device, generator, latents)
        extra_step_kwargs = self.prepare_extra_step_kwargs(generator, eta)
        scale_range = np.linspace(scale_range[0], scale_range[1], len(self.scheduler.timesteps))
        if max_iter_to_alter is None:
            max_iter_to_alter = len(self.scheduler.timesteps) + 1

        recurse_steps = max(state.curHyperParams.get("recurse_steps", 1), 1)
        recurse_until = state.curHyperParams.get("recurse_until", 20)
        if len(thresholds) == 0:
            thresholds = {0: float("inf")}
        renoise_gen = None
        if recurse_steps > 1:
            seed = generator.initial_seed() if generator is not None else 0
            renoise_gen = torch.Generator("cpu").manual_seed(seed)   # CPU stream: reproducible on every device

        loss_kw = dict(attention_store=attention_store, attention_res=attention_res,
                       smooth_attentions=smooth_attentions, sigma=sigma, kernel_size=kernel_size, normalize_eot=sd_2_1)
        num_warmup_steps = len(timesteps) - num_inference_steps * self.scheduler.order
        graphed = (self.use_cuda_graphs and latents.is_cuda and state.config.diagnostic_level == 0
                   and do_classifier_free_guidance and latents.shape[0] == 1 and cross_attention_kwargs is None)
        attention_store.text_kv = None      # the eager loop does not own the embedding buffers: no K/V caching
        self.control_mode = "eager"
        if graphed:
            G = self._step_graphs(attention_store, loss_kw, prompt_embeds, guidance_scale, latents)
            if self.device_side_control and callback is None and not run_standard_sd:
                self.control_mode = "device"
                if renoise_gen is None:
                    renoise_gen = torch.Generator("cpu").manual_seed(0)      # recurse_steps == 1: never drawn from
                latents = self._denoise_device(G, latents, timesteps, thresholds, scale_range, scale_factor,
                                               recurse_steps, recurse_until, max_iter_to_alter, renoise_gen)
            else:
                self.control_mode = "host-graphs"
                latents = self._denoise_graphed(G, latents, timesteps, thresholds, scale_range, scale_factor,
                                                recurse_steps, recurse_until, max_iter_to_alter, run_standard_sd,
                                                renoise_gen, callback, callback_steps)
            timesteps = []      # the eager loop below has nothing left to do
        with self.progress_bar(total=num_inference_steps) as progress_bar:
            for i, t in enumerate(timesteps):
                t = int(t)
                if i:
                    ops.nvtx_pop()
                ops.nvtx_push("denoise_step %d (eager)" % i)
                for recurse_step in range(0, recurse_steps):
                    did_we_update = False
                    state.cur_time_step_iter = i
                    helpers.log(f"iteration {i}", True)
                    with torch.enable_grad():
                        latents = latents.clone().detach().requires_grad_(True)
                        # text-conditioned forward with the autograd graph: fills the attention accumulators
                        self.unet(latents, t, encoder_hidden_states=prompt_embeds[1].unsqueeze(0),
                                  cross_attention_kwargs=cross_attention_kwargs)
                        self._count_pass("eval")
                        max_attention_per_index = self._aggregate_and_get_max_attention_per_token(**loss_kw)
                        if not run_standard_sd:
                            loss, losses, unscaled_losses = self._compute_loss(losses_dict=max_attention_per_index)
                            if not self.meets_threshold(i, thresholds, unscaled_losses):
                                did_we_update = True
                                loss, latents, max_attention_per_index = self._perform_iterative_refinement_step(
                                    latents=latents, loss=loss, threshold=thresholds[i], text_embeddings=prompt_embeds,
                                    text_input=text_inputs, attention_store=attention_store,
                                    step_size=scale_factor * np.sqrt(scale_range[i]), t=t,
                                    attention_res=attention_res, smooth_attentions=smooth_attentions,
                                    max_refinement_steps=10, sigma=sigma, kernel_size=kernel_size, normalize_eot=sd_2_1)
                            if (not state.config.only_update_on_threshold_steps and i < max_iter_to_alter) or \
                                    (i in state.config.thresholds):
                                # NB: like the reference (:1001) this tests the PRE-refinement unscaled losses
                                if not self.meets_threshold(-1, state.config.thresholds, unscaled_losses):
                                    did_we_update = True
                                    loss, losses, unscaled_losses = self._compute_loss(
                                        losses_dict=max_attention_per_index)
                                    if self._nonzero(loss):
                                        latents = self._update_latent(latents=latents, loss=loss,
                                                                      step_size=scale_factor * np.sqrt(scale_range[i]))
                    latent_model_input = torch.cat([latents] * 2) if do_classifier_free_guidance else latents
                    latent_model_input = self.scheduler.scale_model_input(latent_model_input, t)
                    noise_pred = self.unet(latent_model_input, t, encoder_hidden_states=prompt_embeds,
                                           cross_attention_kwargs=cross_attention_kwargs).sample
                    self._count_pass("cfg")
                    if do_classifier_free_guidance:
                        noise_pred_uncond, noise_pred_text = noise_pred.chunk(2)
                        noise_pred = noise_pred_uncond + guidance_scale * (noise_pred_text - noise_pred_uncond)
                    ddim_output = self.scheduler.step(noise_pred, t, latents, **extra_step_kwargs)
                    latents = ddim_output.prev_sample
                    if state.config.diagnostic_level > 0:
                        self.save_image(ddim_output.pred_original_sample, "pred")
                    if i == len(timesteps) - 1 or ((i + 1) > num_warmup_steps and (i + 1) % self.scheduler.order == 0):
                        progress_bar.update()
                        if callback is not None and i % callback_steps == 0:
                            callback(i, t, latents)
                    if i > recurse_until or not did_we_update:
                        break
                    if recurse_step != (recurse_steps - 1):
                        latents = self._renoise(latents, t, renoise_gen)
            if len(timesteps):
                ops.nvtx_pop()

        latents = latents.detach()
        has_nsfw_concept = False
        if output_type == "latent":
            image = latents
        else:
            image = self.decode_latents(latents)
            if output_type == "pil":
                image = self.numpy_to_pil(image)
        if not return_dict:
            return (image, has_nsfw_concept)
        return StableDiffusionPipelineOutput(images=image, nsfw_content_detected=has_nsfw_concept)

    # ------------------------------------------------------------------------------------------------ diagnostics
    def get_innermost_folder(self):
        return str(state.cur_seed)

    def get_token(self, index):
        return self.tokenizer.decode(self.tokenizer(state.config.prompt)['input_ids'][index])

    def save_image(self, latent, tag):
        image = self.numpy_to_pil(self.decode_latents(latent.detach()))
        fname = helpers.get_meta_prompt_clean() + state.get_name() + "_" + tag
        for ch in "[]:.":
            fname = fname.replace(ch, "_")
        out = state.config.output_path / helpers.get_inner_folder_name() / self.get_innermost_folder()
        out.mkdir(exist_ok=True, parents=True)
        image[0].save(out / (fname + ".png"))
