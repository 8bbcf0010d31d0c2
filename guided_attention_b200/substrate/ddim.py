"""DDIM scheduler restatement (the members the reference pipeline touches, `pipeline_guided_attention.py:883-888,
1011, 1027-1029, 1046-1050`).

diffusers is absent offline, so this is a *definition* shared by the oracle and the product path.  Settings follow
what `DDIMScheduler.from_config(<SD-1.4 PNDM config>)` is assumed to yield in diffusers 0.12.1 (SURVEY.md 8c [memory]):
scaled_linear betas 0.00085 -> 0.012 over 1000 train steps, steps_offset=1, set_alpha_to_one=False, eta=0, epsilon
prediction, and clip_sample=False: the checkpoints' own scheduler_config.json files (CompVis/stable-diffusion-v1-4,
stabilityai/stable-diffusion-2-1-base) carry `"clip_sample": false` and `from_config` honours the key [memory, not
checkable offline], so latent-space pred_x0 is NOT clamped.  The key stays overridable (`DDIMScheduler(clip_sample=True)`
gives the class default of diffusers' DDIMScheduler) and `from_config` passes it through when the config has it.  For 50 steps: timesteps 981, 961, ..., 1
(matches the comment at reference `utils/shared_state.py:8`).
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace

import numpy as np
import torch


@dataclass
class DDIMOutput:
    prev_sample: torch.Tensor
    pred_original_sample: torch.Tensor


class DDIMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1,
                 clip_sample=False, set_alpha_to_one=False):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, steps_offset=steps_offset, clip_sample=clip_sample,
                                      set_alpha_to_one=set_alpha_to_one, beta_schedule="scaled_linear")
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    @classmethod
    def from_config(cls, config):
        kw = {k: getattr(config, k) for k in ("num_train_timesteps", "beta_start", "beta_end", "steps_offset",
                                               "clip_sample", "set_alpha_to_one") if hasattr(config, k)}
        return cls(**kw)

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        # kept on the host on purpose: the loop reads `int(t)` every step and a device tensor would sync
        self.timesteps = torch.from_numpy(ts)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def coefficients(self, timestep):
        """(sqrt(1 - a_t), sqrt(a_t), sqrt(a_prev), sqrt(1 - a_prev)) as python floats for this step."""
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_prev = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else float(self.final_alpha_cumprod)
        return ((1 - a_t) ** 0.5, a_t ** 0.5, a_prev ** 0.5, (1 - a_prev) ** 0.5)

    def step(self, model_output, timestep, sample, eta=0.0, generator=None, coeffs=None, **_):
        """eta = 0 DDIM update.  The arithmetic runs in fp32 and is rounded to the sample dtype once; `coeffs` may be
        python floats (eager, no device sync) or four fp32 device scalars (CUDA-graph replay: same kernels, the step's
        coefficients are data instead of baked-in constants) -- both give bit-identical results."""
        c = self.coefficients(timestep) if coeffs is None else coeffs
        x, eps = sample.float(), model_output.float()
        pred_x0 = (x - c[0] * eps) / c[1]
        if self.config.clip_sample:
            pred_x0 = pred_x0.clamp(-1, 1)
        prev = c[2] * pred_x0 + c[3] * eps
        return DDIMOutput(prev_sample=prev.to(sample.dtype), pred_original_sample=pred_x0.to(sample.dtype))
