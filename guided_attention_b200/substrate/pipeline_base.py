"""Minimal stand-in for diffusers' `StableDiffusionPipeline` base: only the members the guided pipeline calls
(reference `pipeline_guided_attention.py:850, 865, 883, 894-906, 924, 1060, 1067`).

No checkpoints exist offline: the text encoder and the VAE are out of scope, prompts are
supplied as synthetic `prompt_embeds`, and `decode_latents` is a fixed, parameter-free placeholder so that the
pipeline's return type keeps its shape.
"""
from __future__ import annotations

import contextlib
from types import SimpleNamespace

import numpy as np
import torch

from .ddim import DDIMScheduler
from .tokenizer import WhitespaceTokenizer


class _NullProgress:
    def update(self, n=1):
        pass


class StableDiffusionPipelineBase:
    vae_scale_factor = 8

    def __init__(self, unet=None, scheduler=None, tokenizer=None, text_encoder=None, vae=None):
        self.unet = unet
        self.scheduler = scheduler if scheduler is not None else DDIMScheduler()
        self.tokenizer = tokenizer if tokenizer is not None else WhitespaceTokenizer()
        self.text_encoder = text_encoder if text_encoder is not None else SimpleNamespace(
            dtype=unet.dtype if unet is not None else torch.float32, config=SimpleNamespace())
        self.vae = vae

    @property
    def _execution_device(self):
        return self.unet.device

    def to(self, device):
        self.unet.to(device)
        return self

    def check_inputs(self, prompt, height, width, callback_steps, negative_prompt=None, prompt_embeds=None,
                     negative_prompt_embeds=None):
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if callback_steps is None or not isinstance(callback_steps, int) or callback_steps <= 0:
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps}.")
        if prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`.")

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator,
                        latents=None):
        shape = (batch_size, num_channels_latents, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if latents is None:
            # always draw on the CPU generator then copy: reproducible on every device
            if generator is not None and generator.device.type != "cpu":
                seed = generator.initial_seed()
                generator = torch.Generator("cpu").manual_seed(seed)
            latents = torch.randn(shape, generator=generator, dtype=torch.float32)
        latents = latents.to(device=device, dtype=dtype)
        return latents * self.scheduler.init_noise_sigma

    def prepare_extra_step_kwargs(self, generator, eta):
        return {"eta": eta, "generator": generator}

    @contextlib.contextmanager
    def progress_bar(self, total=None):
        yield _NullProgress()

    def decode_latents(self, latents):
        """Placeholder decode (no VAE offline): 8x nearest upsample of the first three latent channels."""
        x = latents.detach().float() / 0.18215
        x = torch.nn.functional.interpolate(x[:, :3], scale_factor=8.0, mode="nearest")
        x = (x / 2 + 0.5).clamp(0, 1)
        return x.cpu().permute(0, 2, 3, 1).numpy()

    @staticmethod
    def numpy_to_pil(images):
        from PIL import Image
        if images.ndim == 3:
            images = images[None]
        images = (images * 255).round().astype("uint8")
        return [Image.fromarray(im) for im in images]


class StableDiffusionPipelineOutput:
    def __init__(self, images, nsfw_content_detected=False):
        self.images = images
        self.nsfw_content_detected = nsfw_content_detected
