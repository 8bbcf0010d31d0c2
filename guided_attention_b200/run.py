"""Driver -- mirror of the reference's `run.py`: `--meta_prompt` bracket syntax, `parseMetaPrompt`, `run_on_prompt`,
`execute`, the custom-loss plug-in API (`CustomLossBase`, `ToLeftOf`, `register_custom_loss`).

Differences that follow from the offline setting (documented in DESIGN.md): no checkpoint can be loaded, so `setup`
builds a random-init SD-shaped UNet (`substrate.build_unet`) and prompts are encoded as synthetic embeddings seeded from
the prompt; pyrallis is absent so the same RunConfig fields are exposed through argparse (`python -m
guided_attention_b200.run --meta_prompt "a [robot:.6,.3,.4,.55] and a [blue vase:.2,.3,.4,.55]" --seeds 28
--half_precision True`).
"""
from __future__ import annotations

import argparse
import ast
import zlib
from abc import ABC, abstractmethod
from pathlib import Path
from typing import List

import torch

from . import helpers, ptp_utils, shared_state
from .config import RunConfig
from .pipeline_guided_attention import GuidedAttention
from .ptp_utils import AttentionStore
from .substrate import DDIMScheduler, UNetConfig, WhitespaceTokenizer, build_unet


def load_model(config: RunConfig, device=None, unet_config: UNetConfig = None, unet_seed: int = 0):
    """Random-init stand-in for `GuidedAttention.from_pretrained` (reference run.py:18-29)."""
    device = torch.device(device) if device is not None else torch.device("cuda:0")
    if unet_config is None:
        unet_config = UNetConfig.sd21_base() if config.sd_2_1 else UNetConfig.sd14()
    dtype = torch.float16 if config.half_precision else torch.float32
    unet = build_unet(unet_config, seed=unet_seed, dtype=dtype, device=device)
    pipe = GuidedAttention(unet=unet, scheduler=DDIMScheduler(), tokenizer=WhitespaceTokenizer())
    # the guided loop replays three captured device programs instead of re-issuing ~10^3 launches per UNet pass; the
    # eager loop stays available (`pipe.use_cuda_graphs = False`, diagnostics, use_optimizer=False only for graphs)
    pipe.use_cuda_graphs = device.type == "cuda"
    return pipe


def synthetic_prompt_embeds(prompt: str, cross_attention_dim: int, n_ctx: int = 77, seed: int = 1234):
    """(2, n_ctx, dim) fp32 on CPU: row 0 = unconditional, row 1 = conditional.  The conditional
    row is drawn from a generator seeded by (seed, prompt) so different prompts get different embeddings."""
    g0 = torch.Generator("cpu").manual_seed(seed)
    uncond = torch.randn(1, n_ctx, cross_attention_dim, generator=g0)
    g1 = torch.Generator("cpu").manual_seed((seed * 1000003 + zlib.crc32(prompt.encode("utf-8"))) % (2 ** 31))
    cond = torch.randn(1, n_ctx, cross_attention_dim, generator=g1)
    return torch.cat([uncond, cond])


def run_on_prompt(prompt: List[str], model: GuidedAttention, controller: AttentionStore, seed: torch.Generator,
                  config: RunConfig, prompt_embeds=None, output_type="pil"):
    """reference run.py:44-67."""
    if controller is not None:
        ptp_utils.register_attention_control(model, controller)
    if prompt_embeds is None:
        text = prompt[0] if isinstance(prompt, list) else prompt
        prompt_embeds = synthetic_prompt_embeds(text, model.unet.config.cross_attention_dim)
    outputs = model(prompt=prompt, attention_store=controller, attention_res=config.attention_res,
                    guidance_scale=config.guidance_scale, generator=seed,
                    num_inference_steps=config.n_inference_steps, max_iter_to_alter=config.max_iter_to_alter,
                    run_standard_sd=config.run_standard_sd, thresholds=config.thresholds,
                    scale_factor=config.scale_factor, scale_range=config.scale_range,
                    smooth_attentions=config.smooth_attentions, sigma=config.sigma, kernel_size=config.kernel_size,
                    sd_2_1=config.sd_2_1, prompt_embeds=prompt_embeds[1:2], negative_prompt_embeds=prompt_embeds[0:1],
                    output_type=output_type)
    return outputs.images[0] if output_type != "latent" else outputs.images


def get_indices(tokenized_prompt, tokens):
    """First position where `tokens` occurs in `tokenized_prompt`; None if absent (reference run.py:69-73, including
    its range that never tests the final window)."""
    n = len(tokens)
    for i in range(0, len(tokenized_prompt) - n):
        if tokenized_prompt[i:i + n] == tokens:
            return list(range(i, i + n))


def overrideConfig(config):
    if 'meta_prompt' in shared_state.curHyperParams:
        config.meta_prompt = shared_state.curHyperParams['meta_prompt']
    if 'thresholds' in shared_state.curHyperParams:
        config.thresholds = shared_state.curHyperParams["thresholds"]


def parseMetaPrompt(config):
    """meta-prompt -> config.prompt / meta_info / custom_loss / token_dict (reference run.py:81-91)."""
    config.prompt, config.meta_info, config.custom_loss = helpers.parse_prompt(config.meta_prompt)
    shared_state.config = config
    tok = config.stable.tokenizer
    tokenized_prompt = tok(config.prompt)['input_ids']
    token_dict = {}
    for sub, kind, payload in config.meta_info:
        ids = tok(sub)['input_ids'][1:-1]
        for idx in get_indices(tokenized_prompt, ids):
            token_dict[idx] = {'word': tok.decode(tokenized_prompt[idx]), 'loss_type': kind, 'loss': payload,
                               'subprompt': sub}
    config.token_dict = token_dict


def execute(config, output_type="pil", save=True):
    """All seeds x hyper-parameter states (reference run.py:93-135).  Launched plainly, every seed runs sequentially on
    this process's GPU like the reference; launched under `torchrun` (one process per GPU), `config.seeds` is sharded
    round-robin `seed_idx % world_size` (`sweep.shard_seeds`, BASELINE config 5) with no collective on the hot path, and
    for `output_type="latent"` the per-seed results are gathered once at the end so that every rank returns the full
    list in (seed, hyper-state) order.  One `AttentionStore` serves every seed (`reset()` between images, which is all
    a fresh store would give) so the CUDA graphs captured for the first image are replayed for the rest."""
    from . import sweep
    rank, world, _ = sweep.init_distributed() if sweep.dist_env()[1] > 1 else (0, 1, 0)
    images = []
    image_path = None
    controller = AttentionStore()
    n_states = len(shared_state.get_hyperparam_states())
    for seed in sweep.shard_seeds(list(config.seeds), rank, world):
        for hyper in shared_state.get_hyperparam_states():
            shared_state.curHyperParams = hyper
            overrideConfig(config)
            parseMetaPrompt(config)
            helpers.log_clear()
            shared_state.cur_seed = seed
            g = torch.Generator('cpu').manual_seed(seed)   # CPU stream: same latents on every device
            controller.reset()
            image = run_on_prompt(prompt=config.prompt, model=config.stable, controller=controller, seed=g,
                                  config=config, output_type=output_type)
            if save and output_type == "pil":
                out_dir = config.output_path / helpers.get_inner_folder_name()
                out_dir.mkdir(exist_ok=True, parents=True)
                name = helpers.dictToString(shared_state.curHyperParams)
                image_path = out_dir / f'{seed}{name}.png'
                try:
                    image.save(image_path)
                except OSError:
                    image_path = out_dir / f'{seed}.png'
                    image.save(image_path)
                helpers.log_save(out_dir / f'{seed}.txt')
            images.append(image)
    if output_type == "latent" and world > 1:
        local = torch.cat(images) if images else None
        # the only communication of the sweep: final latents, 32 KB per image (NCCL over NVLink on the GPU box)
        like = local if local is not None else torch.zeros((0, config.stable.unet.in_channels) + (
            config.stable.unet.config.sample_size,) * 2, dtype=config.stable.unet.dtype,
            device=config.stable._execution_device)
        per_seed = like.reshape((-1, n_states) + tuple(like.shape[1:]))
        full = sweep.gather_results(per_seed, len(config.seeds), rank, world)
        return list(full.reshape((-1, 1) + tuple(like.shape[1:])))
    return images if output_type != "pil" else image_path


def setup(config, device=None, unet_config=None):
    shared_state.config = config
    config.stable = load_model(config, device=device, unet_config=unet_config)


# ------------------------------------------------------------------------------------------ custom-loss plug-ins
class CustomLossBase(ABC):
    """Plug-in contract of the reference (run.py:148-171): `calc_loss(cross_attention_maps, text_args)` receives the
    renormalised maps (res, res, n_text_tokens) -- here a differentiable fp32 CUDA tensor produced by the tail kernel."""

    @abstractmethod
    def calc_loss(self, cross_attention_maps, text_args: str) -> torch.Tensor:
        pass

    def subprompts_of_interest(self, text_args: str) -> list:
        return []

    def parse_text_args(self, text_args: str):
        return ast.literal_eval(text_args)

    def find_indices_for_sub_prompt(self, sub_prompt):
        tok = shared_state.config.stable.tokenizer
        full = tok(shared_state.config.prompt)['input_ids'][1:-1]
        sub = tok(sub_prompt)['input_ids'][1:-1]
        for i in range(len(full) - len(sub) + 1):
            if full[i:i + len(sub)] == sub:
                return list(range(i, i + len(sub)))

    def get_map_for_token(self, cross_attention_maps, token_index: int, pixel_wise_normalization: True):
        image_map = cross_attention_maps[:, :, token_index]
        if pixel_wise_normalization:
            image_map = image_map / image_map.sum()
        return image_map


class ToLeftOf(CustomLossBase):
    """`[CustomLoss:toLeftOf (a,b)]`: 9 * max(0, (cx_a + 0.2 W - cx_b) / W) (reference run.py:174-225), vectorised; keeps
    the reference's normalisation of the right-hand centre by len(left)."""

    def calc_loss(self, cross_attention_maps, text_args: str) -> torch.Tensor:
        args = self.parse_text_args(self.quote_items_in_tuple(text_args))
        left = self.find_indices_for_sub_prompt(args[0])
        right = self.find_indices_for_sub_prompt(args[1])
        width = cross_attention_maps.shape[1]
        left_c = sum(self.calc_weighted_center(self.get_map_for_token(cross_attention_maps, i, True))[0] / len(left)
                     for i in left)
        right_c = sum(self.calc_weighted_center(self.get_map_for_token(cross_attention_maps, i, True))[0] / len(left)
                      for i in right)
        loss = (left_c + .2 * width - right_c) / width * 9
        return torch.clamp(loss, min=0).reshape(1)

    def calc_loss_from_stats(self, stats, token_indices, res: int, text_args: str) -> torch.Tensor:
        """Fused path: the same loss from the tail kernel's per-token raw-map statistics -- the centre of
        mass of the un-smoothed map of every tracked token is `stats[n, GA_STAT_RAW_COL]`, differentiable through the
        tail backward kernel -- instead of reductions over the materialised (res, res, 75) maps.  `token_indices[n]` is
        the token index of stats row n (token index = map column + 1).  Returns None when a token of the two
        sub-prompts is not tracked (the caller then falls back to `calc_loss`)."""
        from . import _cabi as abi
        args = self.parse_text_args(self.quote_items_in_tuple(text_args))
        left = self.find_indices_for_sub_prompt(args[0])
        right = self.find_indices_for_sub_prompt(args[1])
        rows = {idx - 1: n for n, idx in enumerate(token_indices)}
        if any(i not in rows for i in left + right):
            return None
        cx = stats[:, abi.GA_STAT_RAW_COL]
        left_c = sum(cx[rows[i]] / len(left) for i in left)
        right_c = sum(cx[rows[i]] / len(left) for i in right)      # the reference divides by len(left) here too
        loss = (left_c + .2 * res - right_c) / res * 9
        return torch.clamp(loss, min=0).reshape(1)

    def subprompts_of_interest(self, text_args: str) -> list:
        return list(self.parse_text_args(self.quote_items_in_tuple(text_args)))

    def quote_items_in_tuple(self, text_args):
        items = text_args.strip('()').split(',')
        return "(" + ",".join(f"'{item.strip()}'" for item in items) + ")"

    def calc_weighted_center(self, imageNormalized):
        res_y, res_x = imageNormalized.shape
        xs = torch.arange(res_x, dtype=imageNormalized.dtype, device=imageNormalized.device) + .5
        ys = torch.arange(res_y, dtype=imageNormalized.dtype, device=imageNormalized.device) + .5
        return (imageNormalized * xs[None, :]).sum(), (imageNormalized * ys[:, None]).sum()


def register_custom_loss(name: str, customLoss: CustomLossBase):
    if not hasattr(shared_state.config, "registered_loss_functions"):
        shared_state.config.registered_loss_functions = {}
    shared_state.config.registered_loss_functions[name] = customLoss


def _str2bool(v):
    return str(v).lower() in ("1", "true", "yes", "y")


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--meta_prompt", required=True)
    ap.add_argument("--seeds", type=int, nargs="+", default=[42])
    ap.add_argument("--sd_2_1", type=_str2bool, default=False)
    ap.add_argument("--half_precision", type=_str2bool, default=False)
    ap.add_argument("--output_path", type=Path, default=Path("./outputs"))
    ap.add_argument("--n_inference_steps", type=int, default=50)
    ap.add_argument("--guidance_scale", type=float, default=7.5)
    ap.add_argument("--attention_res", type=int, default=16)
    ap.add_argument("--run_standard_sd", type=_str2bool, default=False)
    ap.add_argument("--diagnostic_level", type=int, default=0)
    a = ap.parse_args(argv)
    config = RunConfig(meta_prompt=a.meta_prompt, seeds=a.seeds, sd_2_1=a.sd_2_1, half_precision=a.half_precision,
                       output_path=a.output_path, n_inference_steps=a.n_inference_steps,
                       guidance_scale=a.guidance_scale, attention_res=a.attention_res,
                       run_standard_sd=a.run_standard_sd, diagnostic_level=a.diagnostic_level)
    setup(config)
    register_custom_loss("toLeftOf", ToLeftOf())
    print(execute(config))



def setup_prompt(meta_prompt: str, hyper=None, cfg_kw=None, register: bool = True):
    """`setup` + `parseMetaPrompt` of the reference (run.py:81-91, 139-145) for stacks without a CLIP tokenizer: installs a
    RunConfig with the whitespace tokenizer as `shared_state.config`, the shipped hyper-parameters (plus `hyper`
    overrides) as `shared_state.curHyperParams`, and builds `config.token_dict`.  Used by the microbenchmarks and, through
    tests/gpu_harness.py, by the tests."""
    import tempfile
    import types
    from .config import RunConfig
    from .substrate import WhitespaceTokenizer
    cfg = RunConfig(meta_prompt=meta_prompt, output_path=tempfile.mkdtemp(prefix="ga_run_"), **(cfg_kw or {}))
    shared_state.config = cfg
    cfg.stable = types.SimpleNamespace(tokenizer=WhitespaceTokenizer())
    if register:
        register_custom_loss("toLeftOf", ToLeftOf())
    hp = shared_state.get_hyperparam_states()[0]
    hp.update(hyper or {})
    shared_state.curHyperParams = hp
    overrideConfig(cfg)
    parseMetaPrompt(cfg)
    shared_state.cur_time_step_iter, shared_state.cur_seed, shared_state.sub_iteration = 0, 0, 0
    return cfg

if __name__ == '__main__':
    main()
