"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference sources from /root/reference on CPU.

Only `oracle/gen_golden.py` (fixture generation, run in the build container) and `tests/` (when /root/reference is
mounted) may import this.  Nothing in the product package, `bench.py` or `smoke()` imports it: /root/reference does
not exist on the GPU box.  No reference source is copied; the modules are imported where they lie.

Recipe: pre-seed `sys.modules` with placeholders for the third-party packages the
reference imports but that are absent offline (diffusers, matplotlib, IPython, pyrallis), add the names transformers
5.x dropped, and neutralise `.cuda()`.  The diffusers placeholders are backed by this repo's substrate
(`guided_attention_b200.substrate`): the substrate is the shared *definition* of everything diffusers owns, so the
reference's own pipeline `__call__` can run end to end on CPU against the very same UNet/DDIM the product path uses.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("GA_REFERENCE_ROOT", "/root/reference")

_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pipeline_guided_attention.py"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _TorchProxy:
    """Stands in for the name `torch` inside the reference pipeline module: forwards everything, but maps the two
    hard-coded CUDA call sites (`torch.Generator('cuda')` :918 and `torch.randn(..., device="cuda")` :1050) to CPU."""

    def __getattr__(self, k):
        return getattr(torch, k)

    @staticmethod
    def Generator(device="cpu"):
        return torch.Generator("cpu")

    @staticmethod
    def randn(*a, **k):
        k["device"] = "cpu"
        return torch.randn(*a, **k)

    @staticmethod
    def Tensor(*a, **k):
        return torch.Tensor(*a, **k)


def load():
    """Returns a namespace with the reference modules: state, helpers, ptp_utils, gaussian_smoothing, config,
    pipeline, run."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")

    from guided_attention_b200 import substrate as sub

    # --- placeholders --------------------------------------------------------------------------------------------
    plt = _mod("matplotlib.pyplot", imsave=lambda *a, **k: None, ioff=lambda: None, plot=lambda *a, **k: None,
               legend=lambda *a, **k: None, savefig=lambda *a, **k: None, clf=lambda: None)
    _mod("matplotlib", pyplot=plt)
    ipd = _mod("IPython.display", display=lambda *a, **k: None)
    _mod("IPython", display=ipd)
    _mod("pyrallis", wrap=lambda *a, **k: (lambda f: f))

    # transformers 5.x is installed but dropped CLIPFeatureExtractor and resolves names lazily; the reference only
    # uses these three names as annotations, so a placeholder module is enough (and much faster to import).
    _saved_transformers = sys.modules.get("transformers")
    _mod("transformers", **{n: type(n, (), {}) for n in ("CLIPFeatureExtractor", "CLIPTextModel", "CLIPTokenizer")})

    class _Logger:
        def info(self, *a, **k):
            pass
        warning = info
        debug = info

    logging_ns = types.SimpleNamespace(get_logger=lambda name=None: _Logger())
    _mod("diffusers", DDIMScheduler=sub.DDIMScheduler)
    _mod("diffusers.configuration_utils", FrozenDict=dict)
    u2c = types.SimpleNamespace(UNet2DConditionOutput=sub.UNet2DConditionOutput)
    _mod("diffusers.models", AutoencoderKL=type("AutoencoderKL", (), {}), UNet2DConditionModel=sub.UNet2DConditionModel,
         unet_2d_condition=u2c)
    _mod("diffusers.models.cross_attention", CrossAttention=sub.CrossAttention)
    _mod("diffusers.schedulers", KarrasDiffusionSchedulers=type("KarrasDiffusionSchedulers", (), {}))
    _mod("diffusers.utils", deprecate=lambda *a, **k: None, is_accelerate_available=lambda: False, logging=logging_ns,
         randn_tensor=None, replace_example_docstring=lambda *a, **k: (lambda f: f))
    _mod("diffusers.pipelines")
    _mod("diffusers.pipelines.pipeline_utils", DiffusionPipeline=type("DiffusionPipeline", (), {}))
    _mod("diffusers.pipelines.stable_diffusion", StableDiffusionPipeline=sub.StableDiffusionPipelineBase,
         StableDiffusionPipelineOutput=sub.StableDiffusionPipelineOutput)
    _mod("diffusers.pipelines.stable_diffusion.safety_checker",
         StableDiffusionSafetyChecker=type("StableDiffusionSafetyChecker", (), {}))

    # --- neutralise .cuda() (CPU runs only; this process must not be a GPU test process) ---------------------------
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self

    # The reference's package is called `utils`, `config`, `run`: import them under their own names from its root.
    clash = [n for n in ("utils", "config", "run", "pipeline_guided_attention") if n in sys.modules]
    for n in clash:
        del sys.modules[n]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        state = importlib.import_module("utils.shared_state")
        helpers = importlib.import_module("utils.helpers")
        gs = importlib.import_module("utils.gaussian_smoothing")
        ptp = importlib.import_module("utils.ptp_utils")
        cfg = importlib.import_module("config")
        pipe = importlib.import_module("pipeline_guided_attention")
        run = importlib.import_module("run")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    if _saved_transformers is not None:
        sys.modules["transformers"] = _saved_transformers
    else:
        del sys.modules["transformers"]
    pipe.torch = _TorchProxy()
    run.torch = _TorchProxy()
    _loaded = types.SimpleNamespace(state=state, helpers=helpers, gaussian_smoothing=gs, ptp_utils=ptp, config=cfg,
                                    pipeline=pipe, run=run)
    return _loaded


def make_pipeline(ref, unet, tokenizer=None):
    """Instantiate the reference's own `GuidedAttention` on the substrate UNet (CPU)."""
    from guided_attention_b200 import substrate as sub
    p = ref.pipeline.GuidedAttention(unet=unet, scheduler=sub.DDIMScheduler(),
                                     tokenizer=tokenizer or sub.WhitespaceTokenizer())
    return p


def make_config(ref, meta_prompt, output_path, stable=None, **kw):
    """A reference RunConfig wired the way `run.setup` + `run.main` leave it (run.py:139-145, 235-246)."""
    from pathlib import Path
    cfg = ref.config.RunConfig(meta_prompt=meta_prompt, output_path=Path(output_path), **kw)
    ref.state.config = cfg
    cfg.stable = stable
    ref.run.register_custom_loss("toLeftOf", ref.run.ToLeftOf())
    return cfg
