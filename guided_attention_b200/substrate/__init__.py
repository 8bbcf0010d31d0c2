"""Shape substrate: SD-shaped UNet, DDIM step, tokenizer shim, pipeline base (diffusers/CLIP are absent offline)."""
from .unet import UNetConfig, UNet2DConditionModel, UNet2DConditionOutput, CrossAttention, build_unet
from .ddim import DDIMScheduler
from .tokenizer import WhitespaceTokenizer
from .pipeline_base import StableDiffusionPipelineBase, StableDiffusionPipelineOutput
