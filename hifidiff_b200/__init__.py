"""hifidiff_b200 — B200-native (sm_100a) implementation of HifiDiff's reverse-sampling hot path.

Drop-in for the reference's module API (`Denoiser`, `FusedDenoiser`, `FacialRefiner`,
`UNet2DOutput`, the `ddim_sample` loop); all per-timestep work runs in hand-written CUDA behind
the C ABI in `include/hifidiff_b200.h`.  See DESIGN.md.
"""
from .modules import Denoiser, FusedDenoiser, UNet2DOutput
from .conditioning import FacialPriorGuidance, FacialRefiner, ResNet50
from .restoration import CoarseRestoration
from .schedulers import DDIMScheduler, DDPMScheduler
from .sampler import ddim_sample, ddpm_sample, sample, sample_sharded, shard_bounds
from .pipeline import ddim_sample_images, encode_latent, from_vae_range, initial_noise, to_vae_range

__all__ = ["Denoiser", "FusedDenoiser", "UNet2DOutput", "FacialPriorGuidance", "FacialRefiner", "ResNet50", "CoarseRestoration",
           "DDIMScheduler", "DDPMScheduler", "ddim_sample", "ddpm_sample", "sample", "sample_sharded",
           "shard_bounds", "ddim_sample_images", "encode_latent", "initial_noise", "from_vae_range", "to_vae_range"]
