"""Restatement of the pixel-level glue around the sampling loop (TEST INFRA — never imported by the product).

  to_vae_range    train_refiner.py:56-61    x.clamp(0, 1) * 2.0 - 1.0
  from_vae_range  train_refiner.py:64-69    ((x + 1.0) / 2.0).clamp(0, 1)
  encode_latent   train_refiner.py:72-83    bicubic resize -> to_vae_range -> vae.encode(...).latent_dist.sample() * sf

Pinned by tests/test_pipeline_host.py::test_pipeline_ref_matches_reference_source, which (where /root/reference is
present) executes the reference's own function bodies, extracted from train_refiner.py with `ast`, on the same
inputs.  The VAE itself (SD-2.1 AutoencoderKL) is external to the reference tree and unavailable offline.
"""
import torch
import torch.nn.functional as F


def to_vae_range(x: torch.Tensor) -> torch.Tensor:
    return x.clamp(0, 1) * 2.0 - 1.0


def from_vae_range(x: torch.Tensor) -> torch.Tensor:
    return ((x + 1.0) / 2.0).clamp(0, 1)


def encode_latent(vae, images: torch.Tensor, scaling_factor: float, image_res: int = 128) -> torch.Tensor:
    images = F.interpolate(images, size=(image_res, image_res), mode="bicubic", align_corners=False)
    images = to_vae_range(images)
    latents = vae.encode(images).latent_dist.sample()
    return latents * scaling_factor
