"""Pixel-to-pixel mirror of the reference's `ddim_sample` (train_refiner.py:86-125; copies at test_refiner.py:58-95):

    cr_face   = cr_module(ln_face)                                   # CoarseRestoration, native (hd_cr_forward)
    cr_latent = encode_latent(vae, cr_face, scaling_factor)          # the caller's VAE (diffusers AutoencoderKL)
    latents   = reverse sampling with FacialRefiner(latents, t, cr_face, cr_latent)   # native: IDC + FPG + hd_sample
    images    = from_vae_range(vae.decode(latents / scaling_factor).sample)

Everything the reference computes with its own modules runs on this library's kernels; the VAE is the one external
object (SD-2.1 `AutoencoderKL`, not part of the reference tree and not available offline — SURVEY.md §8f row 4): it is
passed in and only needs `.encode(x).latent_dist.sample()` and `.decode(z).sample`, as diffusers' class provides.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from .conditioning import FacialRefiner
from .sampler import sample
from .schedulers import _SchedulerBase


def to_vae_range(x: torch.Tensor) -> torch.Tensor:
    """[0, 1] -> [-1, 1], clamped first (train_refiner.py:56-61: `x.clamp(0, 1) * 2.0 - 1.0`)."""
    return x.clamp(0, 1) * 2.0 - 1.0


def from_vae_range(x: torch.Tensor) -> torch.Tensor:
    """[-1, 1] -> [0, 1], clamped (train_refiner.py:64-69)."""
    return ((x + 1.0) / 2.0).clamp(0, 1)


@torch.no_grad()
def encode_latent(vae, images: torch.Tensor, scaling_factor: float, image_res: int = 128) -> torch.Tensor:
    """train_refiner.py:72-83.  The bicubic resize is skipped when the size already matches: at scale 1 with
    align_corners=False every output pixel sits on an input pixel and the cubic weights are exactly (0, 1, 0, 0)."""
    if images.shape[-1] != image_res or images.shape[-2] != image_res:
        images = F.interpolate(images, size=(image_res, image_res), mode="bicubic", align_corners=False)
    return vae.encode(to_vae_range(images)).latent_dist.sample() * scaling_factor


def initial_noise(n_faces: int, latent_res: int = 16, seed: int = 0, first_face: int = 0) -> torch.Tensor:
    """x_T for faces [first_face, first_face + n_faces): face i is drawn from its own generator keyed by
    (seed, global face index), so any sharding of a set of faces over ranks starts every face from the same noise
    (the reference draws one batch tensor from the global RNG, train_refiner.py:101-104)."""
    out = torch.empty((n_faces, 4, latent_res, latent_res), dtype=torch.float32)
    for i in range(n_faces):
        g = torch.Generator(device="cpu").manual_seed((int(seed) * 0x9E3779B1 + first_face + i) & 0x7FFFFFFFFFFFFFFF)
        out[i] = torch.randn((4, latent_res, latent_res), generator=g)
    return out


@torch.no_grad()
def ddim_sample_images(ln_face: torch.Tensor, unet: FacialRefiner, vae, cr_module, scheduler: _SchedulerBase,
                       scaling_factor: float = 0.18215, num_inference_steps: int = 50, *, image_res: int = 128,
                       x_T: Optional[torch.Tensor] = None, seed: int = 0, first_face: int = 0) -> torch.Tensor:
    """The reference's `ddim_sample(ln_face, unet, vae, cr_module, scheduler, accelerator, scaling_factor,
    num_inference_steps)` (train_refiner.py:86-125) without the accelerator argument.  `x_T` replaces the draw from
    the global RNG (:101-104) so that runs are reproducible; when omitted it is drawn per global face index
    (`initial_noise`), so shards of one set of faces (same seed, their own first_face) see the same noise."""
    if ln_face.device.type != "cuda":
        raise RuntimeError("hifidiff_b200 has no CPU path: ln_face must be a CUDA tensor")
    bs, latent_res = ln_face.shape[0], image_res // 8
    if x_T is None:
        x_T = initial_noise(bs, latent_res, seed, first_face).to(ln_face.device)
    cr_face = cr_module(ln_face)
    cr_latent = encode_latent(vae, cr_face, scaling_factor, image_res).to(torch.float32).contiguous()
    latents = sample(unet, x_T, scheduler, num_inference_steps, cr_face=cr_face, cr_latent=cr_latent, seed=seed,
                     first_face=first_face)
    return from_vae_range(vae.decode(latents / scaling_factor).sample)
