"""Functional CPU restatement of the condition-only networks (TEST INFRASTRUCTURE).

These run once per face, outside the per-timestep loop:
  FacialPriorGuidance.forward   models/fpg/model.py:46-64   (NAFBlock: models/fpg/naf.py:105-126)
  ResNet (IDC) forward          models/idc/model.py:119-135 (Bottleneck: models/idc/model.py:39-55)
  FacialRefiner.forward         models/refiner.py:32-38
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .denoiser_ref import (BN_EPS, _encoder_trunk, fused_denoiser_forward, pixel_shuffle_up)

Tensor = torch.Tensor
SD = Dict[str, Tensor]

IDC_LAYERS = (3, 4, 6, 3)   # models/idc/model.py:165


def fpg_forward(sd: SD, cr_latent: Tensor, prefix: str = "", taps: Optional[dict] = None) -> List[Tensor]:
    """Five multi-scale priors: 2048@1, 1024@2, 512@4, 256@8, 128@16 (fpg/model.py:46-64)."""
    x = F.conv2d(cr_latent, sd[prefix + "intro.weight"], sd[prefix + "intro.bias"], padding=1)
    x, skips = _encoder_trunk(sd, x, None, prefix, taps)
    x = pixel_shuffle_up(sd, prefix + "convs.0.0.weight", x, factor=1)
    priors = [x]
    for j in range(1, 5):
        x = pixel_shuffle_up(sd, f"{prefix}convs.{j}.0.weight", x, factor=2) + skips[-j]
        priors.append(x)
    return priors


def _bn(sd: SD, p: str, x: Tensor) -> Tensor:
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"],
                        sd[p + "bias"], training=False, eps=BN_EPS)


def _bottleneck(sd: SD, p: str, x: Tensor, stride: int, project: bool) -> Tensor:
    h = F.relu(_bn(sd, p + "batch_norm1.", F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"])))
    h = F.relu(_bn(sd, p + "batch_norm2.",
                   F.conv2d(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"], stride=stride, padding=1)))
    h = _bn(sd, p + "batch_norm3.", F.conv2d(h, sd[p + "conv3.weight"], sd[p + "conv3.bias"]))
    if project:
        x = _bn(sd, p + "i_downsample.1.",
                F.conv2d(x, sd[p + "i_downsample.0.weight"], sd[p + "i_downsample.0.bias"], stride=stride))
    return F.relu(h + x)


def idc_forward(sd: SD, cr_face: Tensor, prefix: str = "") -> Tensor:
    """ResNet-50 trunk without fc -> (B, 2048, 1, 1) (idc/model.py:119-135), BN in eval mode."""
    x = F.conv2d(cr_face, sd[prefix + "conv1.weight"], None, stride=2, padding=3)
    x = F.relu(_bn(sd, prefix + "batch_norm1.", x))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, nblk in enumerate(IDC_LAYERS):
        for bi in range(nblk):
            stride = 2 if (bi == 0 and li > 0) else 1
            x = _bottleneck(sd, f"{prefix}layer{li + 1}.{bi}.", x, stride, project=(bi == 0))
    x = F.adaptive_avg_pool2d(x, 1)
    return x.reshape(x.shape[0], -1, 1, 1)


def refiner_forward(sd: SD, latents: Tensor, timesteps, cr_face: Tensor, cr_latent: Tensor,
                    taps: Optional[dict] = None) -> Tensor:
    """FacialRefiner.forward (refiner.py:32-38) with keys prefixed denoiser./fpg./idc."""
    priors = fpg_forward(sd, cr_latent, "fpg.")
    ident = idc_forward(sd, cr_face, "idc.")
    if taps is not None:
        for j, p in enumerate(priors):
            taps[f"prior{j}"] = p
        taps["identity"] = ident
    return fused_denoiser_forward(sd, latents, timesteps, priors, ident, taps, prefix="denoiser.")
