"""Functional CPU restatement of CoarseRestoration (TEST INFRASTRUCTURE).

  CoarseRestoration.forward   models/cr/model.py:75-88   (NAF_STN_Block.forward :26-31)
  NAFBlock                    models/cr/naf.py:105-126   (== models/fpg/naf.py; restated in denoiser_ref.cond_naf_block)
  STNBlock                    models/cr/stn.py:9-52

CR runs once per face BEFORE the sampling loop (train_refiner.py:106): low-quality non-frontal face
(B,3,128,128) -> coarse frontal face (B,3,128,128), which feeds the IDC network and, through the VAE, the FPG.
Pinned against the unmodified reference by tests/golden/make_golden_cr.py (bit-identical on CPU).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .denoiser_ref import cond_naf_block

Tensor = torch.Tensor
SD = Dict[str, Tensor]

WIDTH = 32                                   # models/cr/model.py:39
ENC = ((32, 128, 2), (64, 64, 2), (128, 32, 4), (256, 16, 8))   # (channels, resolution, NAF blocks)  :59-64
MID = (512, 8, 8)                            # :65
DEC = ((512, 8, 2), (256, 16, 2), (128, 32, 2), (64, 64, 2))    # :66-71


def stn_block(sd: SD, p: str, x: Tensor) -> Tensor:
    """Spatial transformer: localisation CNN -> 2x3 affine -> affine_grid + bilinear grid_sample (stn.py:43-52)."""
    xs = F.conv2d(x, sd[p + "localization.0.weight"], sd[p + "localization.0.bias"])
    xs = F.relu(F.max_pool2d(xs, 2, stride=2))
    xs = F.conv2d(xs, sd[p + "localization.3.weight"], sd[p + "localization.3.bias"])
    xs = F.relu(F.max_pool2d(xs, 2, stride=2))
    xs = xs.reshape(x.shape[0], -1)           # view(-1, fc_size): (C, H, W) order of the NCHW tensor
    h = F.relu(F.linear(xs, sd[p + "fc_loc.0.weight"], sd[p + "fc_loc.0.bias"]))
    theta = F.linear(h, sd[p + "fc_loc.2.weight"], sd[p + "fc_loc.2.bias"]).view(-1, 2, 3)
    grid = F.affine_grid(theta, list(x.shape), align_corners=False)
    return F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


def naf_stn_block(sd: SD, p: str, x: Tensor, num_naf: int, sampling: Optional[str], taps: Optional[dict] = None) -> Tensor:
    for i in range(num_naf):
        x = cond_naf_block(sd, f"{p}nfbs.{i}.", x, None)
    if taps is not None:
        taps[p + "nfbs"] = x
    x = stn_block(sd, p + "stn.", x)
    if taps is not None:
        taps[p + "stn"] = x
    if sampling == "down":
        x = F.conv2d(x, sd[p + "sampling.weight"], sd[p + "sampling.bias"], stride=2)
    elif sampling == "up":
        x = F.pixel_shuffle(F.conv2d(x, sd[p + "sampling.0.weight"]), 2)
    return x


def cr_forward(sd: SD, x: Tensor, prefix: str = "", taps: Optional[dict] = None) -> Tensor:
    """CoarseRestoration.forward (models/cr/model.py:75-88)."""
    x = F.conv2d(x, sd[prefix + "intro.weight"], sd[prefix + "intro.bias"], padding=1)
    skips = []
    for i, (_, _, n) in enumerate(ENC):
        x = naf_stn_block(sd, f"{prefix}encoders.{i}.", x, n, "down", taps)
        skips.append(x)
    x = naf_stn_block(sd, f"{prefix}middle_blocks.", x, MID[2], None, taps)
    for i, (_, _, n) in enumerate(DEC):
        x = x + skips[-1 - i]
        x = naf_stn_block(sd, f"{prefix}decoders.{i}.", x, n, "up", taps)
    return F.conv2d(x, sd[prefix + "outro.weight"], sd[prefix + "outro.bias"], padding=1)
