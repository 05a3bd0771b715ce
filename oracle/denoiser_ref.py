"""Functional CPU restatement of the reference denoiser arithmetic (TEST INFRASTRUCTURE).

Every function takes a flat ``state_dict``-style mapping (reference key names, App. B of
SURVEY.md) plus tensors and returns tensors; nothing here is an ``nn.Module``.  All math is
fp32 NCHW exactly as the reference executes it.  ``taps`` (optional dict) collects named
intermediates for per-layer parity.

Reference anchors (paths relative to /root/reference):
  LayerNorm2d            utils.py:16-24,45-54
  SimpleGate             utils.py:57-60
  SinusoidalPosEmb       models/denoiser/model.py:17-29
  time_mlp               models/denoiser/model.py:46-51 / 152-157
  ConditionalNAFBlock    models/denoiser/conditional_naf.py:103-136
  HybridCrossAttention   models/fpg/hca.py:25-48
  Denoiser.forward       models/denoiser/model.py:106-134
  FusedDenoiser.forward  models/denoiser/model.py:217-266
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

ENC_BLOCKS = (2, 2, 4, 8)   # models/denoiser/model.py:80
MID_BLOCKS = 8              # models/denoiser/model.py:89-91
DEC_BLOCKS = (2, 2, 2, 2)   # models/denoiser/model.py:93
WIDTH = 128                 # models/denoiser/model.py:36
BN_EPS = 1e-5               # torch.nn.BatchNorm2d default, models/fpg/hca.py:14,17,22
LN_EPS = 1e-6               # utils.py:47


def layer_norm_2d(x: Tensor, weight: Tensor, bias: Tensor, eps: float = LN_EPS) -> Tensor:
    """Per-pixel LayerNorm over the channel axis, biased variance (utils.py:16-24)."""
    mu = x.mean(dim=1, keepdim=True)
    var = (x - mu).pow(2).mean(dim=1, keepdim=True)
    y = (x - mu) / (var + eps).sqrt()
    return weight.view(1, -1, 1, 1) * y + bias.view(1, -1, 1, 1)


def simple_gate(x: Tensor) -> Tensor:
    """First channel half times second channel half (utils.py:57-60)."""
    a, b = x.chunk(2, dim=1)
    return a * b


def sinusoidal_embedding(t: Tensor, dim: int = WIDTH) -> Tensor:
    """[sin(t f_i), cos(t f_i)], f_i = exp(-i ln(1e4)/(dim/2-1)) (model.py:22-29)."""
    half = dim // 2
    step = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, device=t.device) * -step)
    ang = t[:, None] * freqs[None, :]
    return torch.cat((ang.sin(), ang.cos()), dim=-1)


def canonical_timesteps(timesteps, batch: int, fused: bool) -> Tensor:
    """The timestep coercion done at the top of both forwards.

    Denoiser (model.py:107-108): python scalars / 0-d tensors are broadcast with
    ``torch.full`` (dtype follows the fill value); tensors are used as they are.
    FusedDenoiser (model.py:218-229): everything is cast to float32, a length-1 tensor is
    expanded to the batch.
    """
    if isinstance(timesteps, (int, float)) or len(timesteps.shape) == 0:
        if fused:
            return torch.full((batch,), float(timesteps), dtype=torch.float32)
        return torch.full((batch,), timesteps)
    if fused:
        timesteps = timesteps.to(dtype=torch.float32)
        if timesteps.shape[0] == 1 and batch > 1:
            timesteps = timesteps.expand(batch)
    return timesteps


def time_mlp(sd: SD, timesteps: Tensor, prefix: str = "") -> Tensor:
    """SinusoidalPosEmb -> Linear(128,1024) -> SimpleGate -> Linear(512,512)."""
    e = sinusoidal_embedding(timesteps)
    e = F.linear(e, sd[prefix + "time_mlp.1.weight"], sd[prefix + "time_mlp.1.bias"])
    e = simple_gate(e)
    return F.linear(e, sd[prefix + "time_mlp.3.weight"], sd[prefix + "time_mlp.3.bias"])


def block_modulation(sd: SD, p: str, temb: Tensor) -> Sequence[Tensor]:
    """Per-block AdaLN vectors (conditional_naf.py:18-22,103-106).

    Returns (shift_att, scale_att, shift_ffn, scale_ffn), each (B, c, 1, 1).
    """
    m = F.linear(simple_gate(temb), sd[p + "mlp.1.weight"], sd[p + "mlp.1.bias"])
    return m[:, :, None, None].chunk(4, dim=1)


def cond_naf_block(sd: SD, p: str, inp: Tensor, temb: Optional[Tensor],
                   taps: Optional[dict] = None) -> Tensor:
    """One (Conditional)NAFBlock.  ``temb=None`` gives the unconditional NAFBlock of FPG
    (models/fpg/naf.py:105-126), which is the same arithmetic without modulation."""
    c = inp.shape[1]
    x = layer_norm_2d(inp, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    if temb is not None:
        shift_att, scale_att, shift_ffn, scale_ffn = block_modulation(sd, p, temb)
        x = x * (scale_att + 1) + shift_att
    x = F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"])
    x = F.conv2d(x, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1, groups=2 * c)
    x = simple_gate(x)
    pooled = x.mean(dim=(2, 3), keepdim=True)
    x = x * F.conv2d(pooled, sd[p + "sca.1.weight"], sd[p + "sca.1.bias"])
    x = F.conv2d(x, sd[p + "conv3.weight"], sd[p + "conv3.bias"])
    y = inp + x * sd[p + "beta"]
    if taps is not None:
        taps[p + "y"] = y
    x = layer_norm_2d(y, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    if temb is not None:
        x = x * (scale_ffn + 1) + shift_ffn
    x = F.conv2d(x, sd[p + "conv4.weight"], sd[p + "conv4.bias"])
    x = simple_gate(x)
    x = F.conv2d(x, sd[p + "conv5.weight"], sd[p + "conv5.bias"])
    return y + x * sd[p + "gamma"]


def _bn_eval(sd: SD, p: str, x: Tensor) -> Tensor:
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"],
                        sd[p + "weight"], sd[p + "bias"], training=False, eps=BN_EPS)


def hca_channel_gate(sd: SD, p: str, f_g: Tensor) -> Tensor:
    """w_c = sigmoid(L2(relu(L1(avgpool + maxpool)))) -> (B, C, 1, 1) (hca.py:33-43)."""
    b = f_g.shape[0]
    pooled = F.adaptive_avg_pool2d(f_g, 1) + F.adaptive_max_pool2d(f_g, 1)
    h = F.linear(pooled.reshape(b, -1), sd[p + "channel_mlp.0.weight"], sd[p + "channel_mlp.0.bias"])
    h = F.linear(F.relu(h), sd[p + "channel_mlp.2.weight"], sd[p + "channel_mlp.2.bias"])
    return torch.sigmoid(h).reshape(b, -1, 1, 1)


def hca_spatial_gate(sd: SD, p: str, f_g: Tensor) -> Tensor:
    """w_s = sigmoid(BN(1x1(relu(BN(1x1(f_g)))))) -> (B, 1, H, W) (hca.py:12-19,45-48)."""
    h = F.conv2d(f_g, sd[p + "spatial_mlp.0.weight"], sd[p + "spatial_mlp.0.bias"])
    h = F.relu(_bn_eval(sd, p + "spatial_mlp.1.", h))
    h = F.conv2d(h, sd[p + "spatial_mlp.3.weight"], sd[p + "spatial_mlp.3.bias"])
    return torch.sigmoid(_bn_eval(sd, p + "spatial_mlp.4.", h))


def hca(sd: SD, p: str, f_g: Tensor, f_d: Tensor) -> Tensor:
    """HybridCrossAttention.forward (hca.py:25-31): gates from the prior, applied to f_d,
    then dense 3x3 + BN(eval) + ReLU."""
    w_c = hca_channel_gate(sd, p, f_g)
    w_s = hca_spatial_gate(sd, p, f_g)
    f_o = f_d + w_c * f_d + w_s * f_d
    f_o = F.conv2d(f_o, sd[p + "fused_mlp.0.weight"], sd[p + "fused_mlp.0.bias"], padding=1)
    return F.relu(_bn_eval(sd, p + "fused_mlp.1.", f_o))


def pixel_shuffle_up(sd: SD, key: str, x: Tensor, factor: int = 2) -> Tensor:
    """1x1 conv (no bias) + PixelShuffle (model.py:95-97 / fpg/model.py:34-44)."""
    return F.pixel_shuffle(F.conv2d(x, sd[key]), factor)


def _encoder_trunk(sd: SD, x: Tensor, temb: Optional[Tensor], prefix: str, taps):
    skips: List[Tensor] = []
    for lvl, nblk in enumerate(ENC_BLOCKS):
        for i in range(nblk):
            x = cond_naf_block(sd, f"{prefix}encoders.{lvl}.{i}.", x, temb, taps)
            if taps is not None:
                taps[f"{prefix}encoders.{lvl}.{i}"] = x
        skips.append(x)
        x = F.conv2d(x, sd[f"{prefix}downs.{lvl}.weight"], sd[f"{prefix}downs.{lvl}.bias"], stride=2)
        if taps is not None:
            taps[f"{prefix}downs.{lvl}"] = x
    return x, skips


def denoiser_forward(sd: SD, latents: Tensor, timesteps, taps: Optional[dict] = None,
                     prefix: str = "") -> Tensor:
    """Denoiser.forward (model.py:106-134) -> epsilon prediction (B,4,H,W)."""
    _, _, height, width = latents.shape
    t = canonical_timesteps(timesteps, latents.shape[0], fused=False)
    temb = time_mlp(sd, t, prefix)
    x = F.conv2d(latents, sd[prefix + "intro.weight"], sd[prefix + "intro.bias"], padding=1)
    if taps is not None:
        taps[prefix + "time_mlp"] = temb
        taps[prefix + "intro"] = x
    x, skips = _encoder_trunk(sd, x, temb, prefix, taps)
    for i in range(MID_BLOCKS):
        x = cond_naf_block(sd, f"{prefix}middle_blks.{i}.", x, temb, taps)
        if taps is not None:
            taps[f"{prefix}middle_blks.{i}"] = x
    for lvl, nblk in enumerate(DEC_BLOCKS):
        x = pixel_shuffle_up(sd, f"{prefix}ups.{lvl}.0.weight", x) + skips[-1 - lvl]
        if taps is not None:
            taps[f"{prefix}ups.{lvl}"] = x
        for i in range(nblk):
            x = cond_naf_block(sd, f"{prefix}decoders.{lvl}.{i}.", x, temb, taps)
            if taps is not None:
                taps[f"{prefix}decoders.{lvl}.{i}"] = x
    x = F.conv2d(x, sd[prefix + "ending.weight"], sd[prefix + "ending.bias"], padding=1)
    return x[..., :height, :width]


def fused_denoiser_forward(sd: SD, latents: Tensor, timesteps, facial_priors: Sequence[Tensor],
                           identity_embedding: Tensor, taps: Optional[dict] = None,
                           prefix: str = "") -> Tensor:
    """FusedDenoiser.forward (model.py:217-266) -> epsilon prediction (B,4,H,W)."""
    batch, _, height, width = latents.shape
    t = canonical_timesteps(timesteps, batch, fused=True)
    temb = time_mlp(sd, t, prefix)
    x = F.conv2d(latents, sd[prefix + "intro.weight"], sd[prefix + "intro.bias"], padding=1)
    if taps is not None:
        taps[prefix + "time_mlp"] = temb
        taps[prefix + "intro"] = x
    x, skips = _encoder_trunk(sd, x, temb, prefix, taps)
    for i in range(MID_BLOCKS):
        x = cond_naf_block(sd, f"{prefix}middle_blks.{i}.", x, temb, taps)
        if taps is not None:
            taps[f"{prefix}middle_blks.{i}"] = x
    idc = F.conv2d(identity_embedding, sd[prefix + "idc_conv.weight"], sd[prefix + "idc_conv.bias"])
    x = x + idc.reshape(batch, *x.shape[1:])
    x = hca(sd, prefix + "hcas.0.", facial_priors[0], x)
    if taps is not None:
        taps[prefix + "hcas.0"] = x
    for lvl, nblk in enumerate(DEC_BLOCKS):
        x = pixel_shuffle_up(sd, f"{prefix}ups.{lvl}.0.weight", x) + skips[-1 - lvl]
        if taps is not None:
            taps[f"{prefix}ups.{lvl}"] = x
        for i in range(nblk):
            x = cond_naf_block(sd, f"{prefix}decoders.{lvl}.{i}.", x, temb, taps)
            if taps is not None:
                taps[f"{prefix}decoders.{lvl}.{i}"] = x
        x = hca(sd, f"{prefix}hcas.{lvl + 1}.", facial_priors[lvl + 1], x)
        if taps is not None:
            taps[f"{prefix}hcas.{lvl + 1}"] = x
    x = F.conv2d(x, sd[prefix + "ending.weight"], sd[prefix + "ending.bias"], padding=1)
    return x[..., :height, :width]
