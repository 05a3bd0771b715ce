"""Condition-only networks and the refiner composition (run once per face, outside the loop).

  FacialPriorGuidance  <- models/fpg/model.py:7-64 (NAFBlock: models/fpg/naf.py:23-126)
  ResNet50 (IDC)       <- models/idc/model.py:10-55,102-166
  FacialRefiner        <- models/refiner.py:10-38

These are SURVEY.md §8(f) "next" rows 1-2: they depend on neither x_t nor t and cost ~5 GFLOP per face
against >= 104 GFLOP for the sampling loop.  The nn.Modules below keep the reference's exact
`state_dict()` layout (and a PyTorch eager `forward` used by the parity tests); on a CUDA device
`FacialRefiner.condition` runs both networks on the sm_100a kernels instead (`hd_fpg_forward`,
`hd_idc_forward`; switches `native_fpg` / `native_idc`), once per (cr_face, cr_latent) pair, and hands
the priors and the identity embedding to the sm_100a `FusedDenoiser` — instead of recomputing them at
every step as refiner.py:33-34 does (same result: they are t-invariant).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn.functional as F
from torch import nn

from .modules import FusedDenoiser, UNet2DOutput, _LayerNorm2dParams, _NAFBlockParams, _NoParams


def _layer_norm_2d(x: torch.Tensor, p: _LayerNorm2dParams, eps: float = 1e-6) -> torch.Tensor:
    mu = x.mean(1, keepdim=True)
    var = (x - mu).pow(2).mean(1, keepdim=True)
    return p.weight.view(1, -1, 1, 1) * ((x - mu) / (var + eps).sqrt()) + p.bias.view(1, -1, 1, 1)


def _gate(x: torch.Tensor) -> torch.Tensor:
    a, b = x.chunk(2, dim=1)
    return a * b


def _naf_block(p: _NAFBlockParams, inp: torch.Tensor) -> torch.Tensor:
    """Unconditional NAFBlock (models/fpg/naf.py:105-126)."""
    x = p.conv1(_layer_norm_2d(inp, p.norm1))
    x = _gate(p.conv2(x))
    x = x * p.sca[1](x.mean(dim=(2, 3), keepdim=True))
    y = inp + p.conv3(x) * p.beta
    x = p.conv5(_gate(p.conv4(_layer_norm_2d(y, p.norm2))))
    return y + x * p.gamma


class FacialPriorGuidance(nn.Module):
    """NAFNet encoder over the CR latent; returns the 5 priors 2048@1, 1024@2, 512@4, 256@8, 128@16."""

    def __init__(self):
        super().__init__()
        width = 32 * 4
        self.intro = nn.Conv2d(4, width, 3, padding=1)
        self.encoders = nn.ModuleList()
        self.downs = nn.ModuleList()
        self.convs = nn.ModuleList()
        chan = width
        for num in (2, 2, 4, 8):
            self.encoders.append(nn.Sequential(*[_NAFBlockParams(chan, None) for _ in range(num)]))
            self.downs.append(nn.Conv2d(chan, 2 * chan, 2, 2))
            chan *= 2
        self.convs.append(nn.Sequential(nn.Conv2d(chan, chan, 1, bias=False), _NoParams()))
        for _ in range(4):
            self.convs.append(nn.Sequential(nn.Conv2d(chan, chan * 2, 1, bias=False), _NoParams()))
            chan //= 2

    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        skips = []
        x = self.intro(x)
        for blocks, down in zip(self.encoders, self.downs):
            for blk in blocks:
                x = _naf_block(blk, x)
            skips.append(x)
            x = down(x)
        x = self.convs[0][0](x)  # PixelShuffle(1) is the identity
        priors = [x]
        for conv, skip in zip(list(self.convs)[1:], skips[::-1]):
            x = F.pixel_shuffle(conv[0](x), 2) + skip
            priors.append(x)
        return priors


class _Bottleneck(nn.Module):
    def __init__(self, cin: int, planes: int, stride: int, projection: Optional[nn.Module]):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 1)
        self.batch_norm1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1)
        self.batch_norm2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1)
        self.batch_norm3 = nn.BatchNorm2d(planes * 4)
        self.i_downsample = projection

    def forward(self, x):
        h = F.relu(self.batch_norm1(self.conv1(x)))
        h = F.relu(self.batch_norm2(self.conv2(h)))
        h = self.batch_norm3(self.conv3(h))
        if self.i_downsample is not None:
            x = self.i_downsample(x)
        return F.relu(h + x)


class ResNet(nn.Module):
    """ResNet-50 trunk without the fc head -> (B, 2048, 1, 1) identity embedding."""

    def __init__(self, layers=(3, 4, 6, 3), num_channels: int = 3):
        super().__init__()
        self.conv1 = nn.Conv2d(num_channels, 64, 7, stride=2, padding=3, bias=False)
        self.batch_norm1 = nn.BatchNorm2d(64)
        cin = 64
        for li, (n, planes) in enumerate(zip(layers, (64, 128, 256, 512))):
            stride = 1 if li == 0 else 2
            # the projection is created before the block (same parameter-creation order as idc/model.py:137-153)
            proj = nn.Sequential(nn.Conv2d(cin, planes * 4, 1, stride=stride), nn.BatchNorm2d(planes * 4))
            blocks = [_Bottleneck(cin, planes, stride, proj)]
            cin = planes * 4
            blocks += [_Bottleneck(cin, planes, 1, None) for _ in range(n - 1)]
            setattr(self, f"layer{li + 1}", nn.Sequential(*blocks))

    def forward(self, x):
        x = F.relu(self.batch_norm1(self.conv1(x)))
        x = F.max_pool2d(x, 3, 2, 1)
        for li in range(4):
            x = getattr(self, f"layer{li + 1}")(x)
        return F.adaptive_avg_pool2d(x, 1).reshape(x.shape[0], -1, 1, 1)


def ResNet50(channels: int = 3) -> ResNet:
    return ResNet((3, 4, 6, 3), channels)


def _bump_epoch(module, incompatible) -> None:
    module._hd_epoch = getattr(module, "_hd_epoch", 0) + 1


def _tensor_key(tensors):
    """Strong references plus the version counters at the time the condition was computed."""
    return tuple((t, t._version) for t in tensors)


def _same_tensors(key, tensors) -> bool:
    return (key is not None and len(key) == len(tensors)
            and all(k[0] is t and k[1] == t._version for k, t in zip(key, tensors)))


class FacialRefiner(nn.Module):
    """`FacialRefiner(latent_res=16, idc_ckpt=None, denoiser_ckpt=None)`;
    `forward(latents, timesteps, cr_face, cr_latent)` -> UNet2DOutput (refiner.py:32-38)."""

    def __init__(self, latent_res=16, idc_ckpt=None, denoiser_ckpt=None):
        super().__init__()
        self.idc = ResNet50()
        self.denoiser = FusedDenoiser(latent_res)
        self.fpg = FacialPriorGuidance()
        if idc_ckpt is not None:
            self.idc.load_state_dict(torch.load(idc_ckpt)["model_state_dict"])
        self.idc.eval()
        if denoiser_ckpt is not None:
            from safetensors.torch import load_file
            weights = load_file(denoiser_ckpt)
            self.denoiser.load_state_dict(weights, strict=False)
            self.fpg.load_state_dict(weights, strict=False)
            for name, param in self.denoiser.named_parameters():
                if name.startswith("intro") or name.startswith("encoders"):
                    param.requires_grad = False
        self._cond_src = None
        self._cond: Optional[tuple] = None
        self.native_fpg = True   # run FPG on the sm_100a kernels (False: PyTorch eager)
        self.native_idc = True   # run the IDC ResNet-50 on the sm_100a kernels (False: PyTorch/cuDNN eager)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._drop_condition())
        # the engine keeps packed copies of the FPG / IDC weights: a load aimed directly at a sub-module
        # (refiner.fpg.load_state_dict(...), refiner.idc.load_state_dict(...)) must drop them as well.  The hooks only
        # bump a counter on the sub-module itself (no reference to the parent: deepcopy-safe); `condition` compares.
        for sub in (self.fpg, self.idc):
            sub._hd_epoch = 0
            sub.register_load_state_dict_post_hook(_bump_epoch)
        self._packed_epoch = (0, 0)

    def _drop_condition(self) -> None:
        self._cond_src = None
        self._cond = None

    def _weights_changed(self) -> None:
        """FPG / IDC parameters were replaced: forget the cached condition and the engine's packed copies."""
        self._drop_condition()
        self.denoiser.invalidate()

    def _sync_weight_epoch(self) -> None:
        epoch = (self.fpg._hd_epoch, self.idc._hd_epoch)
        if epoch != self._packed_epoch:
            self._weights_changed()
            self._packed_epoch = epoch

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)  # .to() / .cuda() / .float(): every cached tensor is stale
        self._weights_changed()
        return out

    @torch.no_grad()
    def condition(self, cr_face: torch.Tensor, cr_latent: torch.Tensor):
        """(priors, identity) for a batch of faces; cached while the SAME tensor objects are passed again, unmodified.

        The cache holds strong references to `cr_face` / `cr_latent` and compares object identity and `_version`:
        an address can never be handed to a fresh tensor while the cached one is alive, so a second batch of faces
        (new tensors, possibly at a recycled address) always recomputes.  The sm_100a kernels never write into
        tensors they were handed as inputs, so `_version` covers every in-place change PyTorch can make."""
        if cr_face is None or cr_latent is None:
            raise ValueError("FacialRefiner needs cr_face and cr_latent")
        if (self.native_fpg and cr_latent.device.type != "cuda") or (self.native_idc and cr_face.device.type != "cuda"):
            # no silent detour through PyTorch eager: the native paths are the product; native_* = False is the opt-out
            raise RuntimeError("hifidiff_b200 has no CPU path: cr_face and cr_latent must be CUDA tensors")
        self._sync_weight_epoch()
        if not _same_tensors(self._cond_src, (cr_face, cr_latent)):
            was_training = self.training
            self.eval()  # BatchNorm must use running statistics on the sampling path
            # full fp32 convolutions: the identity feeds the fp32 correctness mode too, and this runs once per face
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                if self.native_fpg:
                    eng = self.denoiser.engine(cr_latent.shape[0])
                    if not eng.fpg_loaded:
                        eng.load_fpg_state(self.fpg.state_dict())
                    priors = eng.fpg_forward(cr_latent, self.denoiser.config.sample_size, self.denoiser.width)
                else:
                    priors = self.fpg(cr_latent)
                if self.native_idc:
                    want = 8 * self.denoiser.config.sample_size
                    if cr_face.shape[-1] != want or cr_face.shape[-2] != want:
                        # no silent detour through PyTorch: the native ResNet-50 is built for the pipeline's face size
                        raise ValueError(f"cr_face must be (B,3,{want},{want}) for latent size "
                                         f"{self.denoiser.config.sample_size}; set native_idc=False to run the "
                                         f"PyTorch module on other sizes")
                    eng = self.denoiser.engine(cr_face.shape[0])
                    if not eng.idc_loaded:
                        eng.load_idc_state(self.idc.state_dict())
                    ident = eng.idc_forward(cr_face)
                else:
                    ident = self.idc(cr_face)
            self.train(was_training)
            self._cond = ([p.contiguous() for p in priors], ident.contiguous())
            self._cond_src = _tensor_key((cr_face, cr_latent))
        return self._cond

    def forward(self, latents, timesteps, cr_face, cr_latent):
        priors, ident = self.condition(cr_face, cr_latent)
        return self.denoiser(latents, timesteps, priors, ident)
