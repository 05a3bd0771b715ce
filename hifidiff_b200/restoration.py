"""CoarseRestoration (CR) — the stage BEFORE the sampling loop (reference
models/cr/model.py:8-88, models/cr/stn.py:9-52, call site train_refiner.py:106).

This module reproduces the reference's constructor, `forward(x)` contract ((B,3,128,128) low-quality face ->
(B,3,128,128) coarse frontal face) and `state_dict()` layout exactly (checked against the unmodified reference by
tests/golden/make_golden_cr.py: same keys, order, shapes, dtypes and seeded default init), so checkpoints load
unchanged.  On a CUDA device `forward` runs on the library's own kernels (`hd_load_cr_weights` / `hd_cr_forward`:
fp32 NHWC stream, every contraction on the tensor cores with split-precision operands and fp32 accumulation —
tcgen05 3 x bf16 at c >= 128, row-scaled 3 x fp16 / 3xTF32 `mma.sync` GEMMs for the wide shallow stages, the STN
localisation conv as a scaled fp16 implicit GEMM with fused pool + ReLU — and affine-grid bilinear resampling;
DESIGN.md §1 row f.3).  There is no implicit fallback: a CPU tensor or a call that needs
autograd raises; `native = False` is the explicit opt-out to the plain PyTorch arithmetic below (training, and the
CPU tests that pin this module's layout and arithmetic against the reference).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .conditioning import _naf_block
from .modules import _Engine, _NAFBlockParams


class _STNBlock(nn.Module):
    """Spatial transformer (models/cr/stn.py:9-52): localisation CNN -> 2x3 affine -> bilinear resampling."""

    def __init__(self, in_ch: int, in_res: int):
        super().__init__()
        k = (3, 1) if in_res <= 8 else (5, 3) if in_res <= 16 else (7, 5) if in_res <= 32 else (9, 7)
        fc_res = (in_res - k[0] - 2 * k[1] + 3) // 4
        self.fc_size = 10 * fc_res * fc_res
        self.localization = nn.Sequential(nn.Conv2d(in_ch, 8, kernel_size=k[0]), nn.MaxPool2d(2, stride=2), nn.ReLU(True),
                                          nn.Conv2d(8, 10, kernel_size=k[1]), nn.MaxPool2d(2, stride=2), nn.ReLU(True))
        hidden = int(math.sqrt(self.fc_size))
        self.fc_loc = nn.Sequential(nn.Linear(self.fc_size, hidden), nn.ReLU(True), nn.Linear(hidden, 6))
        self.fc_loc[2].weight.data.zero_()
        self.fc_loc[2].bias.data.copy_(torch.tensor([1, 0, 0, 0, 1, 0], dtype=torch.float))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        theta = self.fc_loc(self.localization(x).view(-1, self.fc_size)).view(-1, 2, 3)
        grid = F.affine_grid(theta, x.size(), align_corners=False)
        return F.grid_sample(x, grid, align_corners=False)


class _NAFSTNBlock(nn.Module):
    """`NAF_STN_Block(in_channel, in_resolution, num_naf, sampling)` (models/cr/model.py:8-31)."""

    def __init__(self, c: int, res: int, num_naf: int, sampling: Optional[str] = None):
        super().__init__()
        self.nfbs = nn.Sequential(*[_NAFBlockParams(c, None) for _ in range(num_naf)])
        self.stn = _STNBlock(c, res)
        if sampling == "down":
            self.sampling = nn.Conv2d(c, 2 * c, 2, 2)
        elif sampling == "up":
            self.sampling = nn.Sequential(nn.Conv2d(c, 2 * c, 1, bias=False), nn.PixelShuffle(2))
        else:
            self.sampling = nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for blk in self.nfbs:
            x = _naf_block(blk, x)
        return self.sampling(self.stn(x))


class CoarseRestoration(nn.Module):
    """`CoarseRestoration()`; `forward(x)` (B,3,128,128) -> (B,3,128,128) (models/cr/model.py:34-88)."""

    def __init__(self):
        super().__init__()
        w = 32
        self.intro = nn.Conv2d(3, w, 3, padding=1)
        self.outro = nn.Conv2d(w, 3, 3, padding=1)
        self.encoders = nn.Sequential(_NAFSTNBlock(w, 128, 2, "down"), _NAFSTNBlock(2 * w, 64, 2, "down"),
                                      _NAFSTNBlock(4 * w, 32, 4, "down"), _NAFSTNBlock(8 * w, 16, 8, "down"))
        self.middle_blocks = _NAFSTNBlock(16 * w, 8, 8)
        self.decoders = nn.Sequential(_NAFSTNBlock(16 * w, 8, 2, "up"), _NAFSTNBlock(8 * w, 16, 2, "up"),
                                      _NAFSTNBlock(4 * w, 32, 2, "up"), _NAFSTNBlock(2 * w, 64, 2, "up"))
        self.native = True      # CUDA inputs run on the library's kernels (False: PyTorch ops)
        self.tensor_cores = True  # 1x1 / down / up convs and the STN localisation conv as split-precision (3 x bf16) tensor-core GEMMs; False: FFMA everywhere
        self._engine = None
        self._engine_dev = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    def invalidate(self) -> None:
        """Drop the library's packed copy of the weights; the next CUDA call re-reads the parameters."""
        if self._engine is not None:
            self._engine.close()
        self._engine = None
        self._engine_dev = None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate()
        return out

    def engine(self) -> _Engine:
        p = next(self.parameters())
        if p.device.type != "cuda":
            raise RuntimeError("the native CoarseRestoration path runs only on CUDA (sm_100a)")
        if self._engine is None or self._engine_dev != (p.device, bool(self.tensor_cores)):
            self.invalidate()
            with torch.cuda.device(p.device):
                torch.cuda.synchronize()
                eng = _Engine(_lib.HD_MODEL_DENOISER, 16, _lib.HD_PRECISION_BF16 if self.tensor_cores else _lib.HD_PRECISION_FP32,
                              p.device, 1, 1, False)
                eng.load_cr_state(self.state_dict())
            self._engine, self._engine_dev = eng, (p.device, bool(self.tensor_cores))
        return self._engine

    @torch.no_grad()
    def _forward_native(self, x: torch.Tensor) -> torch.Tensor:
        if tuple(x.shape[1:]) != (3, 128, 128):
            raise ValueError(f"the native CoarseRestoration takes (B,3,128,128) faces, got {tuple(x.shape)}; "
                             f"set native=False for other sizes")
        return self.engine().cr_forward(x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.native:
            # the library path: CUDA, inference (the reference calls CR under no_grad with frozen weights,
            # train_refiner.py:380-381,86).  No silent detour: anything else must opt out with native=False.
            if not x.is_cuda:
                raise RuntimeError("hifidiff_b200.CoarseRestoration runs on CUDA (sm_100a); there is no CPU path. "
                                   "Set native=False to evaluate the plain PyTorch arithmetic instead.")
            if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
                raise RuntimeError("the native CoarseRestoration is inference-only: call it under torch.no_grad() "
                                   "(or freeze the weights), or set native=False for the autograd-capable PyTorch ops")
            return self._forward_native(x)
        skips = []
        x = self.intro(x)
        for enc in self.encoders:
            x = enc(x)
            skips.append(x)
        x = self.middle_blocks(x)
        for dec, skip in zip(self.decoders, skips[::-1]):
            x = dec(x + skip)
        return self.outro(x)
