"""CoarseRestoration (CR) — the stage BEFORE the sampling loop (reference
models/cr/model.py:8-88, models/cr/stn.py:9-52, call site train_refiner.py:106).

Status: boundary + oracle only.  This module reproduces the reference's constructor, `forward(x)` contract
((B,3,128,128) low-quality face -> (B,3,128,128) coarse frontal face) and `state_dict()` layout exactly (checked
against the unmodified reference by tests/golden/make_golden_cr.py: same keys, order, shapes, dtypes and seeded
default init), so checkpoints load unchanged.  Its arithmetic is ordinary PyTorch ops for now — it is NOT yet on
the sm_100a kernels and is not part of any parity or throughput claim; the kernel plan is DESIGN.md §6 item 4.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .conditioning import _naf_block
from .modules import _NAFBlockParams


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

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        skips = []
        x = self.intro(x)
        for enc in self.encoders:
            x = enc(x)
            skips.append(x)
        x = self.middle_blocks(x)
        for dec, skip in zip(self.decoders, skips[::-1]):
            x = dec(x + skip)
        return self.outro(x)
