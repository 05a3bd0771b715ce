"""Runs one plain (non-graph) denoise step at batch B with HD_LV_TRACE=1 so the level-chain kernel prints its
per-phase timeline (median clocks over CTAs) to stderr."""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import hifidiff_b200 as H  # noqa: E402
from hifidiff_b200 import testing  # noqa: E402
from gpu_util import build  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m, sd = build(H.FusedDenoiser, seed=2, precision="bf16", max_batch=B, use_graph=False)
priors, ident = testing.synthetic_condition(B, 16, seed=1, device="cuda")
x = torch.randn(B, 4, 16, 16, device="cuda")
for _ in range(3):
    y = m(x, 500, priors, ident).sample
m.engine().synchronize()
print("finite", bool(torch.isfinite(y).all()))
