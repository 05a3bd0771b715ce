"""Phase timeline (SM clocks) of the fused per-face block kernel: HD_FACE_TRACE=1 python tools/face_trace.py"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
if "HD_PAIR_TRACE" not in os.environ and "HD_QUAD_TRACE" not in os.environ:
    os.environ.setdefault("HD_FACE_TRACE", "0")
import torch

import hifidiff_b200 as H
from hifidiff_b200 import testing
from gpu_util import build
from util import inputs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m, sd = build(H.FusedDenoiser, seed=2, precision="bf16", max_batch=B, use_graph=False)
x = inputs("latents", B, seed=1).cuda()
priors, ident = testing.synthetic_condition(B, 16, seed=1)
cond = ([p.cuda() for p in priors], ident.cuda())
for _ in range(3):
    m(x, 500, *cond)
m.engine().synchronize()
