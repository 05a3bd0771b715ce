"""Warm per-kernel timing of one denoise step (CUDA events between plain launches)."""
import collections
import ctypes as C
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import hifidiff_b200 as H  # noqa: E402
from hifidiff_b200 import testing  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 10  # negative: each launch 16x back to back
with torch.device("meta"):
    m = H.FusedDenoiser(16)
sd0 = m.state_dict()
sd = testing.random_state({k: v.shape for k, v in sd0.items()}, {k: v.dtype for k, v in sd0.items()}, seed=2, eps_gain=0.15)
m = m.to_empty(device="cuda")
m.load_state_dict(sd)
m.eval().configure(precision="bf16", max_batch=B, max_steps=8, use_graph=False)
priors, ident = testing.synthetic_condition(B, 16, seed=0)
pc, ic = [p.cuda() for p in priors], ident.cuda()
x = torch.randn(B, 4, 16, 16).cuda()
for _ in range(2):
    m(x, 500, pc, ic)
torch.cuda.synchronize()
eng = m.engine()
cap, stride = 1024, 160
ms = (C.c_float * cap)()
labels = C.create_string_buffer(cap * stride)
n = C.c_int32()
eng.check(eng.lib.hd_profile_step(eng.handle, B, REPS, ms, labels, stride, cap, C.byref(n)), "hd_profile_step")
rows = [(labels.raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode(), ms[i] * 1e3) for i in range(n.value)]
tot = sum(v for _, v in rows)
print(f"# B={B}: {n.value} launches, {tot:.1f} us per step ({'each launch 16x in its own CUDA graph, best of ' + str(-REPS) if REPS < 0 else 'event-to-event plain launches'}, warm)")
agg = collections.defaultdict(lambda: [0, 0.0])
for lab, v in rows:
    key = re.sub(r" M=.*", "", lab)
    agg[key][0] += 1
    agg[key][1] += v
for k, (cnt, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{t:9.1f} us {100 * t / tot:5.1f}% n={cnt:3d} avg {t / cnt:7.2f}  {k}")
print("# first block, a level-3 block, a level-4 block:")
for i, (lab, v) in enumerate(rows):
    if i < 11 or 95 <= i < 105 or 150 <= i < 162 or i >= n.value - 4:
        print(f"{i:4d} {v:8.2f} us  {lab}")
