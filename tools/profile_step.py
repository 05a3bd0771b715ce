"""One FusedDenoiser denoise step at batch B for ncu (`--profile-from-start off`): the profiled
region is a single plain-launch (non-graph) step after warm-up."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import hifidiff_b200 as H  # noqa: E402
from hifidiff_b200 import testing  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
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
for _ in range(3):
    m(x, 500, pc, ic)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    m(x, 500, pc, ic)
e1.record()
torch.cuda.synchronize()
print(f"plain-launch step at B={B}: {e0.elapsed_time(e1) / 5:.3f} ms")
torch.cuda.profiler.start()
m(x, 500, pc, ic)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
