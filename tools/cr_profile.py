"""One native CoarseRestoration pass over 32 faces for ncu (--profile-from-start off)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hifidiff_b200 as H  # noqa: E402
from hifidiff_b200 import testing  # noqa: E402

with torch.device("meta"):
    m = H.CoarseRestoration()
s0 = m.state_dict()
sd = testing.random_state({k: v.shape for k, v in s0.items()}, {k: v.dtype for k, v in s0.items()}, seed=4)
m = m.to_empty(device="cuda")
m.load_state_dict(sd)
m.eval()
x = torch.rand(int(sys.argv[1]) if len(sys.argv) > 1 else 32, 3, 128, 128, device="cuda")
with torch.no_grad():
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
