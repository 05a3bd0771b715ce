"""CoarseRestoration against PyTorch eager fp32 on the same GPU, per HD_CR_CHUNK (faces per pass)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hifidiff_b200 as H  # noqa: E402
from hifidiff_b200 import testing  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
with torch.device("meta"):
    m = H.CoarseRestoration()
s0 = m.state_dict()
sd = testing.random_state({k: v.shape for k, v in s0.items()}, {k: v.dtype for k, v in s0.items()}, seed=4)
m = m.to_empty(device="cuda")
m.load_state_dict(sd)
m.eval()
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(B, 3, 128, 128, device="cuda", generator=g)
with torch.no_grad():
    y = m(x)
    torch.cuda.synchronize()
    m.native = False
    ref = torch.cat([m(x[i:i + 16]) for i in range(0, B, 16)])
d = (y - ref).abs().flatten(1).amax(1)
print("chunk", os.environ.get("HD_CR_CHUNK", "default"), "max |native - eager| per face: max %.3e  median %.3e  worst face %d" % (float(d.max()), float(d.median()), int(d.argmax())),
      "ref range", float(ref.min()), float(ref.max()))
print(" per-32 block max:", ["%.1e" % float(d[i:i + 32].max()) for i in range(0, B, 32)])
