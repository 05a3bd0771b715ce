"""Device time of the native CoarseRestoration for a batch of faces, and its per-kernel breakdown (CUDA events)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hifidiff_b200 as H  # noqa: E402
from hifidiff_b200 import testing  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
with torch.device("meta"):
    m = H.CoarseRestoration()
s0 = m.state_dict()
sd = testing.random_state({k: v.shape for k, v in s0.items()}, {k: v.dtype for k, v in s0.items()}, seed=4)
m = m.to_empty(device="cuda")
m.load_state_dict(sd)
m.eval()
x = torch.rand(B, 3, 128, 128, device="cuda")
with torch.no_grad():
    for _ in range(2):
        y = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        y = m(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"native CR: {ms:.2f} ms for {B} faces ({ms / B * 1e3:.1f} us/face), finite={bool(torch.isfinite(y).all())}")
    m.native = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    xs = x[:32]
    for _ in range(2):
        m(xs)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        m(xs)
    e1.record()
    torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / 3 * (B / 32)
    print(f"PyTorch eager fp32 (TF32 off) on the same GPU: {ms_t:.2f} ms per {B} faces (measured on 32)")
