import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import hifidiff_b200 as H
from hifidiff_b200 import testing
from hifidiff_b200.conditioning import ResNet50, FacialPriorGuidance
from gpu_util import build
B = 256
m, sd = build(H.FusedDenoiser, seed=2, precision="bf16", max_batch=B)
eng = m.engine(B)
def rand_state(mod, seed):
    with torch.device("meta"):
        mm = mod()
    s0 = mm.state_dict()
    return {k: v.cuda() for k, v in testing.random_state({k: v.shape for k, v in s0.items()}, {k: v.dtype for k, v in s0.items()}, seed=seed).items()}
eng.load_idc_state(rand_state(ResNet50, 8))
eng.load_fpg_state(rand_state(FacialPriorGuidance, 7))
face = torch.rand((B, 3, 128, 128)).cuda()
lat = torch.randn((B, 4, 16, 16)).cuda()
for name, fn in (("idc", lambda: eng.idc_forward(face)), ("fpg", lambda: eng.fpg_forward(lat))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, "ms", e0.elapsed_time(e1) / 3)
