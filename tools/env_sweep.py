"""In-process sweep of the per-handle tuning switches (read from the environment at hd_create): ms per denoise step of
a 50-step DDIM run at B faces for each setting, plus the distance of x_0 from the default setting's x_0.

Usage: python tools/env_sweep.py [B] "HD_MAX_SPLIT=2" "HD_CTA_TARGET=100,HD_SCA_TARGET=64" ...
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import hifidiff_b200 as H  # noqa: E402
from hifidiff_b200 import testing  # noqa: E402

args = sys.argv[1:]
B = int(args.pop(0)) if args and args[0].isdigit() else 256
settings = [""] + args
STEPS = 50
with torch.device("meta"):
    m = H.FusedDenoiser(16)
sd0 = m.state_dict()
sd = testing.random_state({k: v.shape for k, v in sd0.items()}, {k: v.dtype for k, v in sd0.items()}, seed=2, eps_gain=0.15)
m = m.to_empty(device="cuda")
m.load_state_dict(sd)
m.eval()
priors, ident = testing.synthetic_condition(B, 16, seed=0)
pc, ic = [p.cuda() for p in priors], ident.cuda()
x = torch.randn(B, 4, 16, 16, generator=torch.Generator().manual_seed(0)).cuda()
sched = H.DDIMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon", clip_sample=False)
base = None
for s in settings:
    kv = [p.split("=", 1) for p in s.split(",") if p]
    for k, v in kv:
        os.environ[k] = v
    try:
        m.configure(precision="bf16", max_batch=B, max_steps=STEPS)  # new handle: reads the environment
        for _ in range(2):
            out = H.ddim_sample(m, x, sched, STEPS, facial_priors=pc, identity_embedding=ic)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = H.ddim_sample(m, x, sched, STEPS, facial_priors=pc, identity_embedding=ic)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / STEPS)
        m.engine().synchronize()
        if base is None:
            base = out.clone()
        d = float((out - base).norm() / base.norm())
        print(f"B={B} {s or 'default':40s} {best:.4f} ms/step  x0 vs default rel-L2 {d:.2e}  finite={bool(torch.isfinite(out).all())}",
              flush=True)
    finally:
        for k, _ in kv:
            os.environ.pop(k, None)
