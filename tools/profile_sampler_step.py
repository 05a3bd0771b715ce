"""One full sampler step (denoise + x_{t-1} update, DDPM with Philox noise) for ncu (--profile-from-start off)."""
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
sched = H.DDPMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon", clip_sample=False)
H.sample(m, x, sched, 4, facial_priors=pc, identity_embedding=ic, seed=1)   # warm-up: 4 steps
torch.cuda.synchronize()
torch.cuda.profiler.start()
H.sample(m, x, sched, 1, facial_priors=pc, identity_embedding=ic, seed=1)   # profiled: 1 step (t = 0: no noise)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
