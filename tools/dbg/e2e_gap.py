import sys, time
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import hifidiff_b200 as H
from hifidiff_b200 import testing
B, T = 256, 200
dev = torch.device("cuda", 0)
with torch.device("meta"):
    model = H.FusedDenoiser(16)
sd0 = model.state_dict()
sd = testing.random_state({k: v.shape for k, v in sd0.items()}, {k: v.dtype for k, v in sd0.items()}, seed=2, eps_gain=0.15)
model = model.to_empty(device=dev); model.load_state_dict(sd); del sd
model.eval().configure(precision="bf16", max_batch=B, max_steps=T, use_graph=True)
sched = H.DDPMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon", clip_sample=False)
x_host = torch.randn((B, 4, 16, 16)).pin_memory()
priors_h, ident_h = testing.synthetic_condition(B, 16, seed=0)
priors_h = [p.pin_memory() for p in priors_h]; ident_h = ident_h.pin_memory()
x_dev = x_host.to(dev); priors_d = [p.to(dev) for p in priors_h]; ident_d = ident_h.to(dev)
out_host = torch.empty((B, 4, 16, 16)).pin_memory()
def ev_time(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, 1e3 * (time.perf_counter() - t0) / n
def h2d():
    return x_host.to(dev, non_blocking=True), [p.to(dev, non_blocking=True) for p in priors_h], ident_h.to(dev, non_blocking=True)
def setc_dev(): model.set_condition(priors_d, ident_d)
def setc_fresh():
    xd, pd, idd = h2d(); model.set_condition(pd, idd)
def samp(): H.sample(model, x_dev, sched, T, facial_priors=priors_d, identity_embedding=ident_d, seed=99)
def resident():
    model.set_condition(priors_d, ident_d); return H.sample(model, x_dev, sched, T, facial_priors=priors_d, identity_embedding=ident_d, seed=99)
def e2e():
    xd, pd, idd = h2d(); model.set_condition(pd, idd)
    x0 = H.sample(model, xd, sched, T, facial_priors=pd, identity_embedding=idd, seed=99); out_host.copy_(x0, non_blocking=True)
for name, fn in (("h2d", h2d), ("set_condition(dev)", setc_dev), ("h2d+set_condition", setc_fresh), ("sample", samp), ("resident", resident), ("e2e", e2e)):
    print(name, "event ms %.2f wall ms %.2f" % ev_time(fn))
