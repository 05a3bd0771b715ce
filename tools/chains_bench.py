"""Throughput of C concurrent independent sampling chains (C engines, B/C faces each)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hifidiff_b200 as H  # noqa: E402
from hifidiff_b200 import testing  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
with torch.device("meta"):
    proto = H.FusedDenoiser(16)
sd0 = proto.state_dict()
sd = testing.random_state({k: v.shape for k, v in sd0.items()}, {k: v.dtype for k, v in sd0.items()}, seed=2, eps_gain=0.15)
sched = H.DDPMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon", clip_sample=False)
for C in (1, 2, 4):
    b = B // C
    models, xs, conds = [], [], []
    for i in range(C):
        with torch.device("meta"):
            m = H.FusedDenoiser(16)
        m = m.to_empty(device="cuda")
        m.load_state_dict(sd)
        m.eval().configure(precision="bf16", max_batch=b, max_steps=T, use_graph=True)
        pri, idt = testing.synthetic_condition(b, 16, seed=i)
        pri, idt = [p.cuda() for p in pri], idt.cuda()
        m.set_condition(pri, idt)
        models.append(m)
        conds.append((pri, idt))
        xs.append(torch.randn(b, 4, 16, 16).cuda())

    streams = [torch.cuda.Stream() for _ in range(C)]

    def run():
        outs = []
        for m, x, (pri, idt), st in zip(models, xs, conds, streams):
            with torch.cuda.stream(st):  # one user stream per chain: calls on one stream are ordered by design
                outs.append(H.sample(m, x, sched, T, facial_priors=pri, identity_embedding=idt, seed=1, first_face=0))
        return outs

    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"chains={C} faces/chain={b}: {dt * 1e3 / T:.3f} ms per denoise step of {B} faces -> {B / dt / (1000 / T):.1f} faces/s @1000 steps")
    for m in models:
        m.invalidate()
    del models, xs, conds
    torch.cuda.empty_cache()
