import ctypes as C
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import hifidiff_b200 as H
from hifidiff_b200 import testing, _lib
from hifidiff_b200.schedulers import StepCoef
from hifidiff_b200.sampler import _coef_array
from gpu_util import build
from util import inputs, rel_l2

batch = 5
x = inputs("latents", batch, seed=41)
priors, ident = testing.synthetic_condition(batch, 16, seed=41)
cond = ([p.cuda() for p in priors], ident.cuda())
m, sd = build(H.FusedDenoiser, seed=2, precision="bf16", eps_gain=0.15, max_batch=8, max_steps=6)
eng = m.engine(batch)
m._ensure_condition(cond[0], cond[1], batch)
t = 830.0
for trial in range(2):
    eps_mod = m(x.cuda(), t, *cond).sample.clone()
    eps_mod2 = m(x.cuda(), t, *cond).sample.clone()
    xa = x.cuda().clone()
    arr = _coef_array([StepCoef(t, 0.0, 1.0, 0.0, 0.0, 1.0, 0.0, 0.0)])
    eng.check(eng.lib.hd_sample(eng.handle, xa.data_ptr(), arr, 1, C.c_uint64(1), C.c_int64(0), batch, None, None), "hd_sample")
    eng.synchronize()
    print("trial", trial, "module twice equal:", torch.equal(eps_mod, eps_mod2), "fused eps == module eps:", torch.equal(xa, eps_mod),
          "rel", rel_l2(xa, eps_mod))
m.configure(use_graph=False)
eng = m.engine(batch)
m._ensure_condition(cond[0], cond[1], batch)
eps_mod = m(x.cuda(), t, *cond).sample.clone()
xa = x.cuda().clone()
eng.check(eng.lib.hd_sample(eng.handle, xa.data_ptr(), arr, 1, C.c_uint64(1), C.c_int64(0), batch, None, None), "hd_sample")
eng.synchronize()
print("no graph: fused eps == module eps:", torch.equal(xa, eps_mod), rel_l2(xa, eps_mod))
