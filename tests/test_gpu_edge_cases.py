"""GPU: ragged and edge-case inputs through the module API (row tails across 128-row tiles, partial
face groups in the dwconv / split-K paths, batch growth, empty batches, state reuse)."""
import pytest
import torch

import hifidiff_b200 as H
from hifidiff_b200 import testing
from oracle import denoiser_ref

from gpu_util import build
from util import inputs, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fused32():
    m, sd = build(H.FusedDenoiser, seed=2, precision="fp32", max_batch=4)
    yield m, sd
    m.invalidate()


@pytest.fixture(scope="module")
def fused16():
    m, sd = build(H.FusedDenoiser, seed=2, precision="bf16", max_batch=4)
    yield m, sd
    m.invalidate()


@pytest.mark.parametrize("batch", [1, 3, 130])
def test_ragged_batch_fp32(fused32, batch):
    """130 faces = one full 128-face tile + 2 at the 1x1 level, 65 tiles at 8x8, ... every tail path."""
    m, sd = fused32
    x = inputs("latents", batch, seed=20 + batch)
    priors, ident = testing.synthetic_condition(batch, 16, seed=batch)
    t = torch.arange(batch) * 7 % 1000                    # per-face timesteps
    out = m(x.cuda(), t.cuda(), [p.cuda() for p in priors], ident.cuda()).sample
    m.engine().synchronize()
    assert m.max_batch >= batch                            # the engine grew its workspace on demand
    with torch.no_grad():
        ref = denoiser_ref.fused_denoiser_forward(sd, x, t, priors, ident)
    assert rel_l2(out, ref) <= 1e-5
    # faces are independent: face 0 alone gives the same answer
    one = m(x[:1].cuda(), t[:1].cuda(), [p[:1].contiguous().cuda() for p in priors], ident[:1].contiguous().cuda()).sample
    assert rel_l2(one, ref[:1]) <= 1e-5


@pytest.mark.parametrize("batch", [1, 5, 130])
def test_ragged_batch_bf16(fused16, batch):
    m, sd = fused16
    x = inputs("latents", batch, seed=40 + batch)
    priors, ident = testing.synthetic_condition(batch, 16, seed=100 + batch)
    out = m(x.cuda(), 321, [p.cuda() for p in priors], ident.cuda()).sample
    m.engine().synchronize()
    with torch.no_grad():
        ref = denoiser_ref.fused_denoiser_forward(sd, x, 321, priors, ident)
    assert rel_l2(out, ref) <= 1e-2
    assert torch.isfinite(out).all()


def test_empty_and_malformed_inputs_raise(fused16):
    m, _ = fused16
    priors, ident = testing.synthetic_condition(2, 16, seed=0)
    pc, ic = [p.cuda() for p in priors], ident.cuda()
    with pytest.raises((RuntimeError, ValueError)):
        m(torch.zeros(0, 4, 16, 16).cuda(), 5, [p[:0] for p in pc], ic[:0])
    with pytest.raises(ValueError):
        m(torch.zeros(2, 4, 16, 16).cuda(), torch.tensor([1, 2, 3]), pc, ic)          # 3 timesteps for 2 faces
    with pytest.raises(ValueError):
        m(torch.zeros(2, 4, 16, 16).cuda(), 5, pc[:4], ic)                             # 4 priors
    with pytest.raises(ValueError):
        m(torch.zeros(2, 4, 16, 16).cuda(), 5, [pc[1]] + pc[1:], ic)                   # wrong prior shape


def test_sampler_ragged_and_repeatable(fused16):
    m, sd = fused16
    sched = H.DDIMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon",
                            clip_sample=True, clip_sample_range=3.0)              # test_refiner.py:166-171 variant
    x = inputs("latents", 3, seed=77).cuda()
    priors, ident = testing.synthetic_condition(3, 16, seed=77)
    pc, ic = [p.cuda() for p in priors], ident.cuda()
    a = H.ddim_sample(m, x, sched, 7, facial_priors=pc, identity_embedding=ic)
    b = H.ddim_sample(m, x, sched, 7, facial_priors=pc, identity_embedding=ic)
    m.engine().synchronize()
    assert torch.equal(a, b) and torch.isfinite(a).all()
    assert not torch.equal(a, x)                                                    # input untouched, output new
    c = H.ddim_sample(m, x, sched, 8, facial_priors=pc, identity_embedding=ic)      # different schedule -> new table
    assert not torch.equal(a, c)


def test_edge_convs_on_tensor_cores_and_fused_scheduler_step():
    """intro / ending 3x3 on mma.sync with split-precision operands (edge_convs.cuh) and the scheduler step fused
    behind the ending conv.  (1) The intro tap stays fp32-grade against the oracle and eps within the bf16 bar.
    (2) hd_sample's fused launch (ending conv + x_{t-1} update + step advance) is BIT-IDENTICAL to the unfused
    sequence driven from the host — module forward (same conv kernel, eps to HBM) then hd_sampler_update — step by
    step over 6 DDPM steps with explicit noise on a ragged batch.  (3) Against the CUDA-core kernels with separate launches
    (HD_EDGE_MMA=0): 2 launches fewer per step.  (Two bf16 runs whose stems differ by 1e-6 decorrelate their
    rounding noise, so eps of the two variants agree to the bf16 noise floor only, not bit for bit.)"""
    import ctypes as C
    import os
    from hifidiff_b200.sampler import _coef_array
    batch, steps, seed, first = 5, 6, 5, 17
    x = inputs("latents", batch, seed=41)
    priors, ident = testing.synthetic_condition(batch, 16, seed=41)
    cond = ([p.cuda() for p in priors], ident.cuda())
    sched = H.DDPMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon",
                            clip_sample=False)
    m, sd = build(H.FusedDenoiser, seed=2, precision="bf16", eps_gain=0.15, max_batch=8, max_steps=steps)
    out, taps = m.forward_with_taps(x.cuda(), 321, ["intro"], *cond)
    x0 = H.ddpm_sample(m, x.cuda(), sched, steps, facial_priors=cond[0], identity_embedding=cond[1], seed=seed, first_face=first)
    m.engine().synchronize()
    launches_fused = m.engine().info().launches_per_step
    # the same scheduler steps one at a time, with explicit noise, two ways: (a) hd_sample(n_steps = 1) = the fused
    # launch, (b) module forward (eps to HBM) + hd_sampler_update.  Both see a one-row time table, so every input
    # of the step is the same bit pattern.
    sched.set_timesteps(steps)
    coefs = sched.step_coefficients()
    eng = m.engine()
    z = torch.randn((steps, batch, 1024), generator=torch.Generator().manual_seed(43)).cuda()
    xa, xb = x.cuda().clone(), x.cuda().clone()
    for i, t in enumerate(sched.timesteps.tolist()):
        arr = _coef_array([coefs[i]])
        zi = z[i].contiguous()
        eng.check(eng.lib.hd_sample(eng.handle, xa.data_ptr(), arr, 1, C.c_uint64(seed), C.c_int64(first), batch, zi.data_ptr(),
                                    None), "hd_sample")
        eps = m(xb, t, *cond).sample
        eng.check(eng.lib.hd_sampler_update(eng.handle, xb.data_ptr(), eps.data_ptr(), arr, 0, C.c_uint64(seed), C.c_int64(first),
                                            batch, zi.data_ptr(), None), "hd_sampler_update")
    eng.synchronize()
    xs = xb
    x0_steps = xa
    ref_taps = {}
    with torch.no_grad():
        ref = denoiser_ref.fused_denoiser_forward(sd, x, 321, priors, ident, ref_taps)
    e_intro, e_eps = rel_l2(taps["intro"], ref_taps["intro"]), rel_l2(out.sample, ref)
    m.invalidate()
    os.environ["HD_EDGE_MMA"] = "0"
    try:
        m0, _ = build(H.FusedDenoiser, seed=2, precision="bf16", eps_gain=0.15, max_batch=8, max_steps=steps)
        out0 = m0(x.cuda(), 321, *cond).sample
        x00 = H.ddpm_sample(m0, x.cuda(), sched, steps, facial_priors=cond[0], identity_embedding=cond[1], seed=seed, first_face=first)
        m0.engine().synchronize()
        launches_plain = m0.engine().info().launches_per_step
        m0.invalidate()
    finally:
        del os.environ["HD_EDGE_MMA"]
    print(f"edge convs: intro vs oracle {e_intro:.2e}, eps vs oracle {e_eps:.2e} (CUDA-core variant {rel_l2(out0, ref):.2e}); "
          f"DDPM-{steps} fused launch == forward + hd_sampler_update: {torch.equal(x0_steps, xs)}; x0 vs CUDA-core variant {rel_l2(x0, x00):.2e}; "
          f"launches {launches_fused} vs {launches_plain}")
    assert e_intro <= 2e-5 and e_eps <= 1e-2 and rel_l2(out0, ref) <= 1e-2
    assert torch.equal(x0_steps, xs)
    assert rel_l2(x0, x00) <= 1e-2
    assert launches_plain - launches_fused == 2
