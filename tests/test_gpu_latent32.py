"""GPU: latent size 32 (`--image_res 256`, train_refiner.py:27): FusedDenoiser(32) one step per layer against the
oracle and the reference's stored eps (tests/golden/fused_step_s32.npz), both precisions; FPG native at latent 32; a
short DDIM trajectory against the oracle loop.  The 32x32 level runs the general (untuned) kernels."""
import numpy as np
import pytest
import torch

import hifidiff_b200 as H
from hifidiff_b200 import testing
from oracle import cond_ref, denoiser_ref, schedulers_ref as R

from gpu_util import build
from util import TRAJ_EPS_GAIN, golden, psnr, rel_l2

pytestmark = pytest.mark.gpu

TAPS = ["intro", "encoders.0.1", "downs.0", "encoders.1.1", "downs.3", "middle_blks.7", "hcas.0", "ups.0", "decoders.0.1",
        "hcas.1", "hcas.2", "decoders.3.1", "hcas.4"]


def _inputs(batch, seed):
    return torch.randn((batch, 4, 32, 32), generator=torch.Generator().manual_seed(700 + seed))


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-5), ("bf16", 1e-2)])
def test_fused_step_latent32_per_layer(prec, tol):
    g = golden("fused_step_s32.npz")
    m, sd = build(H.FusedDenoiser, seed=2, precision=prec, max_batch=2, args=(32,))
    x = _inputs(2, seed=1)
    priors, ident = testing.synthetic_condition(2, 32, seed=0)
    t = torch.from_numpy(g["t"])
    out, taps = m.forward_with_taps(x.cuda(), t.cuda(), TAPS, [p.cuda() for p in priors], ident.cuda())
    m.engine().synchronize()
    ref_taps = {}
    with torch.no_grad():
        ref = denoiser_ref.fused_denoiser_forward(sd, x, t, priors, ident, ref_taps)
    assert rel_l2(ref, g["eps"]) <= 1e-6                      # the oracle reproduces the reference's stored eps
    worst = max((rel_l2(taps[k], ref_taps[k]), k) for k in TAPS)
    for k in ("downs.3", "middle_blks.7", "hcas.0", "ups.0", "decoders.0.1", "hcas.1"):
        assert rel_l2(taps[k], g["tap_" + k.replace(".", "_")]) <= tol, k
    print(f"latent 32 {prec}: eps vs reference {rel_l2(out.sample, g['eps']):.3e}; worst tap {worst[1]} {worst[0]:.3e}")
    assert tuple(out.sample.shape) == (2, 4, 32, 32)
    assert worst[0] <= tol
    assert rel_l2(out.sample, g["eps"]) <= tol
    m.invalidate()


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("bf16", 1e-2)])
def test_fpg_native_latent32(prec, tol):
    g = golden("fused_step_s32.npz")
    m, sd = build(H.FusedDenoiser, seed=2, precision=prec, max_batch=2, args=(32,))
    with torch.device("meta"):
        fpg = H.FacialPriorGuidance()
    s0 = fpg.state_dict()
    sdf = testing.random_state({k: v.shape for k, v in s0.items()}, {k: v.dtype for k, v in s0.items()}, seed=7)
    eng = m.engine(2)
    eng.load_fpg_state({k: v.cuda() for k, v in sdf.items()})
    lat = torch.randn((2, 4, 32, 32), generator=torch.Generator().manual_seed(731))
    pri = eng.fpg_forward(lat.cuda(), 32, 128)
    eng.synchronize()
    with torch.no_grad():
        want = cond_ref.fpg_forward(sdf, lat, "")
    for j in range(3):
        assert rel_l2(want[j], g[f"fpg_prior{j}"]) <= 1e-6    # oracle == reference fixture
    errs = [rel_l2(pri[j], want[j]) for j in range(5)]
    print(f"FPG latent 32 {prec}: prior rel-L2 {['%.2e' % e for e in errs]}")
    assert [tuple(p.shape) for p in pri] == [(2, 2048, 2, 2), (2, 1024, 4, 4), (2, 512, 8, 8), (2, 256, 16, 16), (2, 128, 32, 32)]
    assert max(errs) <= tol
    m.invalidate()


def test_ddim_trajectory_latent32():
    m, sd = build(H.FusedDenoiser, seed=2, precision="bf16", eps_gain=TRAJ_EPS_GAIN, max_batch=2, max_steps=10, args=(32,))
    sched = H.DDIMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon",
                            clip_sample=False)
    priors, ident = testing.synthetic_condition(2, 32, seed=4)
    xT = _inputs(2, seed=9)
    x0 = H.ddim_sample(m, xT.cuda(), sched, 10, facial_priors=[p.cuda() for p in priors], identity_embedding=ident.cuda())
    m.engine().synchronize()
    with torch.no_grad():
        want = R.sample_loop(lambda xx, tt: denoiser_ref.fused_denoiser_forward(sd, xx, tt, priors, ident), xT,
                             R.DDIMSchedulerRef(clip_sample=False), 10)
    q = psnr(x0, want)
    print(f"latent 32 DDIM-10 bf16: PSNR vs oracle loop {q:.2f} dB")
    assert q >= 50.0
    m.invalidate()


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("bf16", 1.5e-2)])
def test_refiner_forward_latent32(prec, tol):
    """FacialRefiner(latent_res=32).forward(latents, t, cr_face 256x256, cr_latent 32x32) with IDC ResNet-50 and FPG
    native, against the oracle's refiner (models/refiner.py:32-38)."""
    m, sd = build(H.FacialRefiner, seed=3, precision=prec, max_batch=2, args=(32,))
    x = _inputs(1, seed=2)
    cr_face = torch.rand((1, 3, 256, 256), generator=torch.Generator().manual_seed(741))
    cr_latent = torch.randn((1, 4, 32, 32), generator=torch.Generator().manual_seed(742))
    with torch.no_grad():
        out = m(x.cuda(), torch.tensor([640]), cr_face.cuda(), cr_latent.cuda()).sample
        priors, ident = m.condition(cr_face.cuda(), cr_latent.cuda()) if False else m._cond
        m.denoiser.engine().synchronize()
        taps = {}
        want = cond_ref.refiner_forward(sd, x, torch.tensor([640]), cr_face, cr_latent, taps)
    e_id = rel_l2(ident, taps["identity"])
    e_pr = max(rel_l2(priors[j], taps[f"prior{j}"]) for j in range(5))
    print(f"refiner latent 32 {prec}: identity {e_id:.3e}, worst prior {e_pr:.3e}, eps {rel_l2(out, want):.3e}")
    assert e_id <= min(tol, 1e-2) and e_pr <= min(tol, 1e-2)
    assert rel_l2(out, want) <= tol
    m.denoiser.invalidate()
