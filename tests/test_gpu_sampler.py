"""GPU: the x_{t-1} update kernel, Philox noise, and whole trajectories (hd_sample) against the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

import hifidiff_b200 as H
from hifidiff_b200 import _lib, schedulers as S, testing
from hifidiff_b200.sampler import _coef_array
from oracle import denoiser_ref, philox, schedulers_ref as R

from gpu_util import RawHandle, build
from util import TRAJ_EPS_GAIN, gen, golden, inputs, psnr, rel_l2, state_for

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def raw():
    h = RawHandle(max_batch=64, max_steps=1000)
    yield h
    h.close()


def _update(raw, x, eps, coefs, idx, seed=0, first=0, noise=None):
    xx = x.clone().cuda()
    arr = _coef_array([coefs[idx]])
    raw.check(raw.lib.hd_sampler_update(raw.h, xx.data_ptr(), eps.cuda().data_ptr(), arr, idx, C.c_uint64(seed),
                                        C.c_int64(first), x.shape[0], noise.data_ptr() if noise is not None else None,
                                        None), "hd_sampler_update")
    torch.cuda.synchronize()
    return xx.cpu()


@pytest.mark.parametrize("clip", [False, True])
def test_ddim_update_matches_oracle(raw, clip):
    p = S.DDIMScheduler(beta_schedule="scaled_linear", clip_sample=clip, clip_sample_range=3.0)
    o = R.DDIMSchedulerRef(clip_sample=clip, clip_sample_range=3.0)
    p.set_timesteps(50)
    o.set_timesteps(50)
    coefs = p.step_coefficients()
    x, eps = torch.randn(5, 4, 16, 16, generator=gen(1)) * 3, torch.randn(5, 4, 16, 16, generator=gen(2))
    for idx in (0, 25, 49):
        want = o.step(eps, int(o.timesteps[idx]), x)
        got = _update(raw, x, eps, coefs, idx)
        assert float((got - want).abs().max()) <= 4e-6 * float(want.abs().max())


def test_ddpm_update_explicit_and_philox_noise(raw):
    p = S.DDPMScheduler(beta_schedule="scaled_linear", clip_sample=False)
    o = R.DDPMSchedulerRef(clip_sample=False)
    p.set_timesteps(1000)
    o.set_timesteps(1000)
    coefs = p.step_coefficients()
    x, eps = torch.randn(6, 4, 16, 16, generator=gen(3)), torch.randn(6, 4, 16, 16, generator=gen(4))
    z = torch.randn(6, 4, 16, 16, generator=gen(5))
    for idx in (0, 500, 998, 999):
        t = int(o.timesteps[idx])
        want = o.step(eps, t, x, variance_noise=z)
        got = _update(raw, x, eps, coefs, idx, noise=z.cuda())
        assert float((got - want).abs().max()) <= 4e-6 * float(want.abs().max()), idx
    # Philox path: same integer stream as the numpy oracle, float transform within a few ulp
    idx, seed, first = 123, 0xDEADBEEF12345, 40
    zz = torch.from_numpy(philox.normal_noise(seed, first, 6, idx)).reshape(6, 4, 16, 16)
    want = o.step(eps, int(o.timesteps[idx]), x, variance_noise=zz)
    got = _update(raw, x, eps, coefs, idx, seed=seed, first=first)
    assert float((got - want).abs().max()) <= 1e-5
    # last step (t = 0) adds no noise
    assert torch.equal(_update(raw, x, eps, coefs, 999, seed=1), _update(raw, x, eps, coefs, 999, seed=2))


@pytest.mark.parametrize("prec,min_psnr", [("fp32", 80.0), ("bf16", 35.0)])
def test_ddim50_trajectory_vs_reference(prec, min_psnr):
    """Config 1 shape (B=1, 50 DDIM steps, eta 0): final latent against the reference's own trajectory."""
    g = golden("denoiser_ddim50.npz")
    m, sd = build(H.Denoiser, seed=1, precision=prec, eps_gain=TRAJ_EPS_GAIN, max_batch=4, max_steps=50)
    sched = H.DDIMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon",
                            clip_sample=False)
    xT = inputs("latents", 1, seed=7)
    x0 = H.ddim_sample(m, xT.cuda(), sched, 50)
    m.engine().synchronize()
    q = psnr(x0, g["x0"])
    print(f"DDIM-50 {prec}: PSNR vs reference x0 = {q:.2f} dB, rel-L2 {rel_l2(x0, g['x0']):.3e}")
    assert q >= min_psnr
    m.invalidate()


def test_graph_and_plain_launch_agree_and_sharding_is_invariant():
    m, sd = build(H.Denoiser, seed=1, precision="bf16", eps_gain=TRAJ_EPS_GAIN, max_batch=4, max_steps=20)
    sched = H.DDPMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon",
                            clip_sample=False)
    xT = inputs("latents", 4, seed=11).cuda()
    full = H.ddpm_sample(m, xT, sched, 20, seed=7, first_face=100)
    lo = H.ddpm_sample(m, xT[:2].contiguous(), sched, 20, seed=7, first_face=100)
    hi = H.ddpm_sample(m, xT[2:].contiguous(), sched, 20, seed=7, first_face=102)
    m.engine().synchronize()
    assert torch.equal(full, torch.cat([lo, hi]))            # face trajectories do not depend on the sharding
    assert not torch.equal(full, H.ddpm_sample(m, xT, sched, 20, seed=8, first_face=100))
    m.configure(use_graph=False)
    plain = H.ddpm_sample(m, xT, sched, 20, seed=7, first_face=100)
    m.engine().synchronize()
    assert torch.equal(full, plain)
    m.invalidate()


def test_ddpm_trajectory_matches_oracle_loop():
    """20-step DDPM with explicit noise: CUDA loop vs oracle model + oracle scheduler."""
    m, sd = build(H.Denoiser, seed=1, precision="fp32", eps_gain=TRAJ_EPS_GAIN, max_batch=2, max_steps=20)
    sched = H.DDPMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon",
                            clip_sample=False)
    o = R.DDPMSchedulerRef(clip_sample=False)
    xT = inputs("latents", 2, seed=12)
    z = torch.randn(20, 2, 1024, generator=gen(13))
    got = H.ddpm_sample(m, xT.cuda(), sched, 20, noise=z.cuda())
    m.engine().synchronize()
    with torch.no_grad():
        want = R.sample_loop(lambda xx, tt: denoiser_ref.denoiser_forward(sd, xx, torch.full((2,), tt, dtype=torch.long)),
                             xT, o, 20, noise_fn=lambda i, t: z[i].reshape(2, 4, 16, 16))
    assert psnr(got, want) >= 80.0
    m.invalidate()


@pytest.mark.parametrize("prec,min_psnr", [("fp32", 80.0), ("bf16", 35.0)])
def test_refiner_cascade_trajectory(prec, min_psnr):
    """BASELINE.json configs[4] scaled down: the whole cascade through the sampler call the reference makes
    (`ddim_sample(refiner, ..., cr_face, cr_latent)`, train_refiner.py:86-125 between x_T and the VAE decode) —
    IDC ResNet-50 + FPG + FusedDenoiser all native, 10 DDIM steps, 3 faces, against the oracle's cascade; then the
    same faces sharded 2 + 1 with the condition recomputed per shard."""
    from oracle import cond_ref
    m, sd = build(H.FacialRefiner, seed=3, precision=prec, eps_gain=TRAJ_EPS_GAIN, max_batch=4, max_steps=10, args=())
    sched = H.DDIMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon",
                            clip_sample=False)
    xT = inputs("latents", 3, seed=21)
    cr_face, cr_latent = inputs("cr_face", 3), inputs("cr_latent", 3)
    x0 = H.ddim_sample(m, xT.cuda(), sched, 10, cr_face=cr_face.cuda(), cr_latent=cr_latent.cuda())
    m.denoiser.engine().synchronize()
    with torch.no_grad():
        priors = cond_ref.fpg_forward(sd, cr_latent, "fpg.")
        ident = cond_ref.idc_forward(sd, cr_face, "idc.")
        want = R.sample_loop(lambda xx, tt: denoiser_ref.fused_denoiser_forward(sd, xx, tt, priors, ident, prefix="denoiser."),
                             xT, R.DDIMSchedulerRef(clip_sample=False), 10)
    q = psnr(x0, want)
    print(f"refiner cascade DDIM-10 {prec}: PSNR vs oracle x0 = {q:.2f} dB, rel-L2 {rel_l2(x0, want):.3e}")
    assert q >= min_psnr
    lo = H.ddim_sample(m, xT[:2].contiguous().cuda(), sched, 10, cr_face=cr_face[:2].contiguous().cuda(),
                       cr_latent=cr_latent[:2].contiguous().cuda())
    hi = H.ddim_sample(m, xT[2:].contiguous().cuda(), sched, 10, cr_face=cr_face[2:].contiguous().cuda(),
                       cr_latent=cr_latent[2:].contiguous().cuda(), first_face=2)
    m.denoiser.engine().synchronize()
    # split-K depth may differ with the number of row tiles, so shards agree to round-off, not always to the bit
    assert rel_l2(torch.cat([lo, hi]), x0) <= 2e-3
    m.denoiser.invalidate()


class _StubVAE:
    """Stands in for diffusers' AutoencoderKL (unavailable offline): a fixed linear 8x down / up map with the same
    call surface (`encode(x).latent_dist.sample()`, `decode(z).sample`), deterministic so both sides see the same z."""

    class _Out:
        def __init__(self, t):
            self.sample = t
            self.latent_dist = self

    def __init__(self):
        g = gen(77)
        self.enc = torch.randn(4, 3, generator=g) * 8.0    # latents of roughly unit scale after scaling_factor, like a real VAE
        self.dec = torch.randn(3, 4, generator=g) * 0.02   # decoded images stay inside (0, 1): the clamp must not decide the test

    def encode(self, x):
        z = torch.nn.functional.avg_pool2d(x, 8)
        z = torch.einsum("oc,bchw->bohw", self.enc.to(x.device), z)
        out = self._Out(z)
        out.sample = lambda: z
        return out

    def decode(self, z):
        y = torch.einsum("oc,bchw->bohw", self.dec.to(z.device), z)
        return self._Out(torch.nn.functional.interpolate(y, scale_factor=8, mode="nearest"))


@pytest.mark.parametrize("cr_tensor_cores", [True, False])
def test_pixel_pipeline_matches_oracle_pipeline(cr_tensor_cores):
    """The reference's whole `ddim_sample` (train_refiner.py:86-125) at the pixel level: native CoarseRestoration ->
    VAE encode (stub) -> native IDC + FPG + 10 DDIM steps -> VAE decode (stub), against the same pipeline assembled
    from the CPU oracles."""
    from oracle import cond_ref, cr_ref
    vae = _StubVAE()
    with torch.device("meta"):
        crm = H.CoarseRestoration()
    sd_cr = state_for(crm, seed=4)
    crm = crm.to_empty(device="cuda")
    crm.load_state_dict(sd_cr)
    crm.eval()
    crm.tensor_cores = cr_tensor_cores
    m, sd = build(H.FacialRefiner, seed=3, precision="bf16", eps_gain=TRAJ_EPS_GAIN, max_batch=4, max_steps=10, args=())
    sched = H.DDIMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon",
                            clip_sample=False)
    ln_face = inputs("ln_face", 2, seed=9)
    xT = inputs("latents", 2, seed=23)
    images = H.ddim_sample_images(ln_face.cuda(), m, vae, crm, sched, 0.18215, 10, x_T=xT.cuda())
    m.denoiser.engine().synchronize()
    with torch.no_grad():
        cr_face = cr_ref.cr_forward(sd_cr, ln_face)
        cr_latent = vae.encode(H.to_vae_range(cr_face)).latent_dist.sample() * 0.18215
        priors = cond_ref.fpg_forward(sd, cr_latent, "fpg.")
        ident = cond_ref.idc_forward(sd, cr_face, "idc.")
        x0 = R.sample_loop(lambda xx, tt: denoiser_ref.fused_denoiser_forward(sd, xx, tt, priors, ident, prefix="denoiser."),
                           xT, R.DDIMSchedulerRef(clip_sample=False), 10)
        want = H.from_vae_range(vae.decode(x0 / 0.18215).sample)
    q = psnr(images, want)
    print(f"pixel pipeline (CR tensor_cores={cr_tensor_cores} + VAE stub + refiner cascade, DDIM-10): PSNR vs oracle pipeline {q:.2f} dB")
    assert tuple(images.shape) == (2, 3, 128, 128) and float(images.min()) >= 0.0 and float(images.max()) <= 1.0
    assert q >= 35.0, q
    crm.invalidate()
    m.denoiser.invalidate()
