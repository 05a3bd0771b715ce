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


@pytest.mark.parametrize("prec,min_psnr", [("fp32", 134.0), ("bf16", 59.0)])   # measured on B200: 140.9 / 65.2 dB
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
    q = psnr(got, want)
    print(f"DDPM-20 fp32 explicit noise: PSNR vs oracle loop = {q:.2f} dB")
    assert q >= 120.0
    m.invalidate()


def _fused_sched():
    return H.DDIMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon",
                           clip_sample=False)


@pytest.mark.parametrize("prec,min_psnr", [("fp32", 133.0), ("bf16", 55.0)])   # measured on B200: 139.6 / 61.6 dB
def test_fused_ddim50_trajectory_vs_reference(prec, min_psnr):
    """FusedDenoiser, 50 DDIM steps, synthetic priors / identity: final latent against the trajectory the
    reference's own module produced (tests/golden/fused_ddim50.npz, written by make_golden.py section 5)."""
    g = golden("fused_ddim50.npz")
    m, sd = build(H.FusedDenoiser, seed=2, precision=prec, eps_gain=TRAJ_EPS_GAIN, max_batch=4, max_steps=50)
    priors, ident = testing.synthetic_condition(1, 16, seed=3, device="cuda")
    xT = inputs("latents", 1, seed=8)
    x0 = H.ddim_sample(m, xT.cuda(), _fused_sched(), 50, facial_priors=priors, identity_embedding=ident)
    m.engine().synchronize()
    q = psnr(x0, g["x0"])
    print(f"FusedDenoiser DDIM-50 {prec}: PSNR vs reference x0 = {q:.2f} dB, rel-L2 {rel_l2(x0, g['x0']):.3e}")
    assert q >= min_psnr
    m.invalidate()


def test_ddpm1000_bf16_trajectory_vs_oracle():
    """BASELINE.json configs[2] scaled to 2 faces: the full 1000-step DDPM ancestral loop of the conditional
    denoiser in bf16 (the headline precision) with explicit noise, against the fp32 CPU oracle loop (model oracle
    pinned to the reference; scheduler oracle restated from diffusers 0.32.2).  This is the bf16-drift-over-1000-steps
    check SURVEY.md section 7 asks for."""
    import time
    m, sd = build(H.FusedDenoiser, seed=2, precision="bf16", eps_gain=TRAJ_EPS_GAIN, max_batch=2, max_steps=1000)
    sched = H.DDPMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon",
                            clip_sample=False)
    priors, ident = testing.synthetic_condition(2, 16, seed=5)
    xT = inputs("latents", 2, seed=31)
    z = torch.randn(1000, 2, 1024, generator=gen(32))
    got = H.ddpm_sample(m, xT.cuda(), sched, 1000, noise=z.cuda(), facial_priors=[p.cuda() for p in priors],
                        identity_embedding=ident.cuda())
    m.engine().synchronize()
    t0 = time.time()
    with torch.no_grad():
        want = R.sample_loop(lambda xx, tt: denoiser_ref.fused_denoiser_forward(sd, xx, torch.full((2,), tt, dtype=torch.long),
                                                                                priors, ident),
                             xT, R.DDPMSchedulerRef(clip_sample=False), 1000,
                             noise_fn=lambda i, t: z[i].reshape(2, 4, 16, 16))
    q = psnr(got, want)
    print(f"DDPM-1000 bf16, 2 faces: PSNR vs fp32 oracle loop = {q:.2f} dB, rel-L2 {rel_l2(got, want):.3e} "
          f"(oracle loop {time.time() - t0:.0f} s on {torch.get_num_threads()} threads)")
    assert torch.isfinite(got).all()
    assert q >= 54.0                        # measured on B200: 60.0 dB
    m.invalidate()


@pytest.mark.parametrize("prec,min_psnr", [("fp32", 129.0), ("bf16", 56.0)])   # measured on B200: 135.8 / 62.6 dB
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
    from oracle import cond_ref, cr_ref, pipeline_ref
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
        cr_latent = pipeline_ref.encode_latent(vae, cr_face, 0.18215, 128)
        priors = cond_ref.fpg_forward(sd, cr_latent, "fpg.")
        ident = cond_ref.idc_forward(sd, cr_face, "idc.")
        x0 = R.sample_loop(lambda xx, tt: denoiser_ref.fused_denoiser_forward(sd, xx, tt, priors, ident, prefix="denoiser."),
                           xT, R.DDIMSchedulerRef(clip_sample=False), 10)
        want = pipeline_ref.from_vae_range(vae.decode(x0 / 0.18215).sample)
    q = psnr(images, want)
    print(f"pixel pipeline (CR tensor_cores={cr_tensor_cores} + VAE stub + refiner cascade, DDIM-10): PSNR vs oracle pipeline {q:.2f} dB")
    assert tuple(images.shape) == (2, 3, 128, 128) and float(images.min()) >= 0.0 and float(images.max()) <= 1.0
    assert q >= 35.0, q
    crm.invalidate()
    m.denoiser.invalidate()


def _oracle_pipeline(sd_cr, sd, vae, ln_face, xT, steps):
    from oracle import cond_ref, cr_ref, pipeline_ref
    with torch.no_grad():
        cr_face = cr_ref.cr_forward(sd_cr, ln_face)
        cr_latent = pipeline_ref.encode_latent(vae, cr_face, 0.18215, 128)
        priors = cond_ref.fpg_forward(sd, cr_latent, "fpg.")
        ident = cond_ref.idc_forward(sd, cr_face, "idc.")
        x0 = R.sample_loop(lambda xx, tt: denoiser_ref.fused_denoiser_forward(sd, xx, tt, priors, ident, prefix="denoiser."),
                           xT, R.DDIMSchedulerRef(clip_sample=False), steps)
        return pipeline_ref.from_vae_range(vae.decode(x0 / 0.18215).sample), cr_face


def test_two_successive_batches_each_get_their_own_condition():
    """The reference's val / test loops call ddim_sample once per batch (train_refiner.py:213-299).  Two batches of
    equal shape and different faces, back to back through the same modules: each must be refined with ITS identity
    and priors.  (Round 1 keyed the condition cache on (data_ptr, _version, shape); the caching allocator hands the
    second batch's cr_face / cr_latent the first batch's addresses, so batch 2 was sampled with batch 1's
    condition.)  Each result is compared with the oracle pipeline; the out-of-range CR output also exercises the
    clamp of to_vae_range (train_refiner.py:56-61)."""
    vae = _StubVAE()
    with torch.device("meta"):
        crm = H.CoarseRestoration()
    sd_cr = state_for(crm, seed=4)
    crm = crm.to_empty(device="cuda")
    crm.load_state_dict(sd_cr)
    crm.eval()
    m, sd = build(H.FacialRefiner, seed=3, precision="bf16", eps_gain=TRAJ_EPS_GAIN, max_batch=4, max_steps=10, args=())
    sched = _fused_sched()
    xT = inputs("latents", 2, seed=23)
    faces = [inputs("ln_face", 2, seed=9), inputs("ln_face", 2, seed=10) * 1.6 - 0.3]   # batch 2 leaves [0, 1]
    outs, ptrs = [], []
    for ln_face in faces:
        images = H.ddim_sample_images(ln_face.cuda(), m, vae, crm, sched, 0.18215, 10, x_T=xT.cuda())
        m.denoiser.engine().synchronize()
        outs.append(images.cpu())
        ptrs.append(tuple(t.data_ptr() for t, _ in m._cond_src))
        del images
    wants = [_oracle_pipeline(sd_cr, sd, vae, f, xT, 10) for f in faces]
    qs = [psnr(o, w[0]) for o, w in zip(outs, wants)]
    cross = psnr(outs[1], wants[0][0])
    oob = float(((wants[1][1] < 0) | (wants[1][1] > 1)).float().mean())
    print(f"two batches: PSNR vs own oracle {qs[0]:.2f} / {qs[1]:.2f} dB; batch 2 vs batch 1's oracle {cross:.2f} dB; "
          f"addresses recycled: {ptrs[0] == ptrs[1]}; CR output outside [0,1]: {100 * oob:.1f} %")
    assert qs[0] >= 35.0 and qs[1] >= 35.0
    assert cross < 30.0                     # the two batches really differ
    # the module API, called like the reference's loop body (refiner.py:32-38), on fresh tensors per batch
    eps = []
    with torch.no_grad():
        for ln_face in faces:
            cr_face = crm(ln_face.cuda())
            cr_latent = H.encode_latent(vae, cr_face, 0.18215, 128).float().contiguous()
            eps.append(m(xT.cuda(), 500, cr_face, cr_latent).sample.cpu())
            del cr_face, cr_latent
        fresh, _ = build(H.FacialRefiner, seed=3, precision="bf16", eps_gain=TRAJ_EPS_GAIN, max_batch=4, max_steps=10, args=())
        cr_face = crm(faces[1].cuda())
        cr_latent = H.encode_latent(vae, cr_face, 0.18215, 128).float().contiguous()
        want2 = fresh(xT.cuda(), 500, cr_face, cr_latent).sample.cpu()
    assert torch.equal(eps[1], want2)       # batch 2 through the used module == batch 2 through a fresh module
    assert not torch.equal(eps[0], eps[1])
    crm.invalidate()
    m.denoiser.invalidate()
    fresh.denoiser.invalidate()


def test_update_kernel_reproduces_diffusers_known_answers(raw):
    """The CUDA update kernel (`hd_sampler_update`), driven step by step through the full loops of diffusers' own
    scheduler tests (tests/schedulers/test_scheduler_ddim.py / test_scheduler_ddpm.py `test_full_loop_no_noise`; fixtures
    and harness in tests/util.py, validated on CPU by tests/test_schedulers.py): 10 DDIM steps and all 1000 ancestral
    DDPM steps with `torch.manual_seed(0)` noise must land on upstream's asserted checksums, at upstream's tolerances."""
    from util import KAT_CFG, KAT_DDIM, KAT_DDPM, kat_ddim_loop, kat_ddpm_loop
    p = S.DDIMScheduler(prediction_type="epsilon", **KAT_CFG)
    p.set_timesteps(10)
    c = p.step_coefficients(eta=0.0)
    s, m, _ = kat_ddim_loop(lambda i, x, eps: _update(raw, x, eps, c, i), p.timesteps.tolist())
    print(f"DDIM-10 through the kernel: sum |x| {s:.4f} (diffusers {KAT_DDIM[0]}), mean {m:.6f} ({KAT_DDIM[1]})")
    assert abs(s - KAT_DDIM[0]) < 1e-2 and abs(m - KAT_DDIM[1]) < 1e-3
    q = S.DDPMScheduler(prediction_type="epsilon", **KAT_CFG)
    q.set_timesteps(1000)
    d = q.step_coefficients()
    s, m, _ = kat_ddpm_loop(lambda i, x, eps, z: _update(raw, x, eps, d, i, noise=z.cuda()))
    print(f"DDPM-1000 through the kernel: sum |x| {s:.4f} (diffusers {KAT_DDPM[0]}), mean {m:.6f} ({KAT_DDPM[1]})")
    assert abs(s - KAT_DDPM[0]) < 1e-2 and abs(m - KAT_DDPM[1]) < 1e-3
