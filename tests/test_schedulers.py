"""CPU: scheduler restatement (diffusers 0.32.2 semantics, PARITY UNPINNED by the reference) —
self-consistency checks of SURVEY.md App. C, and the product's per-step coefficients against it."""
import numpy as np
import pytest
import torch

from oracle import schedulers_ref as R
from hifidiff_b200 import schedulers as S


def test_alpha_table_anchors():
    s = R.DDIMSchedulerRef()
    a = s.alphas_cumprod
    assert abs(float(a[0]) - 0.999900) < 1e-6
    assert abs(float(a[499]) - 0.33319) < 2e-5
    assert abs(float(a[980]) - 1.070e-3) < 2e-6
    assert abs(float(a[999]) - 7.334e-4) < 2e-7
    assert bool((a[1:] < a[:-1]).all())


def test_timesteps_leading_spacing():
    s = R.DDIMSchedulerRef()
    s.set_timesteps(50)
    assert s.timesteps.tolist() == list(range(980, -1, -20))
    s.set_timesteps(1000)
    assert s.timesteps.tolist() == list(range(999, -1, -1))
    with pytest.raises(ValueError):
        s.set_timesteps(1001)


def test_ddim_inverts_add_noise():
    """DDIM eta=0 fed the true eps maps add_noise(x0, eps, t) onto add_noise(x0, eps, prev_t); last step returns x0."""
    s = R.DDIMSchedulerRef(clip_sample=False)
    s.set_timesteps(50)
    g = torch.Generator().manual_seed(0)
    x0, eps = torch.randn(2, 4, 16, 16, generator=g), torch.randn(2, 4, 16, 16, generator=g)
    for t in (980, 500, 20):
        xt = s.add_noise(x0, eps, torch.tensor([t, t]))
        want = s.add_noise(x0, eps, torch.tensor([t - 20, t - 20]))
        assert float((s.step(eps, t, xt) - want).abs().max()) < 2e-5
    xt = s.add_noise(x0, eps, torch.tensor([0, 0]))
    assert float((s.step(eps, 0, xt) - x0).abs().max()) < 1e-6


def test_ddpm_mean_equals_ddim_eta1_mean():
    d = R.DDPMSchedulerRef(clip_sample=False)
    i = R.DDIMSchedulerRef(clip_sample=False)
    d.set_timesteps(50)
    i.set_timesteps(50)
    g = torch.Generator().manual_seed(1)
    x, eps = torch.randn(1, 4, 16, 16, generator=g), torch.randn(1, 4, 16, 16, generator=g)
    z = torch.zeros_like(x)
    for t in (980, 500, 20):
        assert float((d.step(eps, t, x, variance_noise=z) - i.step(eps, t, x, eta=1.0, variance_noise=z)).abs().max()) < 1e-4  # fp32 cancellation in 1 - a_prev - std^2


def test_clip_sample_variant():
    s = R.DDIMSchedulerRef(clip_sample=True, clip_sample_range=3.0)   # test_refiner.py:166-171
    s.set_timesteps(50)
    x = torch.full((1, 4, 16, 16), 50.0)
    eps = torch.zeros_like(x)
    out = s.step(eps, 500, x)
    a_p = s.alphas_cumprod[480]
    assert float((out - a_p ** 0.5 * 3.0).abs().max()) < 1e-6


def _apply(c, x, eps, z):
    x0 = (x - np.float32(c.sqrt_beta_prod) * eps) / np.float32(c.sqrt_alpha_prod)
    if c.clip > 0:
        x0 = x0.clamp(-c.clip, c.clip)
    return np.float32(c.k_x0) * x0 + np.float32(c.k_eps) * eps + np.float32(c.k_x) * x + np.float32(c.k_noise) * z


@pytest.mark.parametrize("steps", [50, 1000])
@pytest.mark.parametrize("clip", [False, True])
def test_product_coefficients_match_oracle_step(steps, clip):
    g = torch.Generator().manual_seed(2)
    x, eps, z = (torch.randn(2, 4, 16, 16, generator=g) for _ in range(3))
    for P, O, kw in ((S.DDIMScheduler, R.DDIMSchedulerRef, {}), (S.DDPMScheduler, R.DDPMSchedulerRef, {})):
        p = P(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon", clip_sample=clip,
              clip_sample_range=3.0)
        o = O(beta_schedule="scaled_linear", clip_sample=clip, clip_sample_range=3.0)
        p.set_timesteps(steps)
        o.set_timesteps(steps)
        assert p.timesteps.tolist() == o.timesteps.tolist()
        coefs = p.step_coefficients()
        for idx in (0, 1, steps // 2, steps - 2, steps - 1):
            t = int(o.timesteps[idx])
            want = o.step(eps, t, x, variance_noise=z) if O is R.DDPMSchedulerRef else o.step(eps, t, x)
            got = _apply(coefs[idx], x, eps, z)
            assert float((got - want).abs().max()) <= 3e-6 * max(1.0, float(want.abs().max())), (P.__name__, idx)


def test_product_add_noise_and_ctor_kwargs():
    p = S.DDPMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon")
    o = R.DDPMSchedulerRef()
    g = torch.Generator().manual_seed(3)
    x0, n = torch.randn(3, 4, 16, 16, generator=g), torch.randn(3, 4, 16, 16, generator=g)
    t = torch.tensor([0, 500, 999])
    assert torch.equal(p.add_noise(x0, n, t), o.add_noise(x0, n, t))


def test_unsupported_diffusers_kwargs_raise():
    """Non-default spacing / offset / variance options are not silently ignored."""
    S.DDIMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear", timestep_spacing="leading", steps_offset=0,
                    set_alpha_to_one=True)
    for kw in ({"timestep_spacing": "trailing"}, {"steps_offset": 1}, {"set_alpha_to_one": False},
               {"variance_type": "fixed_large"}, {"thresholding": True}, {"rescale_betas_zero_snr": True}):
        with pytest.raises(NotImplementedError):
            S.DDPMScheduler(beta_schedule="scaled_linear", **kw)
    with pytest.raises(TypeError):
        S.DDIMScheduler(not_a_diffusers_argument=1)


def test_oracle_against_the_published_formulas_in_float64():
    """Independent anchor for the (parity-unpinned) scheduler oracle: the closed forms of the papers, evaluated in
    numpy float64 from the beta schedule alone — Ho et al. 2020 eq. 6-7 (posterior mean / beta-tilde), Song et al.
    2021 eq. 12 (DDIM) — against the fp32 op-by-op restatement of diffusers' step()."""
    T = 1000
    betas = np.linspace(1e-4 ** 0.5, 2e-2 ** 0.5, T, dtype=np.float64) ** 2
    abar = np.cumprod(1.0 - betas)
    g = torch.Generator().manual_seed(4)
    x, eps, z = (torch.randn(2, 4, 16, 16, generator=g) for _ in range(3))
    xd, ed, zd = x.double().numpy(), eps.double().numpy(), z.double().numpy()
    for n in (50, 1000):
        ddpm, ddim = R.DDPMSchedulerRef(clip_sample=False), R.DDIMSchedulerRef(clip_sample=False)
        ddpm.set_timesteps(n)
        ddim.set_timesteps(n)
        stride = T // n
        for t in (T - stride, 25 * stride, stride, 0):
            a_t = abar[t]
            a_p = abar[t - stride] if t - stride >= 0 else 1.0
            x0 = (xd - np.sqrt(1 - a_t) * ed) / np.sqrt(a_t)
            # DDIM, sigma = 0
            want = np.sqrt(a_p) * x0 + np.sqrt(1 - a_p) * ed
            got = ddim.step(eps, t, x).double().numpy()
            assert np.abs(got - want).max() <= 2e-5 * max(1.0, np.abs(want).max()), ("ddim", n, t)
            # DDPM posterior q(x_{t-1} | x_t, x0) with the strided alpha_t = abar_t / abar_prev
            alpha = a_t / a_p
            mean = np.sqrt(a_p) * (1 - alpha) / (1 - a_t) * x0 + np.sqrt(alpha) * (1 - a_p) / (1 - a_t) * xd
            var = (1 - a_p) / (1 - a_t) * (1 - alpha)
            want = mean + (np.sqrt(var) * zd if t > 0 else 0.0)
            got = ddpm.step(eps, t, x, variance_noise=z).double().numpy()
            # 1 - abar_t cancels in fp32 as t -> 0 (abar_1 = 0.99978): diffusers' op order carries ~6e-5 there
            assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max()), ("ddpm", n, t)
