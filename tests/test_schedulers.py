"""CPU: scheduler restatement (diffusers 0.32.2 semantics; the reference itself holds no vector at this boundary) —
the known-answer vectors of diffusers' own scheduler tests (bottom of this file), the self-consistency checks of
SURVEY.md App. C, and the product's per-step coefficients against the oracle."""
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
    """Second, independent anchor for the scheduler oracle (next to diffusers' known answers below): the closed forms of the papers, evaluated in
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


# ------------------------------------------------------------------------------------------------------------------
# Known-answer vectors of diffusers' OWN test-suite (huggingface/diffusers, tests/schedulers/test_scheduler_ddim.py and
# test_scheduler_ddpm.py, unchanged across the 0.2x / 0.3x releases incl. the pinned 0.32.2).  The package is not on
# this box, so the fixtures below restate that suite's `dummy_sample_deter`, `dummy_model` and `get_scheduler_config`
# (num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear", clip_sample=True) and the numbers
# its assertions hold.  These pin BOTH scheduler restatements (the oracle's step() and the product's coefficient table)
# to the upstream implementation; upstream's tolerances are 1e-2 on the sum and 1e-3 on the mean, ours are tighter.
# ------------------------------------------------------------------------------------------------------------------
def _dummy_sample_deter():
    b, c, h, w = 4, 3, 8, 8
    n = b * c * h * w
    return (torch.arange(n).reshape(c, h, w, b) / n).permute(3, 0, 1, 2)


def _dummy_model(sample, t):
    return sample * t / (t + 1)


_KAT_CFG = dict(num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear", clip_sample=True)
# (config overrides, |x|.sum(), |x|.mean()) — DDIMSchedulerTest.test_full_loop_no_noise / ..._with_set_alpha_to_one
_DDIM_KATS = [({}, 172.0067, 0.223967), ({"beta_start": 0.01}, 149.8295, 0.1951)]
# DDPMSchedulerTest.test_full_loop_no_noise: 1000 ancestral steps, noise from torch.manual_seed(0)
_DDPM_KAT = (258.9606, 0.3372)


def _ddim_loop(step_fn, timesteps):
    x = _dummy_sample_deter()
    for i, t in enumerate(timesteps):
        x = step_fn(i, t, x, _dummy_model(x, t))
    return x


@pytest.mark.parametrize("over,want_sum,want_mean", _DDIM_KATS)
def test_ddim_full_loop_matches_diffusers_known_answers(over, want_sum, want_mean):
    cfg = dict(_KAT_CFG, **over)
    o = R.DDIMSchedulerRef(**cfg)
    o.set_timesteps(10)
    x = _ddim_loop(lambda i, t, x, eps: o.step(eps, t, x, eta=0.0), o.timesteps.tolist())
    assert abs(float(x.abs().sum()) - want_sum) < 2e-3 and abs(float(x.abs().mean()) - want_mean) < 1e-4
    # the product: host coefficient table + the kernel's elementwise update (restated by _apply above)
    p = S.DDIMScheduler(prediction_type="epsilon", **cfg)
    p.set_timesteps(10)
    assert p.timesteps.tolist() == o.timesteps.tolist() == list(range(900, -1, -100))
    coefs = p.step_coefficients(eta=0.0)
    y = _ddim_loop(lambda i, t, x, eps: _apply(coefs[i], x, eps, torch.zeros_like(x)), p.timesteps.tolist())
    assert abs(float(y.abs().sum()) - want_sum) < 2e-3 and abs(float(y.abs().mean()) - want_mean) < 1e-4


def test_ddpm_full_loop_matches_diffusers_known_answers():
    want_sum, want_mean = _DDPM_KAT

    def run(step_fn):
        g = torch.manual_seed(0)                      # upstream: generator = torch.manual_seed(0)
        x = _dummy_sample_deter()
        for i, t in enumerate(reversed(range(1000))):
            eps = _dummy_model(x, t)
            z = torch.randn(eps.shape, generator=g)   # upstream draws a variance_noise at every step, t = 0 included
            x = step_fn(i, t, x, eps, z)
        return x

    o = R.DDPMSchedulerRef(**_KAT_CFG)                # num_inference_steps unset: prev_t = t - 1
    x = run(lambda i, t, x, eps, z: o.step(eps, t, x, variance_noise=z))
    assert abs(float(x.abs().sum()) - want_sum) < 2e-3 and abs(float(x.abs().mean()) - want_mean) < 1e-4
    p = S.DDPMScheduler(prediction_type="epsilon", **_KAT_CFG)
    p.set_timesteps(1000)
    coefs = p.step_coefficients()
    y = run(lambda i, t, x, eps, z: _apply(coefs[i], x, eps, z))
    assert abs(float(y.abs().sum()) - want_sum) < 2e-3 and abs(float(y.abs().mean()) - want_mean) < 1e-4


def test_variances_match_diffusers_known_answers():
    """DDIMSchedulerTest.test_variance (`_get_variance(t, prev_t)`) and DDPMSchedulerTest.test_variance
    (`_get_variance(t)`, fixed_small), tolerance 1e-5 as upstream."""
    o = R.DDIMSchedulerRef(**_KAT_CFG)
    ac = o.alphas_cumprod

    def ddim_var(t, p):
        return float(((1 - ac[p]) / (1 - ac[t])) * (1 - ac[t] / ac[p]))

    for t, p, want in ((0, 0, 0.0), (420, 400, 0.14771), (980, 960, 0.32460), (487, 486, 0.00979), (999, 998, 0.02)):
        assert abs(ddim_var(t, p) - want) < 1e-5, (t, p)
    # the product's k_noise is sqrt(variance): DDPM at t = 487 / 999 (t = 0 adds no noise), DDIM eta = 1 at 980 -> 960
    p = S.DDPMScheduler(prediction_type="epsilon", **_KAT_CFG)
    p.set_timesteps(1000)
    cf = {int(c.timestep): c for c in p.step_coefficients()}
    assert abs(cf[487].k_noise ** 2 - 0.00979) < 1e-5 and abs(cf[999].k_noise ** 2 - 0.02) < 1e-5 and cf[0].k_noise == 0.0
    d = S.DDIMScheduler(prediction_type="epsilon", **_KAT_CFG)
    d.set_timesteps(50)
    cd = {int(c.timestep): c for c in d.step_coefficients(eta=1.0)}
    assert abs(cd[980].k_noise ** 2 - 0.32460) < 1e-5 and abs(cd[420].k_noise ** 2 - 0.14771) < 1e-5


def test_known_answer_loops_on_the_kernel_shape():
    """The harness the GPU test `test_update_kernel_reproduces_diffusers_known_answers` drives through the CUDA update
    kernel, run here with the kernel's arithmetic restated (`_apply`): one padded (1,4,16,16) face, flattened in logical
    order, reproduces the same known answers, and the padding stays 0."""
    from util import KAT_CFG, KAT_DDIM, KAT_DDPM, kat_ddim_loop, kat_ddpm_loop
    p = S.DDIMScheduler(prediction_type="epsilon", **KAT_CFG)
    p.set_timesteps(10)
    c = p.step_coefficients(eta=0.0)
    s, m, pad = kat_ddim_loop(lambda i, x, eps: _apply(c[i], x, eps, torch.zeros_like(x)), p.timesteps.tolist())
    assert abs(s - KAT_DDIM[0]) < 2e-3 and abs(m - KAT_DDIM[1]) < 1e-4 and pad == 0.0
    q = S.DDPMScheduler(prediction_type="epsilon", **KAT_CFG)
    q.set_timesteps(1000)
    d = q.step_coefficients()
    s, m, pad = kat_ddpm_loop(lambda i, x, eps, z: _apply(d[i], x, eps, z))
    assert abs(s - KAT_DDPM[0]) < 2e-3 and abs(m - KAT_DDPM[1]) < 1e-4 and pad == 0.0   # the noise is zero-padded too
