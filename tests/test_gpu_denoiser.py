"""GPU: Denoiser / FusedDenoiser single step through the module API (-> ctypes -> C ABI -> sm_100a)
against the CPU oracle and the reference's stored outputs, per layer.

Tolerances (BASELINE.json north_star): rel-L2 <= 1e-5 in the fp32 mode, <= 1e-2 in bf16.
"""
import numpy as np
import pytest
import torch

import hifidiff_b200 as H
from hifidiff_b200 import testing
from oracle import denoiser_ref

from gpu_util import build
from util import golden, inputs, rel_l2

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-5, "bf16": 1e-2}
DEN_TAPS = (["time_mlp", "intro"] + [f"encoders.{l}.{i}" for l, n in enumerate((2, 2, 4, 8)) for i in range(n)] +
            [f"downs.{l}" for l in range(4)] + [f"middle_blks.{i}" for i in range(8)] +
            [f"ups.{l}" for l in range(4)] + [f"decoders.{l}.{i}" for l in range(4) for i in range(2)])
FUSED_TAPS = DEN_TAPS + [f"hcas.{j}" for j in range(5)]


@pytest.fixture(scope="module", params=["fp32", "bf16"])
def denoiser(request):
    m, sd = build(H.Denoiser, seed=1, precision=request.param, max_batch=64)
    yield m, sd, request.param
    m.invalidate()


@pytest.fixture(scope="module", params=["fp32", "bf16"])
def fused(request):
    m, sd = build(H.FusedDenoiser, seed=2, precision=request.param, max_batch=64)
    yield m, sd, request.param
    m.invalidate()


def test_denoiser_step_per_layer(denoiser):
    m, sd, prec = denoiser
    g = golden("denoiser_step.npz")
    x = inputs("latents", 2)
    t = torch.from_numpy(g["t"])
    out, taps = m.forward_with_taps(x.cuda(), t.cuda(), DEN_TAPS)
    m.engine().synchronize()
    ref_taps = {}
    with torch.no_grad():
        ref = denoiser_ref.denoiser_forward(sd, x, t, ref_taps)
    assert rel_l2(ref, g["eps"]) < 2e-6                      # oracle == stored reference output
    worst = max((rel_l2(taps[k], ref_taps[k]), k) for k in DEN_TAPS)
    assert worst[0] <= TOL[prec], f"layer {worst[1]} rel-L2 {worst[0]:.3e}"
    assert rel_l2(out.sample, g["eps"]) <= TOL[prec]
    assert out.sample.dtype == torch.float32 and tuple(out.sample.shape) == (2, 4, 16, 16)


def test_denoiser_timestep_forms(denoiser):
    m, sd, prec = denoiser
    g = golden("denoiser_step.npz")
    x = inputs("latents", 2).cuda()
    base = m(x, 500).sample
    assert rel_l2(base, g["eps_t500"]) <= TOL[prec]
    for tv in (500.0, torch.tensor(500), torch.tensor([500]), torch.tensor([500, 500]), torch.tensor([500.0, 500.0]).cuda()):
        assert torch.equal(m(x, tv).sample, base)


def test_fused_step_per_layer(fused):
    m, sd, prec = fused
    g = golden("fused_step.npz")
    x = inputs("latents", 2, seed=1)
    priors, ident = testing.synthetic_condition(2, 16, seed=0)
    t = torch.from_numpy(g["t"])
    out, taps = m.forward_with_taps(x.cuda(), t.cuda(), FUSED_TAPS, [p.cuda() for p in priors], ident.cuda())
    m.engine().synchronize()
    ref_taps = {}
    with torch.no_grad():
        ref = denoiser_ref.fused_denoiser_forward(sd, x, t, priors, ident, ref_taps)
    assert rel_l2(ref, g["eps"]) < 2e-6
    worst = max((rel_l2(taps[k], ref_taps[k]), k) for k in FUSED_TAPS)
    assert worst[0] <= TOL[prec], f"layer {worst[1]} rel-L2 {worst[0]:.3e}"
    assert rel_l2(out.sample, g["eps"]) <= TOL[prec]


def test_fused_condition_cache_and_batch_one(fused):
    m, sd, prec = fused
    priors, ident = testing.synthetic_condition(3, 16, seed=5)
    pc, ic = [p.cuda() for p in priors], ident.cuda()
    x = inputs("latents", 3, seed=4)
    a = m(x.cuda(), 77, pc, ic).sample
    b = m(x.cuda(), torch.tensor([77]), pc, ic).sample       # cached condition, length-1 broadcast (model.py:228-229)
    assert torch.equal(a, b)
    with torch.no_grad():
        ref = denoiser_ref.fused_denoiser_forward(sd, x, 77, priors, ident)
    assert rel_l2(a, ref) <= TOL[prec]
    # ragged batch: one face alone gives the same face-0 result (no cross-face reduction anywhere)
    one = m(x[:1].cuda(), 77, [p[:1].contiguous() for p in pc], ic[:1].contiguous()).sample
    assert rel_l2(one, ref[:1]) <= TOL[prec]


def test_config2_batch64_bf16():
    """BASELINE.json configs[1]: single-timestep forward, batch 64, bf16, per-layer vs the reference arithmetic."""
    m, sd = build(H.FusedDenoiser, seed=2, precision="bf16", max_batch=64)
    x = inputs("latents", 64, seed=9)
    priors, ident = testing.synthetic_condition(64, 16, seed=9)
    taps_wanted = ["intro", "encoders.0.1", "encoders.3.7", "middle_blks.7", "hcas.0", "decoders.1.1", "hcas.4"]
    out, taps = m.forward_with_taps(x.cuda(), 500, taps_wanted, [p.cuda() for p in priors], ident.cuda())
    m.engine().synchronize()
    ref_taps = {}
    with torch.no_grad():
        ref = denoiser_ref.fused_denoiser_forward(sd, x, 500, priors, ident, ref_taps)
    for k in taps_wanted:
        assert rel_l2(taps[k], ref_taps[k]) <= 1e-2, k
    assert rel_l2(out.sample, ref) <= 1e-2
    m.invalidate()


def test_fused_face_kernel_taps_and_plans_agree():
    """The production plan (fused per-face block kernel at 16x16) against the oracle at the taps it exposes,
    and against the one-kernel-per-op plan that serves all other taps."""
    m, sd = build(H.FusedDenoiser, seed=2, precision="bf16", max_batch=64)
    for batch in (3, 40):
        x = inputs("latents", batch, seed=11)
        priors, ident = testing.synthetic_condition(batch, 16, seed=11)
        cond = ([p.cuda() for p in priors], ident.cuda())
        t = torch.arange(batch) * 7 + 3 if batch == 3 else 321
        t_dev = t.cuda() if torch.is_tensor(t) else t
        fast_names = ["encoders.0.1", "encoders.1.1", "middle_blks.7", "decoders.2.1", "decoders.3.1"]
        out_fast, taps_fast = m.forward_with_taps(x.cuda(), t_dev, fast_names, *cond)
        out_dbg, taps_dbg = m.forward_with_taps(x.cuda(), t_dev, fast_names + ["intro"], *cond)
        plain = m(x.cuda(), t_dev, *cond).sample
        m.engine().synchronize()
        ref_taps = {}
        with torch.no_grad():
            ref = denoiser_ref.fused_denoiser_forward(sd, x, t, priors, ident, ref_taps)
        assert torch.equal(plain, out_fast.sample)               # the tapped run used the production plan
        for k in fast_names:
            assert rel_l2(taps_fast[k], ref_taps[k]) <= 1e-2, (batch, k)
            assert rel_l2(taps_fast[k], taps_dbg[k]) <= 6e-3, (batch, k)
        assert rel_l2(out_fast.sample, ref) <= 1e-2
        assert rel_l2(out_fast.sample, out_dbg.sample) <= 8e-3
    m.invalidate()


@pytest.mark.parametrize("batch", [5, 130, 256])
def test_sca_rescale_in_gemm_epilogue_is_exact(batch):
    """At 1x1 spatial the SCA rescale `x * sca(x)` (conditional_naf.py:119) rides in the SCA GEMM's epilogue
    (EPI_MUL, split-K over a cluster at these sizes) instead of a separate scale_rows launch: one, two (ragged) and
    two full 128-row tiles, per-face timesteps.  Same fp32 product, rounded to bf16 once either way: bit-identical
    to the two-kernel form (HD_SCA_MUL=0), 8 launches fewer per step, and within tolerance of the oracle."""
    import os
    x = inputs("latents", batch, seed=13)
    priors, ident = testing.synthetic_condition(batch, 16, seed=13)
    cond = ([p.cuda() for p in priors], ident.cuda())
    t = (torch.arange(batch) * 3 + 1) % 1000
    outs, launches = [], []
    for flag in ("1", "0"):
        os.environ["HD_SCA_MUL"] = flag
        try:
            m, sd = build(H.FusedDenoiser, seed=2, precision="bf16", max_batch=256)
            out, taps = m.forward_with_taps(x.cuda(), t.cuda(), ["middle_blks.7"], *cond)
            m.engine().synchronize()
            launches.append(m.engine().info().launches_per_step)
            outs.append((out.sample.clone(), taps["middle_blks.7"].clone()))
            m.invalidate()
        finally:
            del os.environ["HD_SCA_MUL"]
    print(f"SCA epilogue B={batch}: launches/step {launches[0]} (fused) vs {launches[1]}")
    assert launches[1] - launches[0] == 8
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][0], outs[1][0])
    if batch <= 8:
        ref_taps = {}
        with torch.no_grad():
            ref = denoiser_ref.fused_denoiser_forward(sd, x, t, priors, ident, ref_taps)
        assert rel_l2(outs[0][1], ref_taps["middle_blks.7"]) <= 1e-2
        assert rel_l2(outs[0][0], ref) <= 1e-2


def test_errors_are_loud(denoiser):
    m, _, _ = denoiser
    with pytest.raises(ValueError):
        m(torch.zeros(1, 4, 8, 8).cuda(), 1)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, 16, 16), 1)
