"""CPU: the oracle restatement reproduces the reference's outputs stored under tests/golden/.

The fixtures were produced by tests/golden/make_golden.py from the unmodified reference; these
tests regenerate the same weights/inputs from seeds and run only the oracle, so they also run on
machines without /root/reference.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cond_ref, denoiser_ref, schedulers_ref
from hifidiff_b200 import testing
import hifidiff_b200 as H

from util import (GOLDEN, LEVELS, TRAJ_EPS_GAIN, block_shapes, gen, golden, hca_shapes, inputs, rel_l2, state_for)

TOL = 2e-6  # same torch build: bit-identical in practice; allows a different BLAS blocking


def test_naf_block_levels():
    g = golden("naf_blocks.npz")
    for lvl, (c, n) in enumerate(LEVELS):
        sd = state_for(block_shapes(c), seed=10 + lvl)
        x = torch.randn((2, c, n, n), generator=gen(400 + lvl))
        temb = torch.randn((2, 512), generator=gen(500 + lvl))
        y = denoiser_ref.cond_naf_block(sd, "", x, temb)
        assert rel_l2(y, g[f"level{lvl}"]) < TOL, lvl


def test_hca_levels():
    g = golden("hca.npz")
    for j, (c, n) in enumerate(LEVELS[::-1]):
        sd = state_for(hca_shapes(c), seed=20 + j)
        f_g = torch.randn((2, c, n, n), generator=gen(600 + j))
        f_d = torch.randn((2, c, n, n), generator=gen(700 + j))
        y = denoiser_ref.hca(sd, "", f_g, f_d)
        assert rel_l2(y, g[f"hca{j}"]) < TOL, j


@pytest.fixture(scope="module")
def denoiser_state():
    with torch.device("meta"):
        m = H.Denoiser(16)
    return state_for(m, seed=1)


def test_denoiser_step(denoiser_state):
    g = golden("denoiser_step.npz")
    x = inputs("latents", 2)
    taps = {}
    with torch.no_grad():
        y = denoiser_ref.denoiser_forward(denoiser_state, x, torch.from_numpy(g["t"]), taps)
        y500 = denoiser_ref.denoiser_forward(denoiser_state, x, 500)
    assert rel_l2(y, g["eps"]) < TOL
    assert rel_l2(y500, g["eps_t500"]) < TOL
    assert rel_l2(taps["intro"], g["tap_intro"]) < TOL
    assert rel_l2(taps["middle_blks.7"], g["tap_mid7"]) < TOL
    assert rel_l2(taps["decoders.3.1"], g["tap_dec31"]) < TOL
    assert rel_l2(taps["time_mlp"], g["tap_time"]) < TOL
    names = [str(s) for s in g["tap_names"]]
    assert names == list(taps.keys())
    for k, (mean, std, norm) in zip(names, g["tap_stats"]):
        assert abs(float(taps[k].norm()) - norm) <= 1e-5 * norm, k


def test_fused_step():
    g = golden("fused_step.npz")
    with torch.device("meta"):
        m = H.FusedDenoiser(16)
    sd = state_for(m, seed=2)
    x = inputs("latents", 2, seed=1)
    priors, ident = testing.synthetic_condition(2, 16, seed=0)
    taps = {}
    with torch.no_grad():
        y = denoiser_ref.fused_denoiser_forward(sd, x, torch.from_numpy(g["t"]), priors, ident, taps)
    assert rel_l2(y, g["eps"]) < TOL
    assert rel_l2(taps["hcas.0"], g["tap_hca0"]) < TOL
    assert rel_l2(taps["hcas.4"], g["tap_hca4"]) < TOL
    assert rel_l2(taps["ups.0"], g["tap_up0"]) < TOL


def test_refiner_step():
    g = golden("refiner_step.npz")
    with torch.device("meta"):
        m = H.FacialRefiner()
    sd = state_for(m, seed=3)
    x = inputs("latents", 1, seed=2)
    taps = {}
    with torch.no_grad():
        y = cond_ref.refiner_forward(sd, x, torch.tensor([640]), inputs("cr_face", 1), inputs("cr_latent", 1), taps)
    assert rel_l2(y, g["eps"]) < TOL
    for j in range(5):
        assert rel_l2(taps[f"prior{j}"], g[f"prior{j}"]) < TOL
    assert rel_l2(taps["identity"], g["identity"]) < TOL


def test_ddim_trajectory_prefix(denoiser_state):
    """First 10 of the 50 DDIM steps of the stored reference trajectory (the full 50 run on the GPU test)."""
    g = golden("denoiser_ddim50.npz")
    sd = dict(denoiser_state)
    sd["ending.weight"] = sd["ending.weight"] * TRAJ_EPS_GAIN
    sd["ending.bias"] = sd["ending.bias"] * TRAJ_EPS_GAIN
    sched = schedulers_ref.DDIMSchedulerRef(clip_sample=False)
    sched.set_timesteps(50)
    x = inputs("latents", 1, seed=7)
    with torch.no_grad():
        for i, t in enumerate(sched.timesteps.tolist()[:10]):
            x = sched.step(denoiser_ref.denoiser_forward(sd, x, torch.full((1,), t, dtype=torch.long)), t, x)
            if i == 0:
                assert rel_l2(x, g["x_after_step0"]) < TOL
    assert rel_l2(x, g["x_after_step9"]) < 1e-5


def test_state_dict_layout_matches_reference():
    with open(os.path.join(GOLDEN, "state_dict_layout.json")) as f:
        layout = json.load(f)
    for name, cls, args in (("Denoiser", H.Denoiser, (16,)), ("FusedDenoiser", H.FusedDenoiser, (16,)),
                            ("FacialRefiner", H.FacialRefiner, ())):
        with torch.device("meta"):
            m = cls(*args)
        mine = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in m.state_dict().items()]
        assert mine == layout[name], name
    assert len(layout["FusedDenoiser"]) == 787 and len(layout["Denoiser"]) == 660 and len(layout["FacialRefiner"]) == 1460


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not mounted")
def test_oracle_against_live_reference_block():
    from oracle import ref_shim
    ref = ref_shim.load()
    blk = ref.ConditionalNAFBlock(256, 512)
    sd = state_for(blk, seed=99)
    blk.load_state_dict(sd)
    x, temb = torch.randn(3, 256, 8, 8, generator=gen(1)), torch.randn(3, 512, generator=gen(2))
    with torch.no_grad():
        assert rel_l2(denoiser_ref.cond_naf_block(sd, "", x, temb), blk([x, temb])[0]) < TOL


def test_fused_step_latent32():
    """Latent size 32 (`--image_res 256`): the oracle reproduces the reference's eps, the small-spatial taps and the
    FPG priors stored by tests/golden/make_golden_s32.py (idc_conv reshape to (B, 2048, 2, 2), model.py:198-200,245-246)."""
    g = golden("fused_step_s32.npz")
    with torch.device("meta"):
        m = H.FusedDenoiser(32)
    sd = state_for(m, seed=2)
    assert tuple(sd["idc_conv.weight"].shape) == (8192, 2048, 1, 1)
    x = torch.randn((2, 4, 32, 32), generator=torch.Generator().manual_seed(701))
    priors, ident = testing.synthetic_condition(2, 32, seed=0)
    taps = {}
    with torch.no_grad():
        y = denoiser_ref.fused_denoiser_forward(sd, x, torch.from_numpy(g["t"]), priors, ident, taps)
    assert rel_l2(y, g["eps"]) < TOL
    for k in ("downs.3", "middle_blks.7", "hcas.0", "ups.0", "decoders.0.1", "hcas.1"):
        assert rel_l2(taps[k], g["tap_" + k.replace(".", "_")]) < TOL, k
    with torch.device("meta"):
        fpg = H.FacialPriorGuidance()
    sdf = state_for(fpg, seed=7)
    lat = torch.randn((2, 4, 32, 32), generator=torch.Generator().manual_seed(731))
    with torch.no_grad():
        pri = cond_ref.fpg_forward(sdf, lat, "")
    for j in range(3):
        assert rel_l2(pri[j], g[f"fpg_prior{j}"]) < TOL, j
    for j in range(5):
        assert abs(float(pri[j].mean()) - g["fpg_prior_stats"][j][0]) < 1e-5 and abs(float(pri[j].std()) - g["fpg_prior_stats"][j][1]) < 1e-5
