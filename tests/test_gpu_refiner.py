"""GPU: FacialRefiner.forward(latents, t, cr_face, cr_latent) against the reference's stored output."""
import pytest
import torch

import hifidiff_b200 as H

from gpu_util import build
from util import golden, inputs, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("bf16", 1e-2)])
def test_refiner_step(prec, tol):
    g = golden("refiner_step.npz")
    with torch.no_grad():
        m, sd = build(H.FacialRefiner, seed=3, precision=prec, max_batch=2, args=())
        x = inputs("latents", 1, seed=2).cuda()
        cr_face, cr_latent = inputs("cr_face", 1).cuda(), inputs("cr_latent", 1).cuda()
        out = m(x, torch.tensor([640]), cr_face, cr_latent).sample
        priors, ident = m.condition(cr_face, cr_latent)
        m.denoiser.engine().synchronize()
    for j in range(5):
        assert rel_l2(priors[j], g[f"prior{j}"]) < 1e-4, j   # PyTorch/cuDNN eager ("next" row)
    assert rel_l2(ident, g["identity"]) < 1e-4
    assert rel_l2(out, g["eps"]) <= tol
    m.denoiser.invalidate()
