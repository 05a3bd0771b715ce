"""GPU: FacialRefiner.forward(latents, t, cr_face, cr_latent) against the reference's stored output, with the
FPG prior network on the native sm_100a kernels and on PyTorch eager."""
import pytest
import torch

import hifidiff_b200 as H

from gpu_util import build
from util import golden, inputs, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec,native,tol_prior,tol_eps", [("fp32", True, 2e-5, 2e-5), ("bf16", True, 1e-2, 1.5e-2),
                                                           ("bf16", False, 1e-4, 1e-2)])
def test_refiner_step(prec, native, tol_prior, tol_eps):
    g = golden("refiner_step.npz")
    with torch.no_grad():
        m, sd = build(H.FacialRefiner, seed=3, precision=prec, max_batch=2, args=())
        m.native_fpg = native
        x = inputs("latents", 1, seed=2).cuda()
        cr_face, cr_latent = inputs("cr_face", 1).cuda(), inputs("cr_latent", 1).cuda()
        out = m(x, torch.tensor([640]), cr_face, cr_latent).sample
        priors, ident = m.condition(cr_face, cr_latent)
        m.denoiser.engine().synchronize()
    worst = max(rel_l2(priors[j], g[f"prior{j}"]) for j in range(5))
    print(f"refiner {prec} native_fpg={native}: worst prior rel-L2 {worst:.3e}, eps rel-L2 {rel_l2(out, g['eps']):.3e}")
    assert worst <= tol_prior
    assert rel_l2(ident, g["identity"]) < 1e-4           # ResNet-50 stays on PyTorch/cuDNN (next row 2)
    assert rel_l2(out, g["eps"]) <= tol_eps
    m.denoiser.invalidate()


def test_parent_load_state_dict_invalidates_engine():
    m, sd = build(H.FacialRefiner, seed=3, precision="bf16", max_batch=2, args=())
    x = inputs("latents", 1, seed=2).cuda()
    cr_face, cr_latent = inputs("cr_face", 1).cuda(), inputs("cr_latent", 1).cuda()
    a = m(x, 5, cr_face, cr_latent).sample.clone()
    sd2 = {k: (v * 0.5 if k.endswith("ending.weight") else v) for k, v in sd.items()}
    m.load_state_dict(sd2)                               # goes through the parent: must still repack
    b = m(x, 5, cr_face, cr_latent).sample
    assert not torch.equal(a, b)
    m.denoiser.invalidate()
