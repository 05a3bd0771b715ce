"""GPU: FacialRefiner.forward(latents, t, cr_face, cr_latent) against the reference's stored output, with the
FPG prior network and the IDC ResNet-50 on the native sm_100a kernels and on PyTorch
eager; plus the native IDC alone against the PyTorch module on ragged / multi-chunk batches."""
import pytest
import torch

import hifidiff_b200 as H
from oracle import cond_ref

from gpu_util import build
from util import golden, inputs, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec,native,tol_prior,tol_id,tol_eps", [("fp32", True, 2e-5, 2e-5, 2e-5),
                                                                  ("bf16", True, 1e-2, 1e-2, 1.5e-2),
                                                                  ("bf16", False, 1e-4, 1e-4, 1e-2)])
def test_refiner_step(prec, native, tol_prior, tol_id, tol_eps):
    g = golden("refiner_step.npz")
    with torch.no_grad():
        m, sd = build(H.FacialRefiner, seed=3, precision=prec, max_batch=2, args=())
        m.native_fpg = native
        m.native_idc = native
        x = inputs("latents", 1, seed=2).cuda()
        cr_face, cr_latent = inputs("cr_face", 1).cuda(), inputs("cr_latent", 1).cuda()
        out = m(x, torch.tensor([640]), cr_face, cr_latent).sample
        priors, ident = m.condition(cr_face, cr_latent)
        m.denoiser.engine().synchronize()
    worst = max(rel_l2(priors[j], g[f"prior{j}"]) for j in range(5))
    e_id = rel_l2(ident, g["identity"])
    print(f"refiner {prec} native={native}: worst prior rel-L2 {worst:.3e}, identity rel-L2 {e_id:.3e}, "
          f"eps rel-L2 {rel_l2(out, g['eps']):.3e}")
    assert worst <= tol_prior
    assert tuple(ident.shape) == (1, 2048, 1, 1)
    assert e_id <= tol_id
    assert rel_l2(out, g["eps"]) <= tol_eps
    m.denoiser.invalidate()


@pytest.mark.parametrize("prec,batch,tol", [("fp32", 3, 1e-5), ("bf16", 5, 1e-2), ("bf16", 70, 1e-2)])
def test_idc_native_vs_module(prec, batch, tol):
    """hd_idc_forward against the CPU oracle (the restatement pinned to the reference's output) and against the
    PyTorch ResNet-50 with the same state_dict (fp32 cuDNN, TF32 off): ragged batch and a batch that spans two 64-face
    chunks; host input through the staging path gives the same bits."""
    with torch.no_grad():
        m, sd = build(H.FacialRefiner, seed=5, precision=prec, max_batch=4, args=())
        g = torch.Generator().manual_seed(11)
        face = torch.rand((batch, 3, 128, 128), generator=g)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            ref = m.idc.eval()(face.cuda())
        eng = m.denoiser.engine(4)
        eng.load_idc_state(m.idc.state_dict())
        got = eng.idc_forward(face.cuda())
        eng.synchronize()
        out_h = torch.empty_like(got)
        eng.check(eng.lib.hd_idc_forward(eng.handle, face.contiguous().data_ptr(), 128, out_h.data_ptr(), batch, None),
                  "hd_idc_forward")
        eng.synchronize()
        n_or = min(batch, 6)                  # the oracle arm: the first faces and, for two chunks, the last ones
        pick = list(range(n_or)) + ([batch - 2, batch - 1] if batch > 64 else [])
        want = cond_ref.idc_forward({k: v.float().cpu() for k, v in sd.items()}, face[pick], "idc.")
    e = rel_l2(got, ref)
    worst = max(rel_l2(got[i], ref[i]) for i in range(batch))
    worst_or = max(rel_l2(got[i].cpu(), want[k]) for k, i in enumerate(pick))
    print(f"idc {prec} B={batch}: rel-L2 {e:.3e} (worst face {worst:.3e}); vs the CPU oracle on faces {pick}: worst {worst_or:.3e}")
    assert torch.isfinite(got).all()
    assert worst <= tol
    assert worst_or <= tol
    assert torch.equal(out_h, got)
    with pytest.raises(RuntimeError):
        eng.idc_forward(torch.rand((1, 3, 64, 64)).cuda())   # image size must be 8 x latent size
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
