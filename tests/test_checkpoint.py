"""CPU: the on-disk formats either side of the module boundary (SURVEY.md §3.3 / §5 "Checkpoint / resume").

`FacialRefiner(latent_res, idc_ckpt, denoiser_ckpt)` (models/refiner.py:11-30) reads
  * an IDC checkpoint written by `torch.save({"epoch", "model_state_dict", "optimizer_state_dict"})`
    (pretrain_idc.py:138-146), and
  * the accelerate `model.safetensors` of a pre-trained *unconditional* Denoiser (pretrain_denoiser.py:209-210),
    loaded into BOTH the FusedDenoiser and the FPG encoder with strict=False (refiner.py:22-25), after which the
    denoiser's intro / encoders are frozen (refiner.py:27-30).
The same files must build the same module here; where the reference tree is present the result is compared with the
reference's own constructor, entry by entry.
"""
import os

import pytest
import torch
from safetensors.torch import load_file, save_file

import hifidiff_b200 as H
from oracle import ref_shim

from util import state_for


@pytest.fixture(scope="module")
def ckpts(tmp_path_factory):
    d = tmp_path_factory.mktemp("ckpt")
    with torch.device("meta"):
        den = H.Denoiser(16)
        idc = H.ResNet50()
    sd_den = state_for(den, seed=21)
    sd_idc = state_for(idc, seed=22)
    den_path, idc_path = os.path.join(d, "model.safetensors"), os.path.join(d, "idc_epoch10.pt")
    save_file({k: v.contiguous() for k, v in sd_den.items()}, den_path)
    torch.save({"epoch": 10, "model_state_dict": sd_idc, "optimizer_state_dict": {}}, idc_path)
    return den_path, idc_path, sd_den, sd_idc


def test_safetensors_round_trip_is_bit_exact(ckpts):
    den_path, _, sd_den, _ = ckpts
    back = load_file(den_path)
    assert list(back.keys()) == sorted(sd_den.keys()) or set(back.keys()) == set(sd_den.keys())
    for k, v in sd_den.items():
        assert back[k].dtype == v.dtype and torch.equal(back[k], v), k
    # a strict load into the module the file was written from
    m = H.Denoiser(16)
    res = m.load_state_dict(back, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd_den[k]), k


def test_refiner_constructor_double_load(ckpts):
    den_path, idc_path, sd_den, sd_idc = ckpts
    torch.manual_seed(0)
    m = H.FacialRefiner(latent_res=16, idc_ckpt=idc_path, denoiser_ckpt=den_path)
    sd = m.state_dict()
    # denoiser: every checkpoint key lands (the Denoiser layout is a strict subset of the FusedDenoiser's) ...
    for k, v in sd_den.items():
        assert torch.equal(sd["denoiser." + k], v), k
    # ... and the FusedDenoiser-only tensors keep their constructor init
    extra = [k for k in sd if k.startswith("denoiser.") and k[len("denoiser."):] not in sd_den]
    assert extra and all(k.startswith(("denoiser.hcas.", "denoiser.idc_conv.")) for k in extra)
    # FPG: intro / encoders / downs come from the same file, `convs.*` (FPG-only) stays at init
    shared = [k for k in sd if k.startswith("fpg.") and k[len("fpg."):] in sd_den]
    only = [k for k in sd if k.startswith("fpg.") and k[len("fpg."):] not in sd_den]
    assert len(shared) == 298 and len(only) == 5 and all(k.startswith("fpg.convs.") for k in only)
    for k in shared:
        assert torch.equal(sd[k], sd_den[k[len("fpg."):]]), k
    for k, v in sd_idc.items():
        assert torch.equal(sd["idc." + k], v), k
    # refiner.py:27-30: the denoiser's encoder side is frozen, nothing else
    for name, p in m.denoiser.named_parameters():
        assert p.requires_grad == (not (name.startswith("intro") or name.startswith("encoders"))), name
    assert all(p.requires_grad for p in m.fpg.parameters())
    assert not m.idc.training


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
def test_refiner_constructor_matches_reference(ckpts):
    """Same seed, same two files, the reference's own constructor: identical state_dict, identical freezing."""
    den_path, idc_path, _, _ = ckpts
    ref = ref_shim.load()
    torch.manual_seed(0)
    want = ref.FacialRefiner(latent_res=16, idc_ckpt=idc_path, denoiser_ckpt=den_path)
    torch.manual_seed(0)
    got = H.FacialRefiner(latent_res=16, idc_ckpt=idc_path, denoiser_ckpt=den_path)
    sw, sg = want.state_dict(), got.state_dict()
    assert list(sw.keys()) == list(sg.keys())
    for k in sw:
        assert sw[k].dtype == sg[k].dtype and torch.equal(sw[k], sg[k]), k
    assert ({n: p.requires_grad for n, p in want.named_parameters()}
            == {n: p.requires_grad for n, p in got.named_parameters()})
    # eval-path load (test_refiner.py:162-164): one strict load of the whole refiner file
    full = os.path.join(os.path.dirname(den_path), "refiner.safetensors")
    save_file({k: v.contiguous() for k, v in sw.items()}, full)
    again = H.FacialRefiner(latent_res=16)
    res = again.load_state_dict(load_file(full))
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in again.state_dict().items():
        assert torch.equal(v, sw[k]), k


def test_submodule_loads_drop_the_packed_copies(ckpts):
    """The engine keeps packed copies of the FPG / IDC weights and the refiner caches the hoisted condition: a load
    aimed at a sub-module (`refiner.fpg.load_state_dict`, `refiner.idc.load_state_dict` — how refiner.py:17,24-25
    themselves load) or a `.to()` must drop both, exactly as a load of the whole refiner does."""
    import copy
    _, _, _, sd_idc = ckpts
    m = H.FacialRefiner(latent_res=16)
    calls = []
    orig = m.denoiser.invalidate
    m.denoiser.invalidate = lambda: (calls.append(1), orig())[1]

    def armed():
        m._sync_weight_epoch()
        m._cond, m._cond_src = ("stale",), ("stale",)
        calls.clear()

    armed()
    m._sync_weight_epoch()          # nothing loaded since: the cache survives
    assert not calls and m._cond == ("stale",)
    m.idc.load_state_dict(sd_idc)
    m._sync_weight_epoch()          # what `condition` does first
    assert calls and m._cond is None and m._cond_src is None
    armed()
    m.fpg.load_state_dict(m.fpg.state_dict())
    m._sync_weight_epoch()
    assert calls and m._cond is None and m._cond_src is None
    armed()
    m.load_state_dict(m.state_dict())
    assert m._cond is None and m._cond_src is None
    armed()
    m.to(torch.float32)
    assert calls and m._cond is None and m._cond_src is None
    # the hooks hold no reference to the parent: a deep copy tracks its own sub-modules
    m.denoiser.invalidate = orig
    armed()
    m2 = copy.deepcopy(m)
    m2.idc.load_state_dict(sd_idc)
    m2._sync_weight_epoch()
    m._sync_weight_epoch()
    assert m2._cond is None and m._cond == ("stale",)
