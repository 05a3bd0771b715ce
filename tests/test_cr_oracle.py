"""CPU: CoarseRestoration — the oracle restatement and the module boundary against the
fixtures generated from the unmodified reference (tests/golden/make_golden_cr.py).  The sm_100a kernels for this
stage are the next step; nothing here claims GPU parity."""
import json
import os

import pytest
import torch

import hifidiff_b200 as H
from oracle import cr_ref

from util import GOLDEN, golden, inputs, rel_l2, state_for

TOL = 2e-6


@pytest.fixture(scope="module")
def cr_state():
    with torch.device("meta"):
        m = H.CoarseRestoration()
    return state_for(m, seed=4)


def test_cr_oracle_matches_reference_fixture(cr_state):
    g = golden("cr_forward.npz")
    taps = {}
    with torch.no_grad():
        y = cr_ref.cr_forward(cr_state, inputs("ln_face", 2), taps=taps)
    assert rel_l2(y, g["y"]) < TOL
    assert rel_l2(taps["encoders.0.stn"][:, :4], g["stn0"]) < TOL
    # the spatial transformer is exercised: it moves pixels, and keeps most of the image in view
    assert rel_l2(taps["encoders.0.stn"], taps["encoders.0.nfbs"]) > 1e-3
    assert float((taps["encoders.0.stn"].abs().sum(1) > 0).float().mean()) > 0.8


def test_cr_module_layout_and_forward(cr_state):
    with open(os.path.join(GOLDEN, "cr_layout.json")) as f:
        layout = json.load(f)
    with torch.device("meta"):
        meta = H.CoarseRestoration()
    mine = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in meta.state_dict().items()]
    assert mine == layout and len(layout) == 664
    m = H.CoarseRestoration()
    m.load_state_dict(cr_state)          # strict: every key of the reference layout, nothing else
    m.eval()
    with pytest.raises(RuntimeError), torch.no_grad():
        m(inputs("ln_face", 1))          # the library path has no CPU fallback
    m.native = False                     # explicit opt-out: the module's PyTorch arithmetic
    with torch.no_grad():
        y = m(inputs("ln_face", 2))
    g = golden("cr_forward.npz")
    assert tuple(y.shape) == (2, 3, 128, 128)
    assert rel_l2(y, g["y"]) < TOL


def test_cr_stn_default_init_is_identity():
    """Reference init (stn.py:36-40): zero weight + identity bias in the last FC, so a fresh STN is a no-op resample."""
    torch.manual_seed(0)
    m = H.CoarseRestoration().eval()
    m.native = False
    x = torch.randn(1, 32, 128, 128)
    with torch.no_grad():
        y = m.encoders[0].stn(x)
    assert rel_l2(y, x) < 1e-5


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not mounted")
def test_cr_oracle_against_live_reference_stage():
    """One NAF_STN_Block of the unmodified reference (different seed, batch 3, the 16x16 stage with its 5x5 / 3x3
    localisation convs) against the oracle's restatement."""
    from oracle import ref_shim
    ref_shim.load()
    from models.cr.model import NAF_STN_Block  # type: ignore  # noqa: E402  (path set by ref_shim)
    blk = NAF_STN_Block(256, 16, num_naf=2, sampling="up").eval()
    sd = state_for(blk, seed=123)
    blk.load_state_dict(sd)
    x = torch.randn(3, 256, 16, 16, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        want = blk(x)
        got = cr_ref.naf_stn_block(sd, "", x, 2, "up")
    assert tuple(got.shape) == (3, 128, 32, 32)
    assert rel_l2(got, want) < TOL
