"""Generates tests/golden/cr_forward.npz and cr_layout.json from the UNMODIFIED reference CoarseRestoration
(container only; same discipline as make_golden.py):

    python tests/golden/make_golden_cr.py

1. state_dict layout + seeded default init of hifidiff_b200.CoarseRestoration == the reference's;
2. the CPU oracle (oracle/cr_ref.py) against the reference on identical seeded weights / input, taps included;
3. the reference's output (and two intermediate taps) stored as the fixture.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import cr_ref, ref_shim  # noqa: E402
from hifidiff_b200 import testing  # noqa: E402
import hifidiff_b200 as H  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def rel_l2(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cr_input(batch: int, seed: int = 0):
    g = torch.Generator()
    g.manual_seed(600 + seed)
    return torch.rand((batch, 3, 128, 128), generator=g)


def main():
    torch.set_num_threads(os.cpu_count())
    ref = ref_shim.load()
    torch.manual_seed(0)
    r = ref.CoarseRestoration()
    torch.manual_seed(0)
    m = H.CoarseRestoration()
    rs, ms = r.state_dict(), m.state_dict()
    assert list(rs.keys()) == list(ms.keys()), "key order differs"
    for k in rs:
        assert rs[k].shape == ms[k].shape and rs[k].dtype == ms[k].dtype, k
        assert torch.equal(rs[k], ms[k]), f"default init of {k} differs under the same seed"
    layout = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in rs.items()]
    with open(os.path.join(OUT, "cr_layout.json"), "w") as f:
        json.dump(layout, f)
    print(f"[layout] CoarseRestoration: {len(rs)} entries identical (keys, shapes, dtypes, seeded init)")

    sd = testing.random_state({k: v.shape for k, v in rs.items()}, {k: v.dtype for k, v in rs.items()}, seed=4)
    r.load_state_dict(sd)
    r.eval()
    m.load_state_dict(sd)
    m.eval()
    m.native = False   # the module's PyTorch arithmetic (the library path is CUDA-only)
    x = cr_input(2)
    taps = {}
    with torch.no_grad():
        y_ref = r(x)
        y_orc = cr_ref.cr_forward(sd, x, taps=taps)
        y_mod = m(x)
    e = rel_l2(y_orc, y_ref)
    print(f"[cr] oracle vs reference rel-L2 {e:.2e}; module vs reference {rel_l2(y_mod, y_ref):.2e}; "
          f"out rms {float(y_ref.pow(2).mean().sqrt()):.3f}")
    assert e < 5e-6, e
    assert rel_l2(y_mod, y_ref) < 5e-6
    # the STN really resamples: its output differs from its input, and theta stays near the identity
    assert rel_l2(taps["encoders.0.stn"], taps["encoders.0.nfbs"]) > 1e-3
    with torch.no_grad():
        stn0_ref = r.encoders[0].stn(r.encoders[0].nfbs(r.intro(x)))
    assert rel_l2(taps["encoders.0.stn"], stn0_ref) < 5e-6
    # small fixture: the output and the first 4 channels of the first spatial transformer's output
    np.savez_compressed(os.path.join(OUT, "cr_forward.npz"), y=y_ref.numpy(), stn0=stn0_ref[:, :4].numpy())
    with open(os.path.join(OUT, "cr_summary.json"), "w") as f:
        json.dump({"oracle_vs_ref": e, "module_vs_ref": rel_l2(y_mod, y_ref), "entries": len(rs), "torch": torch.__version__}, f)


if __name__ == "__main__":
    main()
