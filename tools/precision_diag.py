"""Per-layer rel-L2 of the CUDA path against the CPU oracle (run on a GPU box)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import hifidiff_b200 as H  # noqa: E402
from hifidiff_b200 import testing  # noqa: E402
from oracle import denoiser_ref  # noqa: E402
from gpu_util import build  # noqa: E402
from test_gpu_denoiser import FUSED_TAPS  # noqa: E402
from util import inputs, rel_l2  # noqa: E402

out = {}
for prec in sys.argv[1:] or ["bf16"]:
    m, sd = build(H.FusedDenoiser, seed=2, precision=prec, max_batch=8)
    priors, ident = testing.synthetic_condition(2, 16, seed=0)
    pc, ic = [p.cuda() for p in priors], ident.cuda()
    for t in (980, 500, 3):
        x = inputs("latents", 2, seed=1)
        o, taps = m.forward_with_taps(x.cuda(), t, FUSED_TAPS, pc, ic)
        m.engine().synchronize()
        ref_taps = {}
        with torch.no_grad():
            ref = denoiser_ref.fused_denoiser_forward(sd, x, t, priors, ident, ref_taps)
        rows = {k: rel_l2(taps[k], ref_taps[k]) for k in FUSED_TAPS}
        rows["eps"] = rel_l2(o.sample, ref)
        out[f"{prec}_t{t}"] = rows
        print(f"--- {prec} t={t}")
        for k, v in rows.items():
            print(f"{k:16s} {v:.3e}  |ref| rms {float(ref_taps[k].pow(2).mean().sqrt()) if k in ref_taps else float(ref.pow(2).mean().sqrt()):.3f}")
    m.invalidate()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "precision_diag.json"), "w"), indent=1)
