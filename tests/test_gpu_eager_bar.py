"""GPU: the same-box PyTorch-eager bar (last sentence): the oracle's restatement of the reference's
FusedDenoiser forward, run as ordinary PyTorch eager ops ON the B200 (fp32 with TF32 off, and bf16 autocast),
timed next to this library's denoise step at the benchmark batch.  The numbers are printed (run with -s) and
stored in gpurun_out/eager_bar.json; the assertion is only that the native step is not slower than eager fp32."""
import json
import os

import pytest
import torch

import hifidiff_b200 as H
from hifidiff_b200 import testing
from oracle import denoiser_ref

from gpu_util import build
from util import inputs, rel_l2

pytestmark = pytest.mark.gpu

B = 256


def _time(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def test_eager_bar_batch256():
    m, sd = build(H.FusedDenoiser, seed=2, precision="bf16", max_batch=B, max_steps=8)
    sd_dev = {k: v.cuda() for k, v in sd.items()}
    x = inputs("latents", B, seed=5).cuda()
    priors, ident = testing.synthetic_condition(B, 16, seed=5)
    pc, ic = [p.cuda() for p in priors], ident.cuda()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    t = torch.full((B,), 500.0, device="cuda")
    with torch.no_grad():
        ref = denoiser_ref.fused_denoiser_forward(sd_dev, x, t, pc, ic)
        ms_fp32 = _time(lambda: denoiser_ref.fused_denoiser_forward(sd_dev, x, t, pc, ic), 5)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out_ac = denoiser_ref.fused_denoiser_forward(sd_dev, x, t, pc, ic)
            ms_bf16 = _time(lambda: denoiser_ref.fused_denoiser_forward(sd_dev, x, t, pc, ic), 5)
        out = m(x, 500, pc, ic).sample
        ms_native = _time(lambda: m(x, 500, pc, ic), 20)
    m.engine().synchronize()
    res = {"batch": B, "eager_fp32_ms_per_step": ms_fp32, "eager_bf16_autocast_ms_per_step": ms_bf16,
           "native_module_call_ms_per_step": ms_native,
           "native_vs_eager_fp32_rel_l2": rel_l2(out, ref), "eager_autocast_vs_eager_fp32_rel_l2": rel_l2(out_ac.float(), ref),
           "note": "module-API call (includes t-table lookup and launch without graph replay); the sampler's graph-replayed step is in bench.py"}
    print("eager bar:", json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/eager_bar.json", "w") as f:
        json.dump(res, f, indent=1)
    assert res["native_vs_eager_fp32_rel_l2"] <= 1e-2
    assert ms_native < ms_fp32
    m.invalidate()
