"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from hifidiff_b200 import testing

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LEVELS = [(128, 16), (256, 8), (512, 4), (1024, 2), (2048, 1)]
TRAJ_EPS_GAIN = 0.15


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def psnr(a, ref):
    """PSNR of `a` against `ref` with data range = max - min of `ref`."""
    a, ref = torch.as_tensor(a).double().cpu(), torch.as_tensor(ref).double().cpu()
    mse = float((a - ref).pow(2).mean())
    rng = float(ref.max() - ref.min())
    return float("inf") if mse == 0 else 10.0 * np.log10(rng * rng / mse)


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def inputs(kind, batch, seed=0):
    """Same deterministic inputs as tests/golden/make_golden.py."""
    if kind == "latents":
        return torch.randn((batch, 4, 16, 16), generator=gen(100 + seed))
    if kind == "cr_face":
        return torch.rand((batch, 3, 128, 128), generator=gen(200 + seed))
    if kind == "cr_latent":
        return torch.randn((batch, 4, 16, 16), generator=gen(300 + seed))
    if kind == "ln_face":   # CoarseRestoration input, same as tests/golden/make_golden_cr.py::cr_input
        return torch.rand((batch, 3, 128, 128), generator=gen(600 + seed))
    raise KeyError(kind)


def state_for(module_or_shapes, seed, eps_gain=1.0):
    sd = module_or_shapes.state_dict() if hasattr(module_or_shapes, "state_dict") else module_or_shapes
    return testing.random_state({k: v.shape for k, v in sd.items()}, {k: v.dtype for k, v in sd.items()}, seed, eps_gain)


def block_shapes(c, time_dim=512):
    """state_dict shapes of one ConditionalNAFBlock."""
    f = torch.float32
    sh = {"beta": (1, c, 1, 1), "gamma": (1, c, 1, 1), "mlp.1.weight": (4 * c, time_dim // 2), "mlp.1.bias": (4 * c,),
          "conv1.weight": (2 * c, c, 1, 1), "conv1.bias": (2 * c,), "conv2.weight": (2 * c, 1, 3, 3),
          "conv2.bias": (2 * c,), "conv3.weight": (c, c, 1, 1), "conv3.bias": (c,), "sca.1.weight": (c, c, 1, 1),
          "sca.1.bias": (c,), "conv4.weight": (2 * c, c, 1, 1), "conv4.bias": (2 * c,), "conv5.weight": (c, c, 1, 1),
          "conv5.bias": (c,), "norm1.weight": (c,), "norm1.bias": (c,), "norm2.weight": (c,), "norm2.bias": (c,)}
    return {k: torch.empty(v, dtype=f) for k, v in sh.items()}


def hca_shapes(d):
    f = torch.float32
    sh = {"channel_mlp.0.weight": (d, d), "channel_mlp.0.bias": (d,), "channel_mlp.2.weight": (d, d),
          "channel_mlp.2.bias": (d,), "spatial_mlp.0.weight": (d // 2, d, 1, 1), "spatial_mlp.0.bias": (d // 2,)}
    out = {k: torch.empty(v, dtype=f) for k, v in sh.items()}

    def bn(p, n):
        out[p + "weight"] = torch.empty(n)
        out[p + "bias"] = torch.empty(n)
        out[p + "running_mean"] = torch.empty(n)
        out[p + "running_var"] = torch.empty(n)
        out[p + "num_batches_tracked"] = torch.empty((), dtype=torch.int64)
    bn("spatial_mlp.1.", d // 2)
    out["spatial_mlp.3.weight"] = torch.empty((1, d // 2, 1, 1))
    out["spatial_mlp.3.bias"] = torch.empty(1)
    bn("spatial_mlp.4.", 1)
    out["fused_mlp.0.weight"] = torch.empty((d, d, 3, 3))
    out["fused_mlp.0.bias"] = torch.empty(d)
    bn("fused_mlp.1.", d)
    return out


# ---- diffusers' scheduler known-answer loops on the update kernel's own tensor shape -----------------------------
# The fixtures of huggingface/diffusers tests/schedulers (see tests/test_schedulers.py) are (4,3,8,8) = 768 values; the
# update kernel works on whole faces of 4*16*16 = 1024 values.  The update is elementwise, so the 768 values ride in the
# head of one (1,4,16,16) face and the 256 padding values stay exactly 0 (x = 0, eps = 0 * t / (t + 1) = 0).
KAT_CFG = dict(num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear", clip_sample=True)
KAT_DDIM = (172.0067, 0.223967)    # DDIMSchedulerTest.test_full_loop_no_noise: sum |x|, mean |x|
KAT_DDPM = (258.9606, 0.3372)      # DDPMSchedulerTest.test_full_loop_no_noise


def kat_sample():
    b, c, h, w = 4, 3, 8, 8
    n = b * c * h * w
    return (torch.arange(n).reshape(c, h, w, b) / n).permute(3, 0, 1, 2)


def _as_face(v):
    out = torch.zeros(1024, dtype=torch.float32)
    out[:768] = v.reshape(-1)
    return out.reshape(1, 4, 16, 16)


def kat_ddim_loop(step, timesteps):
    """step(i, x_face, eps_face) -> x_face for step index i; returns (sum |x|, mean |x|, padding max |x|)."""
    x = _as_face(kat_sample())
    for i, t in enumerate(timesteps):
        x = step(i, x, x * t / (t + 1))
    v = x.reshape(-1)
    return float(v[:768].abs().sum()), float(v[:768].abs().mean()), float(v[768:].abs().max())


def kat_ddpm_loop(step):
    """step(i, x_face, eps_face, z_face) for i = 0..999 (t = 999 - i); noise as upstream: torch.manual_seed(0),
    one (4,3,8,8) draw per step, t = 0 included."""
    g = torch.manual_seed(0)
    x = _as_face(kat_sample())
    for i, t in enumerate(reversed(range(1000))):
        z = _as_face(torch.randn((4, 3, 8, 8), generator=g))
        x = step(i, x, x * t / (t + 1), z)
    v = x.reshape(-1)
    return float(v[:768].abs().sum()), float(v[:768].abs().mean()), float(v[768:].abs().max())
