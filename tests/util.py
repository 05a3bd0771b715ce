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
