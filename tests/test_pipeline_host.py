"""CPU: host-side pieces of the pixel-to-pixel pipeline mirror (hifidiff_b200/pipeline.py; reference
train_refiner.py:56-83): range conversion, the bicubic-resize shortcut, and the no-CPU-path rule."""
import pytest
import torch
import torch.nn.functional as F

import hifidiff_b200 as H


class _Dist:
    def __init__(self, z):
        self.z = z
        self.latent_dist = self

    def sample(self):
        return self.z


class _PoolVAE:
    def encode(self, x):
        return _Dist(F.avg_pool2d(x, 8)[:, :1].repeat(1, 4, 1, 1))


def test_vae_range_round_trip():
    x = torch.rand(2, 3, 8, 8)
    assert torch.allclose(H.from_vae_range(H.to_vae_range(x)), x, atol=1e-7)
    assert float(H.from_vae_range(torch.tensor([-3.0, 3.0])).min()) == 0.0
    assert float(H.from_vae_range(torch.tensor([-3.0, 3.0])).max()) == 1.0


def test_encode_latent_resize_shortcut_is_exact():
    """At equal size the reference's bicubic F.interpolate (train_refiner.py:74-79) is the identity, so skipping it
    changes nothing; at another size the resize runs."""
    vae = _PoolVAE()
    x = torch.rand(2, 3, 128, 128)
    same = F.interpolate(x, size=(128, 128), mode="bicubic", align_corners=False)
    assert torch.equal(same, x)
    z = H.encode_latent(vae, x, 0.18215, 128)
    want = vae.encode(H.to_vae_range(same)).latent_dist.sample() * 0.18215
    assert torch.equal(z, want) and tuple(z.shape) == (2, 4, 16, 16)
    small = torch.rand(1, 3, 64, 64)
    z2 = H.encode_latent(vae, small, 0.18215, 128)
    assert tuple(z2.shape) == (1, 4, 16, 16)


def test_pipeline_has_no_cpu_path():
    with pytest.raises(RuntimeError):
        H.ddim_sample_images(torch.rand(1, 3, 128, 128), None, None, None, None)
