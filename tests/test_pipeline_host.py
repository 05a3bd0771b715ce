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


def test_to_vae_range_clamps_like_the_reference():
    """train_refiner.py:56-61 clamps to [0, 1] before scaling; the CR output routinely leaves that range."""
    from oracle import pipeline_ref
    x = torch.tensor([-0.5, 0.0, 0.25, 1.0, 1.7])
    assert torch.equal(H.to_vae_range(x), torch.tensor([-1.0, -1.0, -0.5, 1.0, 1.0]))
    y = torch.randn(2, 3, 16, 16) * 0.8 + 0.5          # a third of the values fall outside [0, 1]
    assert float(y.min()) < 0 and float(y.max()) > 1
    assert torch.equal(H.to_vae_range(y), pipeline_ref.to_vae_range(y))
    assert torch.equal(H.from_vae_range(y * 4 - 2), pipeline_ref.from_vae_range(y * 4 - 2))
    z = H.encode_latent(_PoolVAE(), y.repeat(1, 1, 8, 8), 0.18215, 128)
    assert torch.equal(z, pipeline_ref.encode_latent(_PoolVAE(), y.repeat(1, 1, 8, 8), 0.18215, 128))


def test_pipeline_ref_matches_reference_source():
    """Pins oracle/pipeline_ref.py: runs the reference's own to_vae_range / from_vae_range / encode_latent bodies,
    cut out of train_refiner.py with ast (the script itself executes argparse and dataset code at import)."""
    import ast
    import os
    import types
    from oracle import pipeline_ref
    src_path = "/root/reference/train_refiner.py"
    if not os.path.exists(src_path):
        pytest.skip("reference tree not present (GPU box)")
    tree = ast.parse(open(src_path).read())
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("to_vae_range", "from_vae_range", "encode_latent")]
    assert len(wanted) == 3
    ns = {"torch": torch, "F": F, "args": types.SimpleNamespace(image_res=128)}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), src_path, "exec"), ns)
    x = torch.randn(2, 3, 64, 64) * 0.8 + 0.5
    assert torch.equal(ns["to_vae_range"](x), pipeline_ref.to_vae_range(x))
    assert torch.equal(ns["from_vae_range"](x * 3 - 1), pipeline_ref.from_vae_range(x * 3 - 1))
    assert torch.equal(ns["encode_latent"](_PoolVAE(), x, 0.18215), pipeline_ref.encode_latent(_PoolVAE(), x, 0.18215, 128))


def test_initial_noise_is_keyed_by_global_face_index():
    """Shards of one set of faces (same seed, their own first_face) start every face from the same x_T."""
    full = H.initial_noise(6, 16, seed=5, first_face=10)
    assert tuple(full.shape) == (6, 4, 16, 16)
    assert torch.equal(full[:2], H.initial_noise(2, 16, seed=5, first_face=10))
    assert torch.equal(full[2:], H.initial_noise(4, 16, seed=5, first_face=12))
    assert not torch.equal(full, H.initial_noise(6, 16, seed=6, first_face=10))
    assert abs(float(full.std()) - 1.0) < 0.05


def test_condition_cache_key_is_object_identity_not_address():
    """The stale-condition hazard: a fresh tensor at a recycled address with _version 0 must NOT hit the cache."""
    from hifidiff_b200.conditioning import _same_tensors, _tensor_key
    a, b = torch.rand(2, 3), torch.rand(2, 3)
    key = _tensor_key((a, b))
    assert _same_tensors(key, (a, b))
    a2 = a.clone()                       # equal contents and shape, another object
    assert not _same_tensors(key, (a2, b))
    view = torch.from_numpy(a.numpy())   # another tensor object over the same address, shape and version 0
    assert view.data_ptr() == a.data_ptr() and view._version == a._version == 0
    assert not _same_tensors(key, (view, b))
    a.add_(1.0)                          # in-place change bumps the version
    assert not _same_tensors(key, (a, b))
    assert not _same_tensors(None, (a, b))


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
