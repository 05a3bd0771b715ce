"""CPU: Philox4x32-10 known-answer vectors (Random123 kat_vectors) and noise statistics."""
import numpy as np

from oracle import philox


def _one(c, k):
    out = philox.philox4x32_10(*[np.array([v], dtype=np.uint32) for v in c], k[0], k[1])
    return [int(o[0]) for o in out]


def test_known_answer_vectors():
    assert _one((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert _one((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert _one((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_noise_is_standard_normal_and_shard_invariant():
    z = philox.normal_noise(seed=1234, first_face=0, n_faces=64, step_index=5)
    assert z.shape == (64, 1024) and z.dtype == np.float32
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02
    assert np.isfinite(z).all()
    # faces [32,64) drawn as their own shard are bit-identical
    z2 = philox.normal_noise(seed=1234, first_face=32, n_faces=32, step_index=5)
    assert np.array_equal(z[32:], z2)
    assert not np.array_equal(z, philox.normal_noise(seed=1234, first_face=0, n_faces=64, step_index=6))
