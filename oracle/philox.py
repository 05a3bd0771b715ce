"""Counter-based Gaussian noise for the DDPM ancestral step (TEST INFRASTRUCTURE).

Philox4x32-10 (Salmon et al., SC'11 / Random123) + Box-Muller, keyed so that the noise for
(face, step, element) does not depend on how faces are sharded over GPUs.  The CUDA sampler
kernel (hifidiff_b200/csrc/sampler.cuh) implements the identical integer stream; the float
transform agrees to a few ulp (logf/sinf/cosf rounding).

No reference anchor: the reference draws x_T from torch's global RNG (train_refiner.py:101-104)
and never samples with DDPM; this is the design's own noise definition.

  key     = (seed & 0xffffffff, seed >> 32)
  counter = (group, face_global_index, step_index, 0x48494644)   group = element // 4
  r0..r3  -> u = ((r >> 9) + 0.5) * 2**-23  in (0,1)
  z[4g+0], z[4g+1] = sqrt(-2 ln u0) * (cos, sin)(2 pi u1);  z[4g+2], z[4g+3] from (u2, u3)
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
STREAM_TAG = 0x48494644
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs uint32 arrays (broadcastable); returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for r in range(10):
            if r > 0:
                k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
                k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
            p0 = c0.astype(np.uint64) * _M0
            p1 = c2.astype(np.uint64) * _M1
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
    return c0, c1, c2, c3


def _unit(r: np.ndarray) -> np.ndarray:
    return ((r >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)


def normal_noise(seed: int, first_face: int, n_faces: int, step_index: int, elems_per_face: int = 1024) -> np.ndarray:
    """(n_faces, elems_per_face) float32 standard normals for one sampler step."""
    assert elems_per_face % 4 == 0
    groups = np.arange(elems_per_face // 4, dtype=np.uint32)[None, :]
    faces = (np.arange(n_faces, dtype=np.uint64) + np.uint64(first_face)).astype(np.uint32)[:, None]
    r0, r1, r2, r3 = philox4x32_10(groups, faces, np.uint32(step_index), np.uint32(STREAM_TAG),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    two_pi = np.float32(6.283185307179586)
    out = np.empty((n_faces, elems_per_face // 4, 4), dtype=np.float32)
    for j, (ra, rb) in enumerate(((r0, r1), (r2, r3))):
        rad = np.sqrt(np.float32(-2.0) * np.log(_unit(ra)))
        ang = two_pi * _unit(rb)
        out[:, :, 2 * j] = rad * np.cos(ang)
        out[:, :, 2 * j + 1] = rad * np.sin(ang)
    return out.reshape(n_faces, elems_per_face)
