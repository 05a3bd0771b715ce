"""Reverse-diffusion sampling on the sm_100a engine (latent in, latent out).

`sample` is the loop of the reference's `ddim_sample` (train_refiner.py:86-125; copies at
test_refiner.py:58-95 and pretrain_denoiser.py:76-120) between "x_T drawn" and "VAE decode":
    for t in scheduler.timesteps:  eps = model(x, t, ...).sample ; x = scheduler.step(eps, t, x).prev_sample
executed as ONE library call (`hd_sample`): per-step kernels replayed from a CUDA graph, the
time-modulation table pre-computed for the whole schedule, x_{t-1} update + Philox noise fused in
one elementwise kernel, no host round trip per step.

`sample_sharded` partitions a set of faces over the ranks of a torch.distributed process group
(one process per GPU): each face's trajectory is independent, so ranks never talk during
sampling; the final latents are gathered once (all_gather over NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Sequence, Tuple

import torch

from . import _lib
from .conditioning import FacialRefiner
from .modules import Denoiser, FusedDenoiser, _stream_ptr
from .schedulers import DDIMScheduler, DDPMScheduler, _SchedulerBase


def _coef_array(coefs) -> "C.Array":
    arr = (_lib.HdStepCoef * len(coefs))()
    for i, c in enumerate(coefs):
        arr[i] = _lib.HdStepCoef(c.timestep, c.sqrt_beta_prod, c.sqrt_alpha_prod, c.clip, c.k_x0, c.k_eps, c.k_x,
                                 c.k_noise)
    return arr


@torch.no_grad()
def sample(model, x_T: torch.Tensor, scheduler: _SchedulerBase, num_inference_steps: int = 50, *,
           eta: float = 0.0, facial_priors: Optional[Sequence[torch.Tensor]] = None,
           identity_embedding: Optional[torch.Tensor] = None, cr_face: Optional[torch.Tensor] = None,
           cr_latent: Optional[torch.Tensor] = None, seed: int = 0, first_face: int = 0,
           noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Runs the whole reverse process for a batch of faces and returns x_0 (B,4,S,S) fp32.

    model      Denoiser | FusedDenoiser (+ facial_priors, identity_embedding) | FacialRefiner (+ cr_face, cr_latent)
    x_T        initial noise, passed explicitly (the reference draws it from the global RNG, :101-104)
    scheduler  hifidiff_b200.schedulers.DDIMScheduler / DDPMScheduler
    seed, first_face   Philox key / global index of x_T[0]; noise for face i depends only on
                       (seed, first_face + i, step), so any sharding reproduces the same trajectories
    noise      optional explicit z, (steps, B, 4*S*S), overrides Philox (parity tests)
    """
    if x_T.device.type != "cuda":
        raise RuntimeError("hifidiff_b200 has no CPU path: x_T must be a CUDA tensor")
    if x_T.shape[0] == 0:  # an empty shard (fewer faces than ranks in `sample_sharded`): nothing to launch
        return (model.denoiser if isinstance(model, FacialRefiner) else model)._check_latents(x_T).clone()
    if isinstance(model, FacialRefiner):
        facial_priors, identity_embedding = model.condition(cr_face, cr_latent)
        model = model.denoiser
    scheduler.set_timesteps(num_inference_steps, device="cpu")
    if isinstance(scheduler, DDIMScheduler):
        coefs = scheduler.step_coefficients(eta=eta)
    else:
        coefs = scheduler.step_coefficients()
    b = x_T.shape[0]
    if len(coefs) > model.max_steps:
        model.configure(max_steps=len(coefs))
    if isinstance(model, FusedDenoiser):
        model._ensure_condition(facial_priors, identity_embedding, b)
    x = model._check_latents(x_T).clone()
    eng = model.engine(b)
    nz = None
    if noise is not None:
        nz = noise.to(device=x.device, dtype=torch.float32).contiguous()
        if nz.numel() != len(coefs) * x.numel():
            raise ValueError("noise must be (steps, B, 4*S*S)")
    arr = _coef_array(coefs)
    with torch.cuda.device(x.device):
        eng.check(eng.lib.hd_sample(eng.handle, x.data_ptr(), arr, len(coefs), C.c_uint64(seed), C.c_int64(first_face),
                                    b, nz.data_ptr() if nz is not None else None, _stream_ptr(x.device)), "hd_sample")
    return x


def ddim_sample(model, x_T, scheduler: DDIMScheduler, num_inference_steps: int = 50, **kw) -> torch.Tensor:
    """`ddim_sample` restricted to latent-in/latent-out (eta = 0.0 as at train_refiner.py:120)."""
    return sample(model, x_T, scheduler, num_inference_steps, eta=0.0, **kw)


def ddpm_sample(model, x_T, scheduler: DDPMScheduler, num_inference_steps: int = 1000, **kw) -> torch.Tensor:
    """Ancestral sampling (BASELINE.json config 3); the reference itself never samples with DDPM."""
    return sample(model, x_T, scheduler, num_inference_steps, **kw)


def shard_bounds(n_faces: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of global face indices owned by `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_faces, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sample_sharded(run_local: Callable[[int, int], torch.Tensor], n_faces: int, *, group=None,
                   gather: bool = True) -> torch.Tensor:
    """Sample `n_faces` faces over the ranks of a process group.

    run_local(lo, hi) -> latents of faces [lo, hi) on this rank.  No collective runs during
    sampling; with gather=True every rank returns all n_faces latents, in global order, after one
    all_gather of the final latents (16 MB at 4096 faces)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return run_local(0, n_faces)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n_faces, world, rank)
    local = run_local(lo, hi)
    if not gather:
        return local
    counts = [shard_bounds(n_faces, world, r) for r in range(world)]
    cap = max(h - l for l, h in counts)
    padded = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: hi - lo] = local
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    return torch.cat([bufs[r][: h - l] for r, (l, h) in enumerate(counts)], dim=0)
