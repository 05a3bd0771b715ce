"""Host-side schedulers: the subset of the diffusers==0.32.2 API the reference calls, producing
the per-step coefficient table the CUDA sampler consumes (include/hifidiff_b200.h: hd_step_coef).

Reference call sites mirrored (constructor kwargs, `set_timesteps`, `.timesteps`, `add_noise`):
  train_refiner.py:337-348,109 ; pretrain_denoiser.py:261-272,99 ; test_refiner.py:166-171,85.
`scheduler.step(...)` itself is not executed on the host: its arithmetic is the elementwise
sm_100a kernel `sampler_update_kernel`, and `step_coefficients()` evaluates the scalar part of
diffusers' `step` in fp32 torch ops in the same order (so the scalars round the same way).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch


@dataclass
class StepCoef:
    timestep: float
    sqrt_beta_prod: float
    sqrt_alpha_prod: float
    clip: float
    k_x0: float
    k_eps: float
    k_x: float
    k_noise: float


class _SchedulerBase:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 2e-2,
                 beta_schedule: str = "linear", prediction_type: str = "epsilon", clip_sample: bool = True,
                 clip_sample_range: float = 1.0, **unused):
        # diffusers kwargs this restatement implements only at their default value: anything else must not be
        # silently ignored (it would sample with leading / offset-0 / fixed_small behaviour regardless)
        supported_defaults = {"trained_betas": None, "set_alpha_to_one": True, "steps_offset": 0,
                              "timestep_spacing": "leading", "thresholding": False, "rescale_betas_zero_snr": False,
                              "variance_type": "fixed_small", "dynamic_thresholding_ratio": 0.995,
                              "sample_max_value": 1.0}
        for k, v in unused.items():
            if k not in supported_defaults:
                raise TypeError(f"unexpected scheduler argument '{k}'")
            if v != supported_defaults[k]:
                raise NotImplementedError(f"{k}={v!r}: only the diffusers default {supported_defaults[k]!r} is implemented")
        if prediction_type != "epsilon":
            raise NotImplementedError("only prediction_type='epsilon' (the reference's setting)")
        if beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                        dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(f"beta_schedule={beta_schedule}")
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.one = torch.tensor(1.0)
        self.num_train_timesteps = num_train_timesteps
        self.clip_sample = clip_sample
        self.clip_sample_range = clip_sample_range
        self.init_noise_sigma = 1.0
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps cannot exceed num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts)
        if device is not None:
            self.timesteps = self.timesteps.to(device)

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        a = ac[timesteps] ** 0.5
        s = (1 - ac[timesteps]) ** 0.5
        while a.dim() < original_samples.dim():
            a, s = a.unsqueeze(-1), s.unsqueeze(-1)
        return a * original_samples + s * noise

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _prev(self, t: int) -> int:
        n = self.num_inference_steps or self.num_train_timesteps
        return t - self.num_train_timesteps // n

    def step_coefficients(self, **kw) -> List[StepCoef]:
        raise NotImplementedError


class DDIMScheduler(_SchedulerBase):
    """DDIMScheduler(set_alpha_to_one=True, timestep_spacing='leading', steps_offset=0)."""

    def step_coefficients(self, eta: float = 0.0) -> List[StepCoef]:
        out = []
        for t in self.timesteps.tolist():
            prev_t = self._prev(t)
            a_t = self.alphas_cumprod[t]
            a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
            beta_t = 1 - a_t
            variance = ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)
            std = eta * variance ** 0.5
            out.append(StepCoef(float(t), float(beta_t ** 0.5), float(a_t ** 0.5),
                                float(self.clip_sample_range) if self.clip_sample else 0.0,
                                float(a_p ** 0.5), float((1 - a_p - std ** 2) ** 0.5), 0.0, float(std)))
        return out


class DDPMScheduler(_SchedulerBase):
    """DDPMScheduler(variance_type='fixed_small', timestep_spacing='leading')."""

    def step_coefficients(self) -> List[StepCoef]:
        out = []
        for t in self.timesteps.tolist():
            prev_t = self._prev(t)
            a_t = self.alphas_cumprod[t]
            a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
            beta_prod_t = 1 - a_t
            beta_prod_p = 1 - a_p
            cur_alpha = a_t / a_p
            cur_beta = 1 - cur_alpha
            c_x0 = (a_p ** 0.5 * cur_beta) / beta_prod_t
            c_xt = cur_alpha ** 0.5 * beta_prod_p / beta_prod_t
            sigma = 0.0
            if t > 0:
                var = torch.clamp(beta_prod_p / beta_prod_t * cur_beta, min=1e-20)
                sigma = float(var ** 0.5)
            out.append(StepCoef(float(t), float(beta_prod_t ** 0.5), float(a_t ** 0.5),
                                float(self.clip_sample_range) if self.clip_sample else 0.0,
                                float(c_x0), 0.0, float(c_xt), sigma))
        return out
