"""Drop-in `nn.Module`s for the reference's denoisers, executing on the sm_100a library.

The classes keep the reference's constructor arguments, `forward` signatures, attribute bag
(`.width`, `.dtype`, `.config.in_channels`, `.config.sample_size`), `UNet2DOutput` return type and
— parameter for parameter, in the same registration order — its `state_dict()` layout
(SURVEY.md App. B), so checkpoints written by the reference load unchanged and a
`torch.manual_seed(s)`-then-construct sequence draws the same initial values.

  Denoiser       <- models/denoiser/model.py:32-134
  FusedDenoiser  <- models/denoiser/model.py:137-266
  block params   <- models/denoiser/conditional_naf.py:13-101
  HCA params     <- models/fpg/hca.py:5-23

The sub-modules here are *parameter holders*: they never run.  `forward` hands the parameters
(once, lazily, or again after `load_state_dict` / `invalidate()`) and the input tensors to
`libhifidiff_b200.so` through ctypes.  There is no PyTorch or CPU fallback: tensors that are not
on a CUDA device raise.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence

import torch
from torch import nn

from . import _lib

_ENC = (2, 2, 4, 8)
_DEC = (2, 2, 2, 2)
_MID = 8


class UNet2DOutput:
    """Return wrapper with `.sample` (models/denoiser/model.py:11-13)."""

    def __init__(self, data):
        self.sample = data


class _Bag:
    """Stand-in for the `diffusers.ConfigMixin()` attribute bag (model.py:39-41)."""


class _NoParams(nn.Module):
    """Placeholder occupying a Sequential slot that holds no parameters in the reference
    (SinusoidalPosEmb, SimpleGate, ReLU, Sigmoid, PixelShuffle, AdaptiveAvgPool2d)."""

    def forward(self, *a, **k):  # pragma: no cover - parameter holders never run
        raise RuntimeError("hifidiff_b200 sub-modules are parameter holders; call the top-level module")


class _LayerNorm2dParams(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))


class _NAFBlockParams(nn.Module):
    """Parameters of one ConditionalNAFBlock (time_dim given) or NAFBlock (time_dim None)."""

    def __init__(self, c: int, time_dim: Optional[int]):
        super().__init__()
        if time_dim:
            self.mlp = nn.Sequential(_NoParams(), nn.Linear(time_dim // 2, 4 * c))
        self.conv1 = nn.Conv2d(c, 2 * c, 1)
        self.conv2 = nn.Conv2d(2 * c, 2 * c, 3, padding=1, groups=2 * c)
        self.conv3 = nn.Conv2d(c, c, 1)
        self.sca = nn.Sequential(_NoParams(), nn.Conv2d(c, c, 1))
        self.conv4 = nn.Conv2d(c, 2 * c, 1)
        self.conv5 = nn.Conv2d(c, c, 1)
        self.norm1 = _LayerNorm2dParams(c)
        self.norm2 = _LayerNorm2dParams(c)
        self.beta = nn.Parameter(torch.zeros((1, c, 1, 1)))
        self.gamma = nn.Parameter(torch.zeros((1, c, 1, 1)))


class _HCAParams(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.channel_mlp = nn.Sequential(nn.Linear(dim, dim), _NoParams(), nn.Linear(dim, dim), _NoParams())
        self.spatial_mlp = nn.Sequential(nn.Conv2d(dim, dim // 2, 1), nn.BatchNorm2d(dim // 2), _NoParams(),
                                         nn.Conv2d(dim // 2, 1, 1), nn.BatchNorm2d(1), _NoParams())
        self.fused_mlp = nn.Sequential(nn.Conv2d(dim, dim, 3, 1, 1), nn.BatchNorm2d(dim), _NoParams())


def _blocks(n: int, c: int, time_dim: Optional[int]) -> nn.Sequential:
    return nn.Sequential(*[_NAFBlockParams(c, time_dim) for _ in range(n)])


class _Engine:
    """Owns one hd_handle for a module; re-created when the module's weights or device change."""

    def __init__(self, model_kind: int, latent_size: int, precision: int, device: torch.device, max_batch: int,
                 max_steps: int, use_graph: bool):
        lib = _lib.load()
        self.lib = lib
        self.handle = C.c_void_p()
        self.fpg_loaded = False
        self.idc_loaded = False
        self.cr_loaded = False
        self.max_batch = max_batch
        self.max_steps = max_steps
        cfg = _lib.HdConfig(C.sizeof(_lib.HdConfig), model_kind, precision, latent_size,
                            device.index if device.index is not None else torch.cuda.current_device(),
                            max_batch, max_steps, 1 if use_graph else 0)
        st = lib.hd_create(C.byref(self.handle), C.byref(cfg))
        if st != _lib.HD_OK:
            msg = lib.hd_last_error(None)
            raise RuntimeError(f"hd_create failed (hd_status {st}): {msg.decode() if msg else '?'}")

    def close(self) -> None:
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.hd_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def check(self, status: int, what: str) -> None:
        _lib.check(self.handle, status, what)

    @staticmethod
    def _descs(state: dict):
        keep = []
        descs = (_lib.HdTensorDesc * len(state))()
        n = 0
        for name, t in state.items():
            if t.dtype == torch.float32:
                dt = 0
            elif t.dtype == torch.int64:
                dt = 1
            else:
                continue
            t = t.detach().contiguous()
            keep.append(t)
            d = descs[n]
            d.name = name.encode()
            d.data = t.data_ptr()
            d.dtype = dt
            d.ndim = min(t.dim(), 4)
            for k in range(d.ndim):
                d.shape[k] = t.shape[k] if k < 3 else int(math.prod(t.shape[3:]))
            if t.dim() == 0:
                d.ndim = 1
                d.shape[0] = 1
            n += 1
        return descs, n, keep

    def load_state(self, state: dict) -> None:
        descs, n, keep = self._descs(state)
        self.check(self.lib.hd_load_weights(self.handle, descs, n, None), "hd_load_weights")
        del keep

    def load_fpg_state(self, state: dict) -> None:
        descs, n, keep = self._descs(state)
        self.check(self.lib.hd_load_fpg_weights(self.handle, descs, n, None), "hd_load_fpg_weights")
        self.fpg_loaded = True
        del keep

    def fpg_forward(self, cr_latent: torch.Tensor, latent_size: int = 16, width: int = 128):
        """FacialPriorGuidance.forward on the sm_100a kernels -> list of 5 fp32 NCHW priors."""
        x = cr_latent.to(torch.float32).contiguous()
        b = x.shape[0]
        outs = [torch.empty((b, width << (4 - j), latent_size >> (4 - j), latent_size >> (4 - j)),
                            dtype=torch.float32, device=x.device) for j in range(5)]
        ptrs = (C.c_void_p * 5)(*[o.data_ptr() for o in outs])
        with torch.cuda.device(x.device):
            self.check(self.lib.hd_fpg_forward(self.handle, x.data_ptr(), ptrs, b, _stream_ptr(x.device)), "hd_fpg_forward")
        return outs

    def load_idc_state(self, state: dict) -> None:
        descs, n, keep = self._descs(state)
        self.check(self.lib.hd_load_idc_weights(self.handle, descs, n, None), "hd_load_idc_weights")
        self.idc_loaded = True
        del keep

    def idc_forward(self, cr_face: torch.Tensor) -> torch.Tensor:
        """ResNet50.forward (models/idc/model.py:123-136) on the sm_100a kernels -> (B,2048,1,1) fp32."""
        x = cr_face.to(torch.float32).contiguous()
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != x.shape[3]:
            raise ValueError(f"cr_face must be (B,3,H,H), got {tuple(x.shape)}")
        out = torch.empty((x.shape[0], 2048, 1, 1), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            self.check(self.lib.hd_idc_forward(self.handle, x.data_ptr(), x.shape[2], out.data_ptr(), x.shape[0],
                                               _stream_ptr(x.device)), "hd_idc_forward")
        return out

    def load_cr_state(self, state: dict) -> None:
        descs, n, keep = self._descs(state)
        self.check(self.lib.hd_load_cr_weights(self.handle, descs, n, None), "hd_load_cr_weights")
        self.cr_loaded = True
        del keep

    def cr_forward(self, ln_face: torch.Tensor) -> torch.Tensor:
        """CoarseRestoration.forward (models/cr/model.py:75-88) on CUDA kernels -> (B,3,128,128) fp32."""
        x = ln_face.to(torch.float32).contiguous()
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != x.shape[3]:
            raise ValueError(f"ln_face must be (B,3,H,H), got {tuple(x.shape)}")
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            self.check(self.lib.hd_cr_forward(self.handle, x.data_ptr(), x.shape[2], out.data_ptr(), x.shape[0],
                                              _stream_ptr(x.device)), "hd_cr_forward")
        return out

    def info(self) -> "_lib.HdInfo":
        info = _lib.HdInfo()
        self.check(self.lib.hd_get_info(self.handle, C.byref(info)), "hd_get_info")
        return info

    def synchronize(self) -> None:
        self.check(self.lib.hd_synchronize(self.handle), "hd_synchronize")


def _stream_ptr(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _DenoiserBase(nn.Module):
    _KIND = _lib.HD_MODEL_DENOISER

    def __init__(self, latent_size: int):
        super().__init__()
        self.width = 32 * 4
        self.dtype = torch.float32
        self.config = _Bag()
        self.config.in_channels = 4
        self.config.sample_size = latent_size

        time_dim = self.width * 4
        self.time_mlp = nn.Sequential(_NoParams(), nn.Linear(self.width, time_dim * 2), _NoParams(),
                                      nn.Linear(time_dim, time_dim))
        self.intro = nn.Conv2d(4, self.width, 3, padding=1)
        self.ending = nn.Conv2d(self.width, 4, 3, padding=1)
        self.encoders = nn.ModuleList()
        self.decoders = nn.ModuleList()
        self.middle_blks = nn.ModuleList()
        self.ups = nn.ModuleList()
        self.downs = nn.ModuleList()
        self._time_dim = time_dim
        # engine settings (not part of the reference API)
        self._engine: Optional[_Engine] = None
        self._engine_key = None
        self.precision = "bf16"
        self.max_batch = 64
        self.max_steps = 1000
        self.use_graph = True
        # a parent's load_state_dict reaches sub-modules through hooks, not through our override
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    # ---- engine management -------------------------------------------------------------------
    def configure(self, precision: Optional[str] = None, max_batch: Optional[int] = None,
                  max_steps: Optional[int] = None, use_graph: Optional[bool] = None):
        """precision: 'bf16' (tcgen05) or 'fp32' (FFMA correctness mode)."""
        if precision is not None:
            if precision not in ("bf16", "fp32"):
                raise ValueError("precision must be 'bf16' or 'fp32'")
            self.precision = precision
        if max_batch is not None:
            self.max_batch = int(max_batch)
        if max_steps is not None:
            self.max_steps = int(max_steps)
        if use_graph is not None:
            self.use_graph = bool(use_graph)
        self.invalidate()
        return self

    def invalidate(self) -> None:
        """Drop the packed copy of the weights; the next call re-reads the parameters."""
        if self._engine is not None:
            self._engine.close()
        self._engine = None
        self._engine_key = None

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate()
        return out

    def engine(self, batch: int = 1) -> _Engine:
        p = next(self.parameters())
        if p.device.type != "cuda":
            raise RuntimeError("hifidiff_b200 modules run only on CUDA (sm_100a); there is no CPU path. "
                               "Move the module with .to('cuda').")
        if batch > self.max_batch:
            self.max_batch = int(batch)
            self.invalidate()
        key = (p.device, self.precision, self.max_batch, self.max_steps, self.use_graph)
        if self._engine is None or self._engine_key != key:
            self.invalidate()
            with torch.cuda.device(p.device):
                torch.cuda.synchronize()
                eng = _Engine(self._KIND, self.config.sample_size,
                              _lib.HD_PRECISION_BF16 if self.precision == "bf16" else _lib.HD_PRECISION_FP32,
                              p.device, self.max_batch, self.max_steps, self.use_graph)
                # the reference evaluates the frequencies with torch fp32 ops (model.py:24-26)
                half = self.width // 2
                freqs = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1))).to(torch.float32).contiguous()
                eng.check(eng.lib.hd_set_time_frequencies(eng.handle, C.c_void_p(freqs.data_ptr())),
                          "hd_set_time_frequencies")
                eng.load_state(self.state_dict())
            self._engine, self._engine_key = eng, key
        return self._engine

    # ---- helpers -----------------------------------------------------------------------------
    def _timesteps(self, timesteps, batch: int, device) -> torch.Tensor:
        """Coerces every form the reference accepts (model.py:107-108, 218-229) to fp32 of length
        1 (shared) or `batch`."""
        if isinstance(timesteps, (int, float)):
            return torch.full((1,), float(timesteps), dtype=torch.float32, device=device)
        t = timesteps.to(device=device, dtype=torch.float32)
        if t.dim() == 0:
            t = t.reshape(1)
        if t.shape[0] not in (1, batch):
            raise ValueError(f"timesteps has {t.shape[0]} entries for a batch of {batch}")
        return t.contiguous()

    def _check_latents(self, latents: torch.Tensor) -> torch.Tensor:
        s = self.config.sample_size
        if latents.dim() != 4 or tuple(latents.shape[1:]) != (4, s, s):
            raise ValueError(f"latents must be (B,4,{s},{s}), got {tuple(latents.shape)}")
        if latents.device.type != "cuda":
            raise RuntimeError("hifidiff_b200 has no CPU path: latents must be a CUDA tensor")
        return latents.to(torch.float32).contiguous()

    def _denoise(self, latents: torch.Tensor, timesteps, taps: Optional[Sequence[str]] = None):
        x = self._check_latents(latents)
        b = x.shape[0]
        eng = self.engine(b)
        t = self._timesteps(timesteps, b, x.device)
        eps = torch.empty_like(x)
        with torch.cuda.device(x.device):
            stream = _stream_ptr(x.device)
            if not taps:
                eng.check(eng.lib.hd_denoise_step(eng.handle, x.data_ptr(), t.data_ptr(), t.shape[0], eps.data_ptr(),
                                                  b, stream), "hd_denoise_step")
                return eps, {}
            outs = {name: torch.empty(self._tap_shape(name, b), dtype=torch.float32, device=x.device) for name in taps}
            names = (C.c_char_p * len(taps))(*[n.encode() for n in taps])
            ptrs = (C.c_void_p * len(taps))(*[outs[n].data_ptr() for n in taps])
            eng.check(eng.lib.hd_denoise_step_taps(eng.handle, x.data_ptr(), t.data_ptr(), t.shape[0], eps.data_ptr(),
                                                   b, names, ptrs, len(taps), stream), "hd_denoise_step_taps")
            return eps, outs

    def _tap_shape(self, name: str, b: int):
        s = self.config.sample_size
        w = self.width
        parts = name.split(".")
        kind = parts[0]
        if kind == "time_mlp":
            return (b, 4 * w)
        if kind == "intro":
            return (b, w, s, s)
        if kind == "encoders":
            lvl = int(parts[1])
        elif kind == "downs":
            lvl = int(parts[1]) + 1
        elif kind == "middle_blks":
            lvl = 4
        elif kind in ("ups", "decoders"):
            lvl = 3 - int(parts[1])
        elif kind == "hcas":
            lvl = 4 - int(parts[1])
        else:
            raise ValueError(f"unknown tap '{name}'")
        return (b, w << lvl, s >> lvl, s >> lvl)

    def forward_with_taps(self, latents, timesteps, taps: Sequence[str], **cond):
        """Like forward, but also returns {tap_name: fp32 NCHW activation} for per-layer parity."""
        raise NotImplementedError


class Denoiser(_DenoiserBase):
    """Unconditional denoiser (pre-training): `forward(latents, timesteps)`."""

    def __init__(self, latent_size):
        super().__init__(latent_size)
        time_dim = self._time_dim
        chan = self.width
        for num in _ENC:
            self.encoders.append(_blocks(num, chan, time_dim))
            self.downs.append(nn.Conv2d(chan, 2 * chan, 2, 2))
            chan *= 2
        self.middle_blks = _blocks(_MID, chan, time_dim)
        for num in _DEC:
            self.ups.append(nn.Sequential(nn.Conv2d(chan, chan * 2, 1, bias=False), _NoParams()))
            chan //= 2
            self.decoders.append(_blocks(num, chan, time_dim))

    def forward(self, latents, timesteps):
        eps, _ = self._denoise(latents, timesteps)
        return UNet2DOutput(eps)

    def forward_with_taps(self, latents, timesteps, taps):
        eps, outs = self._denoise(latents, timesteps, taps)
        return UNet2DOutput(eps), outs


class FusedDenoiser(_DenoiserBase):
    """Conditional denoiser: `forward(latents, timesteps, facial_priors, identity_embedding)`.

    The identity (`idc_conv`) and prior (HCA gate) terms depend on neither x_t nor t; they are
    computed once per distinct (priors, identity) by `set_condition` and reused for every step.
    `forward` calls it automatically when handed tensors it has not seen.
    """
    _KIND = _lib.HD_MODEL_FUSED

    def __init__(self, latent_size):
        super().__init__(latent_size)
        time_dim = self._time_dim
        self.hcas = nn.ModuleList()
        chan = self.width
        for num in _ENC:
            self.encoders.append(_blocks(num, chan, time_dim))
            self.downs.append(nn.Conv2d(chan, 2 * chan, 2, 2))
            chan *= 2
        self.middle_blks = _blocks(_MID, chan, time_dim)
        self.idc_conv = nn.Conv2d(2048, (self.width * 16) * (latent_size // 16) ** 2, (1, 1))
        self.hcas.append(_HCAParams(chan))
        for num in _DEC:
            self.ups.append(nn.Sequential(nn.Conv2d(chan, chan * 2, 1, bias=False), _NoParams()))
            chan //= 2
            self.decoders.append(_blocks(num, chan, time_dim))
            self.hcas.append(_HCAParams(chan))
        self._cond_key = None

    def invalidate(self) -> None:
        super().invalidate()
        self._cond_key = None

    def set_condition(self, facial_priors: Sequence[torch.Tensor], identity_embedding: torch.Tensor) -> None:
        if len(facial_priors) != 5:
            raise ValueError("facial_priors must hold 5 tensors (models/fpg/model.py:46-64)")
        b = identity_embedding.shape[0]
        s = self.config.sample_size
        pri: List[torch.Tensor] = []
        for j, p in enumerate(facial_priors):
            want = (b, self.width << (4 - j), s >> (4 - j), s >> (4 - j))
            if tuple(p.shape) != want:
                raise ValueError(f"facial_priors[{j}] must be {want}, got {tuple(p.shape)}")
            if p.device.type != "cuda":
                raise RuntimeError("hifidiff_b200 has no CPU path: priors must be CUDA tensors")
            pri.append(p.to(torch.float32).contiguous())
        if identity_embedding.numel() != b * 2048:
            raise ValueError("identity_embedding must be (B,2048,1,1)")
        ident = identity_embedding.to(device=pri[0].device, dtype=torch.float32).contiguous()
        eng = self.engine(b)
        ptrs = (C.c_void_p * 5)(*[p.data_ptr() for p in pri])
        with torch.cuda.device(ident.device):
            eng.check(eng.lib.hd_set_condition(eng.handle, ptrs, ident.data_ptr(), b, _stream_ptr(ident.device)),
                      "hd_set_condition")
        # strong references: the storage of a cached tensor cannot be recycled for a fresh one while the key is alive
        self._cond_key = tuple((t, t._version) for t in list(facial_priors) + [identity_embedding])

    def _ensure_condition(self, priors, ident, batch: int) -> None:
        """Re-runs `set_condition` unless the SAME tensor objects, unmodified, were used last time."""
        if priors is None or ident is None:
            raise ValueError("FusedDenoiser needs facial_priors and identity_embedding")
        self.engine(batch)  # may invalidate the condition if the engine is rebuilt
        cur = list(priors) + [ident]
        key = self._cond_key
        if (key is None or len(key) != len(cur)
                or not all(k[0] is t and k[1] == t._version for k, t in zip(key, cur))):
            self.set_condition(priors, ident)

    def forward(self, latents, timesteps, facial_priors, identity_embedding):
        self._ensure_condition(facial_priors, identity_embedding, latents.shape[0])
        eps, _ = self._denoise(latents, timesteps)
        return UNet2DOutput(eps)

    def forward_with_taps(self, latents, timesteps, taps, facial_priors=None, identity_embedding=None):
        self._ensure_condition(facial_priors, identity_embedding, latents.shape[0])
        eps, outs = self._denoise(latents, timesteps, taps)
        return UNet2DOutput(eps), outs
