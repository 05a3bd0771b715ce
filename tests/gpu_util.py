"""GPU-side helpers: build the product modules on cuda:0 with the seeded parity weights."""
import ctypes as C

import torch

import hifidiff_b200 as H
from hifidiff_b200 import _lib

from util import state_for


def build(cls, seed, precision="bf16", eps_gain=1.0, max_batch=8, max_steps=64, use_graph=True, args=(16,)):
    with torch.device("meta"):
        m = cls(*args)
    sd = state_for(m, seed, eps_gain)
    m = m.to_empty(device="cuda")
    m.load_state_dict(sd)
    m.eval()
    target = m.denoiser if hasattr(m, "denoiser") else m
    target.configure(precision=precision, max_batch=max_batch, max_steps=max_steps, use_graph=use_graph)
    return m, sd


class RawHandle:
    """A bare hd_handle (no weights) for kernel-level tests."""

    def __init__(self, max_batch=8, max_steps=8):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        cfg = _lib.HdConfig(C.sizeof(_lib.HdConfig), _lib.HD_MODEL_DENOISER, _lib.HD_PRECISION_BF16, 16, 0, max_batch,
                            max_steps, 0)
        st = self.lib.hd_create(C.byref(self.h), C.byref(cfg))
        assert st == 0, self.lib.hd_last_error(None)

    def check(self, st, what):
        _lib.check(self.h, st, what)

    def close(self):
        if self.h:
            self.lib.hd_destroy(self.h)
            self.h = C.c_void_p()

    def gemm(self, a, w, bias, use_tc):
        m, k = a.shape
        n = w.shape[0]
        out = torch.empty((m, n), dtype=torch.float32, device="cuda")
        self.check(self.lib.hd_debug_gemm(self.h, a.data_ptr(), w.data_ptr(), bias.data_ptr() if bias is not None else None,
                                          out.data_ptr(), m, n, k, int(use_tc), None), "hd_debug_gemm")
        torch.cuda.synchronize()
        return out
