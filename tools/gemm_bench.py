"""TFLOP/s of the tcgen05 GEMM alone (device-timed, back-to-back launches, bias epilogue, fp32 out):
single-CTA 128x128 tiles vs cta_group::2 pairs on 256x256 tiles."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from gpu_util import RawHandle  # noqa: E402

raw = RawHandle()
shapes = [(16384, 4096, 4096), (8192, 8192, 2048), (32768, 2048, 1024), (65536, 256, 1152), (16384, 1024, 512),
          (4096, 2048, 1024), (1024, 2048, 1024), (65536, 256, 128)]
for (m, n, k) in shapes:
    a = torch.randn(m, k, device="cuda")
    w = torch.randn(n, k, device="cuda") / k ** 0.5
    out = torch.empty(m, n, device="cuda")
    row = []
    for mode, name in ((3, "1-CTA 128x128"), (2, "2-CTA 256x256")):
        ms = C.c_float()
        raw.check(raw.lib.hd_debug_gemm_time(raw.h, a.data_ptr(), w.data_ptr(), out.data_ptr(), m, n, k, mode, 20,
                                             C.byref(ms)), "hd_debug_gemm_time")
        row.append(f"{name}: {ms.value * 1e3:8.1f} us {2.0 * m * n * k / (ms.value * 1e-3) / 1e12:7.1f} TFLOP/s")
    print(f"M={m:6d} N={n:5d} K={k:5d}  " + "   ".join(row))
raw.close()
