"""Per-CTA timeline of the tcgen05 GEMM (clock64 deltas, medians over CTAs)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from gpu_util import RawHandle  # noqa: E402

raw = RawHandle()
for (m, n, k) in [(65536, 256, 128), (65536, 128, 128), (256, 2048, 2048), (1024, 2048, 1024), (16384, 512, 256)]:
    a = torch.randn(m, k, device="cuda")
    w = torch.randn(n, k, device="cuda") / k ** 0.5
    out = torch.empty(m, n, device="cuda")
    cap = 4096
    tr = np.zeros((cap, 16), dtype=np.int64)
    nct = C.c_int32()
    raw.check(raw.lib.hd_debug_gemm_trace(raw.h, a.data_ptr(), w.data_ptr(), out.data_ptr(), m, n, k,
                                          tr.ctypes.data, cap, C.byref(nct), None), "trace")
    t = tr[: nct.value]
    names = ["setup", "first_ops", "mma_issue", "acc_ready(from mma)", "stageA", "barrier", "phaseB", "exit"]
    pairs = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8)]
    life = t[:, 8] - t[:, 0]
    gt = t[:, 9] - t[:, 9].min()
    print(f"M={m} N={n} K={k}: {nct.value} CTAs, CTA lifetime median {np.median(life):.0f} cyc (p90 {np.percentile(life, 90):.0f}); "
          f"entry globaltimer spread {gt.max() / 1e3:.1f} us; SMs used {len(set(t[:, 10].tolist()))}")
    for nm, (i, j) in zip(names, pairs):
        d = t[:, j] - t[:, i]
        print(f"   {nm:22s} median {np.median(d):8.0f}  p90 {np.percentile(d, 90):8.0f} cycles")
raw.close()
