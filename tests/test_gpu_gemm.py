"""GPU: the tcgen05/TMA GEMM and the FFMA GEMM alone, through the C ABI, against torch fp32."""
import pytest
import torch

from gpu_util import RawHandle
from util import gen, rel_l2

pytestmark = pytest.mark.gpu

SHAPES = [  # (M, N, K): full tiles, ragged M, single row, deep K, every N/K the network uses
    (128, 128, 64), (128, 128, 128), (256, 256, 128), (200, 128, 192), (1, 4096, 2048), (64, 2048, 2048),
    (300, 512, 1024), (1024, 256, 512), (4096, 128, 128), (130, 1024, 4096), (16384, 256, 128),
]


@pytest.fixture(scope="module")
def raw():
    h = RawHandle()
    yield h
    h.close()


@pytest.mark.parametrize("m,n,k", SHAPES)
def test_tc_gemm_matches_fp32_reference(raw, m, n, k):
    a = torch.randn((m, k), generator=gen(m + n)).cuda()
    w = (torch.randn((n, k), generator=gen(k + 1)) / k ** 0.5).cuda()
    b = torch.randn((n,), generator=gen(3)).cuda()
    out = raw.gemm(a, w, b, use_tc=True)
    # exact statement of what the kernel computes: bf16-rounded operands, fp32 accumulate
    ref_bf = a.bfloat16().double() @ w.bfloat16().double().t() + b.double()
    ref_fp = a.double() @ w.double().t() + b.double()
    assert rel_l2(out, ref_bf) < 2e-5, "tensor-core result differs from bf16-operand / fp32-accumulate math"
    assert rel_l2(out, ref_fp) < 1e-2


TWO_CTA_SHAPES = [(256, 256, 64), (256, 256, 512), (1000, 512, 1024), (4096, 1024, 512), (130, 256, 2048), (16384, 256, 128)]


@pytest.mark.parametrize("m,n,k", TWO_CTA_SHAPES)
def test_two_cta_gemm_matches_fp32_reference(raw, m, n, k):
    """cta_group::2: a CTA pair per 256x256 tile (ragged M: the odd 128-row tile and row tails are masked)."""
    a = torch.randn((m, k), generator=gen(m + n + 1)).cuda()
    w = (torch.randn((n, k), generator=gen(k + 2)) / k ** 0.5).cuda()
    b = torch.randn((n,), generator=gen(4)).cuda()
    out = raw.gemm(a, w, b, use_tc=2)
    ref_bf = a.bfloat16().double() @ w.bfloat16().double().t() + b.double()
    assert rel_l2(out, ref_bf) < 2e-5
    one = raw.gemm(a, w, b, use_tc=3)                     # single-CTA tiles (may split K: different fp32 summation order)
    assert rel_l2(out, one) < 1e-5


@pytest.mark.parametrize("m,n,k", [(1, 1, 64), (70, 130, 36 * 4), (257, 64, 128), (64, 124928 // 64, 256)])
def test_ffma_gemm_matches_fp32_reference(raw, m, n, k):
    a = torch.randn((m, k), generator=gen(m)).cuda()
    w = (torch.randn((n, k), generator=gen(n)) / k ** 0.5).cuda()
    b = torch.randn((n,), generator=gen(5)).cuda()
    out = raw.gemm(a, w, b, use_tc=False)
    ref = a.double() @ w.double().t() + b.double()
    assert rel_l2(out, ref) < 2e-6


def test_tc_gemm_rejects_bad_shapes(raw):
    a = torch.zeros((8, 32)).cuda()
    w = torch.zeros((128, 32)).cuda()
    out = torch.zeros((8, 128)).cuda()
    st = raw.lib.hd_debug_gemm(raw.h, a.data_ptr(), w.data_ptr(), None, out.data_ptr(), 8, 128, 32, 1, None)
    assert st == 1 and b"K % 64" in raw.lib.hd_last_error(raw.h)
