"""CPU: the C-ABI library loads, exports every symbol include/hifidiff_b200.h declares, and the
product fails loudly (no fallback) when there is no sm_100 GPU.  No compute calls here."""
import ctypes as C
import os
import re

import pytest
import torch

import hifidiff_b200 as H
from hifidiff_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_is_built_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = _lib.load()
    assert lib.hd_abi_version() == 1


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "hifidiff_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int32_t|void|const char\*)\s+(hd_[a-z_0-9]+)\s*\(", hdr, flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    lib = C.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_struct_sizes_match_header():
    assert C.sizeof(_lib.HdConfig) == 32
    assert C.sizeof(_lib.HdStepCoef) == 32
    assert C.sizeof(_lib.HdTensorDesc) == 8 + 8 + 4 + 4 + 32
    assert C.sizeof(_lib.HdInfo) == 6 * 4 + 3 * 8 + 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    cfg = _lib.HdConfig(C.sizeof(_lib.HdConfig), _lib.HD_MODEL_DENOISER, _lib.HD_PRECISION_BF16, 16, 0, 4, 50, 1)
    st = lib.hd_create(C.byref(h), C.byref(cfg))
    assert st == 3 and not h.value                       # HD_ERR_UNSUPPORTED
    assert b"no CUDA device" in lib.hd_last_error(None)
    cfg.struct_size = 7
    assert lib.hd_create(C.byref(h), C.byref(cfg)) == 1   # HD_ERR_INVALID
    assert lib.hd_create(None, None) == 1


def test_modules_refuse_cpu_tensors():
    with torch.device("meta"):
        m = H.Denoiser(16)
    m = m.to_empty(device="cpu")
    with pytest.raises(RuntimeError, match="no CPU path|only on CUDA"):
        m(torch.zeros(1, 4, 16, 16), 5)
    with pytest.raises(ValueError):
        m._check_latents(torch.zeros(1, 3, 16, 16))


def test_module_attribute_bag():
    with torch.device("meta"):
        m = H.FusedDenoiser(16)
    assert m.width == 128 and m.dtype == torch.float32
    assert m.config.in_channels == 4 and m.config.sample_size == 16
    assert m.idc_conv.out_channels == 2048
    out = H.UNet2DOutput(torch.zeros(1))
    assert hasattr(out, "sample")


def test_timestep_coercion_forms():
    with torch.device("meta"):
        m = H.FusedDenoiser(16)
    for tv in (7, 7.0, torch.tensor(7), torch.tensor([7]), torch.tensor([7.0])):
        t = m._timesteps(tv, 3, "cpu")
        assert t.dtype == torch.float32 and t.tolist() == [7.0]
    t = m._timesteps(torch.tensor([1, 2, 3]), 3, "cpu")
    assert t.tolist() == [1.0, 2.0, 3.0]
    with pytest.raises(ValueError):
        m._timesteps(torch.tensor([1, 2]), 3, "cpu")


def test_no_oracle_import_in_product():
    """The product package must never import the oracle (a routed-through oracle voids parity)."""
    pkg = os.path.join(ROOT, "hifidiff_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_built_library_is_blackwell_native():
    """Static check on the shipped binary (no GPU needed): it holds sm_100a code only, and the hot kernels carry the
    instructions their design claims — tcgen05.mma (UTCHMMA, incl. cta_group::2), tcgen05.ld (LDTM) and TMA tensor
    loads (UTMALDG) in the GEMM family and in the fused face / pair block kernels, the bulk L2 prefetch (UBLKPF) in
    the GEMMs, mma.sync (HMMA) + ldmatrix (LDSM) in the edge convs.  A rebuild that fell back to CUDA-core paths, or
    for another architecture, fails here before it reaches a GPU box."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    elf = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"\.(sm_\d+a?)\.", elf))
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    per_kernel, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per_kernel[cur] = set()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            per_kernel[cur].add(m.group(1))

    def family(tag):
        ks = [ops for name, ops in per_kernel.items() if tag in name]
        assert ks, f"no kernel matching {tag}"
        return ks

    for ops in family("gemm_tc_kernel") + family("gemm_tc2_kernel"):
        assert {"UTCHMMA", "LDTM", "UTMALDG"} <= ops, ops & {"UTCHMMA", "LDTM", "UTMALDG"}
    assert any("UBLKPF" in ops for ops in family("gemm_tc_kernel"))
    for tag in ("face_block_kernel", "pair_block_kernel"):
        for ops in family(tag):
            assert {"UTCHMMA", "LDTM", "UTMALDG"} <= ops, tag
    for tag in ("intro_mma_kernel", "ending_mma_kernel", "stn_conv_mma_kernel", "gemm_mma3"):
        for ops in family(tag):
            assert "HMMA" in ops, tag
    assert ".2CTA" in sass and "UTCBAR" in sass     # cta_group::2 pairs + tcgen05.commit
