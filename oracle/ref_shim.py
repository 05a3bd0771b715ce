"""Import the UNMODIFIED reference modules from /root/reference (container only; TEST INFRA).

`models/denoiser/model.py:4` does `from diffusers import ConfigMixin` and uses it purely as an
attribute bag (model.py:39-41,144-146); diffusers is not installed, so a 1-class stub module is
injected before the import.  Nothing on the GPU box may call this: /root/reference is absent
there.  Used by tests/golden/make_golden.py and by CPU-side cross-checks that skip when the
reference tree is missing.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HIFIDIFF_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models", "denoiser"))


def load():
    """Returns a namespace with the reference classes."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "diffusers" not in sys.modules:
        stub = types.ModuleType("diffusers")

        class ConfigMixin:  # attribute bag only
            pass

        stub.ConfigMixin = ConfigMixin
        sys.modules["diffusers"] = stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from models.denoiser.model import Denoiser, FusedDenoiser  # type: ignore
    from models.denoiser.conditional_naf import ConditionalNAFBlock  # type: ignore
    from models.fpg.hca import HybridCrossAttention  # type: ignore
    from models.fpg.model import FacialPriorGuidance  # type: ignore
    from models.idc.model import ResNet50  # type: ignore
    from models.refiner import FacialRefiner  # type: ignore
    from models.cr.model import CoarseRestoration  # type: ignore
    return types.SimpleNamespace(Denoiser=Denoiser, FusedDenoiser=FusedDenoiser,
                                 ConditionalNAFBlock=ConditionalNAFBlock,
                                 HybridCrossAttention=HybridCrossAttention,
                                 FacialPriorGuidance=FacialPriorGuidance, ResNet50=ResNet50,
                                 FacialRefiner=FacialRefiner, CoarseRestoration=CoarseRestoration)
