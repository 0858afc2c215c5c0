"""rvtests_b200 -- B200-native engine for the rvtests per-gene association hot path.

The product is the C-ABI library librvtests_b200.so (include/rvtests_b200.h) and the C++
ModelFitter-shaped adapters under rvtests_b200/host/.  This Python package is a thin ctypes
mirror of that ABI used by the tests and by bench.py; it performs no arithmetic itself and has
no CPU fallback -- every compute call fails loudly without a CUDA device.
"""
from .engine import GeneEngine, GeneResult, RvtError, load_library  # noqa: F401

__all__ = ["GeneEngine", "GeneResult", "RvtError", "load_library"]
