"""libpll_b200 - B200-native phylogenetic-likelihood kernels behind the libpll (pll.h) API.

The product is the shared library `libpll_b200/libpll_b200.so` (C host layer + sm_100a CUDA
kernels, sources under `libpll_b200/csrc/`, headers under `include/`).  This Python package
only holds the ctypes mirror of the C API (`binding.py`) used by the tests and the benchmark,
and the deterministic synthetic-workload generator (`synthetic.py`).

There is no CPU fallback: `load()` raises if the CUDA library has not been built, and the
library itself refuses to create a partition without a B200-class device.
"""
from __future__ import annotations

import os

from .binding import (  # noqa: F401
    PLL_ATTRIB_ARCH_AVX,
    PLL_ATTRIB_ARCH_AVX2,
    PLL_ATTRIB_ARCH_CPU,
    PLL_ATTRIB_ARCH_GPU,
    PLL_ATTRIB_ARCH_SSE,
    PLL_ATTRIB_PATTERN_TIP,
    PLL_ATTRIB_RATE_SCALERS,
    PLL_SCALE_BUFFER_NONE,
    OP_DTYPE,
    Partition,
    PllError,
    PllLibrary,
)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpll_b200.so")
REPO_ROOT = os.path.dirname(_HERE)

_lib = None


def load() -> PllLibrary:
    """The product library.  Fails loudly when it has not been built (no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or `make -C libpll_b200/csrc`). There is no CPU fallback.")
        _lib = PllLibrary(LIB_PATH, is_gpu=True)
    return _lib
