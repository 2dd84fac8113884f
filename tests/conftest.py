"""Shared fixtures.

Markers:  @pytest.mark.gpu  - needs a B200 (run by `pytest -m gpu` on the GPU box);
          everything else runs on CPU (`pytest -m "not gpu"`).

Libraries under test / used as checkers:
  gpu_lib   libpll_b200/libpll_b200.so          the product (CUDA); must load on CPU too
  ref_lib   oracle/_ref/libpll_ref.so           the unmodified reference built by oracle/Makefile
                                                (test infrastructure; travels to the GPU box)
  port      oracle/libpll_oracle.so             our plain-C restatement (oracle/pll_oracle.c)
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")


def _has_gpu() -> bool:
    try:
        import libpll_b200

        return libpll_b200.load().plg_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_lib():
    import libpll_b200

    return libpll_b200.load()


@pytest.fixture(scope="session")
def has_gpu():
    return _has_gpu()


@pytest.fixture(scope="session")
def ref_lib():
    from libpll_b200.binding import PllLibrary

    path = os.path.join(ROOT, "oracle", "_ref", "libpll_ref.so")
    if not os.path.exists(path) and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref", "-j8"], check=True,
                       stdout=subprocess.DEVNULL)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libpll_ref.so not available (reference tree absent and no prebuilt copy)")
    return PllLibrary(path, is_gpu=False)


@pytest.fixture(scope="session")
def port_lib():
    """ctypes handle of the oracle's C restatement (oracle/libpll_oracle.so)."""
    from oracle import port as oracle_port

    return oracle_port.load()
