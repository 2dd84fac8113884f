"""GPU: the golden cases recorded from the reference (tests/golden/cases.json - restatements of
reference test/src/00010_NMDU_lkcalc.c, 00011_NMAU_lkcalc.c, examples/unrooted, examples/newton
and the derivatives recipe) executed through the pll.h C-ABI of the CUDA library."""
import json
import os

import pytest

from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP
from golden_runner import execute
from test_oracle_cpu import GOLDEN, compare_outputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
@pytest.mark.parametrize("variant", ["tv", "notv"])
def test_gpu_reproduces_golden_cases(gpu_lib, case, variant):
    attr = PLL_ATTRIB_ARCH_GPU | (PLL_ATTRIB_PATTERN_TIP if variant == "tv" else 0)
    got = execute(gpu_lib, case, attr)
    compare_outputs(got, case["expect"][variant], 1e-10, f"{case['name']}[{variant}]")
