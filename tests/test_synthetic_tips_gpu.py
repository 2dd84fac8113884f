"""GPU: tips generated on the device (pll_gpu_generate_tip_states -> plg_generate_tipchars,
libpll_b200/csrc/gpu/plg_synth.cu) are the characters the host restatement
libpll_b200/synthetic.py:hash_tip_sequence gives, for any slice of the alignment and any number of
pattern slices per partition - so the CPU arm and the device see the same 5 000 x 10 M alignment
without 50 GB of characters ever being made on the host."""
import numpy as np
import pytest

from libpll_b200 import synthetic as S
from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP

pytestmark = pytest.mark.gpu


def _tipchars(part, tip, sites):
    assert part.lib.pll_gpu_sync_tipchars(part.ptr, tip) == 1, part.lib.errmsg()
    return bytes(bytearray(part.p.tipchars[tip][:sites]))


@pytest.mark.parametrize("slices", [1, 3])
def test_device_tips_equal_the_host_restatement(gpu_lib, slices):
    tips, sites, seed, first = 4, 70_000, 43, 9_999_000
    gpu_lib.pll_gpu_set_devices(slices)
    try:
        kw = dict(tips=tips, clv_buffers=2, states=4, sites=sites, rate_matrices=1, prob_matrices=6, rate_cats=4,
                  scale_buffers=2, attributes=PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
        dev, host = gpu_lib.partition(**kw), gpu_lib.partition(**kw)
    finally:
        gpu_lib.pll_gpu_set_devices(0)
    for t in range(tips):
        assert gpu_lib.pll_gpu_generate_tip_states(dev.ptr, t, seed, first) == 1, gpu_lib.errmsg()
        host.set_tip_states(t, S.hash_tip_sequence(seed, t, first, first + sites))
        assert _tipchars(dev, t, sites) == _tipchars(host, t, sites), f"tip {t}"
    # composition: 70 % root state, ~1 % N, ~0.5 % two-state ambiguities
    row = np.frombuffer(_tipchars(dev, 0, sites), np.uint8)
    assert 0.007 < np.mean(row == 15) < 0.013 and 0.003 < np.mean((row == 5) | (row == 10)) < 0.007
    other = np.frombuffer(_tipchars(dev, 1, sites), np.uint8)
    assert 0.5 < np.mean(row == other) < 0.65      # 0.7^2 + 0.3^2 * ... agreement through the common root
    dev.destroy()
    host.destroy()
