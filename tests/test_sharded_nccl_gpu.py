"""GPU (needs two devices): site sharding across PROCESSES with the scalar reduction inside the
library - pll_gpu_comm_init, then the unchanged pll.h calls return the sums over all ranks
(ncclAllReduce of the `logl +=` / `d_f +=` couplings, reference src/core_likelihood_avx.c:1259,
src/core_derivatives_avx2.c:756-765).  Two worker processes, one GPU each, against the same
alignment in ONE partition on one device."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from libpll_b200 import synthetic as S
from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("states,sites", [(4, 20_000), (20, 3_000)])
def test_ranks_get_the_global_values_from_the_plain_calls(gpu_lib, tmp_path, states, sites):
    if gpu_lib.pll_gpu_device_count() < 2:
        pytest.skip("needs two GPUs (NCCL refuses two ranks on one device)")
    id_path = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "sharded_worker.py"), str(r), "2",
                               str(sites), id_path, str(states)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True, cwd=ROOT) for r in range(2)]
    outs = []
    for p in procs:
        out, err = p.communicate(timeout=300)
        assert p.returncode == 0, err[-3000:]
        outs.append(json.loads([l for l in out.splitlines() if l.startswith("{")][-1]))

    w = S.make_workload(24, sites, states=states, seed=23)
    part, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
    want = S.full_evaluation(part, w, pidx)
    a, b = w.root_a, w.root_b
    top = a if a >= w.tips else b
    want_root = part.root_loglikelihood(top, w.scaler_of(top), pidx)
    tab = part.new_sumtable()
    part.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, tab)
    want_d = [part.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), t, pidx, tab) for t in (0.05, 0.3)]
    part.destroy()

    assert outs[0]["lo"] == 0 and outs[0]["hi"] == outs[1]["lo"] and outs[1]["hi"] == sites
    for o in outs:
        assert o["collectives"] >= 5
        assert o["lnl"] == outs[0]["lnl"] and o["derivs"] == outs[0]["derivs"], "ranks disagree"
        assert abs(o["lnl"] - want) <= 1e-12 * abs(want) and o["lnl2"] == o["lnl"]
        assert abs(o["root"] - want_root) <= 1e-12 * abs(want_root)
        np.testing.assert_allclose(np.array(o["derivs"]), np.array(want_d), rtol=1e-10)
    # per-pattern values stay local: the two halves add up to the whole
    assert abs(outs[0]["local_lnl"] + outs[1]["local_lnl"] - want) <= 1e-12 * abs(want)
