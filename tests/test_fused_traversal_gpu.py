"""GPU: the single-kernel traversal (libpll_b200/csrc/gpu/plg_traverse.cu, default for DNA) against
the level-by-level kernels (PLL_GPU_FUSED=0) on the same partitions: every CLV and every scaler
array must be bit-identical, for plain and slot-recycling lists, per-site and per-rate scalers,
with a tile cache of 1, 2 and 3 slots (1 slot forces many reads back from HBM), for alignment
lengths that leave partial tiles, and for 1, 2, 4, 8 and 16 rate categories."""
import numpy as np
import pytest

from libpll_b200 import synthetic as S
from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP, PLL_ATTRIB_RATE_SCALERS
from test_parity_gpu import _caterpillar

pytestmark = pytest.mark.gpu


def _run(gpu_lib, monkeypatch, w, attrs, fused, slots=3):
    monkeypatch.setenv("PLL_GPU_FUSED", "1" if fused else "0")
    monkeypatch.setenv("PLL_GPU_FUSED_SLOTS", str(slots))
    part, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | attrs)
    part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    part.reset_stats()
    part.update_partials(w.ops)
    launches = part.stats()["kernel_launches"]
    clvs = {int(o["parent_clv_index"]): part.get_clv(int(o["parent_clv_index"])).tobytes() for o in w.ops}
    scalers = {int(o["parent_scaler_index"]): part.get_scaler(int(o["parent_scaler_index"])).tobytes()
               for o in w.ops if int(o["parent_scaler_index"]) >= 0}
    lnl = part.edge_loglikelihood(w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b),
                                  w.root_matrix, pidx)
    part.destroy()
    return clvs, scalers, lnl, launches


@pytest.mark.parametrize("tips,sites,cats", [(40, 1000, 4), (150, 4097, 4), (33, 777, 1), (64, 2050, 2), (30, 1500, 8),
                                             (25, 900, 16)])
@pytest.mark.parametrize("rate_scalers", [False, True])
@pytest.mark.parametrize("pattern_tip", [True, False])
def test_fused_equals_level_by_level(gpu_lib, monkeypatch, tips, sites, cats, rate_scalers, pattern_tip):
    w = S.make_workload(tips, sites, states=4, rate_cats=cats, seed=tips + cats)
    attrs = (PLL_ATTRIB_PATTERN_TIP if pattern_tip else 0) | (PLL_ATTRIB_RATE_SCALERS if rate_scalers else 0)
    ref = _run(gpu_lib, monkeypatch, w, attrs, fused=False)
    assert ref[3] > 3, "the level-by-level path launches one kernel per level and kind"
    for slots in (1, 2, 3):
        got = _run(gpu_lib, monkeypatch, w, attrs, fused=True, slots=slots)
        assert got[3] == 2, "pack + traverse"
        assert got[0] == ref[0], f"CLVs differ with {slots} cache slots"
        assert got[1] == ref[1], f"scalers differ with {slots} cache slots"
        assert got[2] == ref[2]


@pytest.mark.parametrize("rate_scalers", [False, True])
def test_fused_with_rescaling_and_recycled_slots(gpu_lib, monkeypatch, rate_scalers):
    attrs = PLL_ATTRIB_PATTERN_TIP | (PLL_ATTRIB_RATE_SCALERS if rate_scalers else 0)
    # long caterpillar: repeated rescaling, every operation hands its result to the next one
    w = _caterpillar(400, 300, 4, seed=5)
    ref = _run(gpu_lib, monkeypatch, w, attrs, fused=False)
    got = _run(gpu_lib, monkeypatch, w, attrs, fused=True)
    assert got[0] == ref[0] and got[1] == ref[1] and got[2] == ref[2]
    assert any(np.frombuffer(v, np.uint32).any() for v in ref[1].values()), "no rescaling happened"
    # a list that recycles CLV / scaler slots is executed in its own order
    w = S.recycle_slots(S.make_workload(120, 900, states=4, seed=9), 9)
    ref = _run(gpu_lib, monkeypatch, w, attrs, fused=False)
    got = _run(gpu_lib, monkeypatch, w, attrs, fused=True)
    assert got[0] == ref[0] and got[1] == ref[1] and got[2] == ref[2]


def test_operations_chained_through_a_scaler_only(gpu_lib, monkeypatch):
    """A legal if odd list in which operation B reads the scaler that operation A wrote while none
    of B's CLVs comes from A: A, D, B, C with C = parent of A and B.  The depth-first reordering of
    the single-kernel traversal (larger subtree first: D, B, A, C) would let B read the scaler
    before A has written it; the planner checks the new order against every read-after-write
    relation of the list and then executes it as given.  Scaler 0 is pre-filled with a marker that
    A (tip-tip: zeroes its scaler, reference src/core_partials_avx.c:598-599) must clear first."""
    from libpll_b200.binding import OP_DTYPE
    import ctypes as C

    w = S.make_workload(8, 640, states=4, seed=21)
    results = []
    for fused in (False, True):
        monkeypatch.setenv("PLL_GPU_FUSED", "1" if fused else "0")
        part, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
        mats = np.arange(w.prob_matrices, dtype=np.uint32)
        part.update_prob_matrices(pidx, mats, np.linspace(0.05, 0.4, w.prob_matrices))
        T, NONE = w.tips, -1
        prev = np.array([(T + 3, NONE, 4, 4, NONE, 5, 5, NONE)], dtype=OP_DTYPE)      # an earlier call
        part.update_partials(prev)
        marker = np.full(w.sites, 7, dtype=np.uint32)
        assert gpu_lib.plg_set_scaler(part.ctx(), 0, marker.ctypes.data_as(C.POINTER(C.c_uint))) == 0
        ops = np.array([(T + 0, 0, 0, 0, NONE, 1, 1, NONE),          # A: tip-tip, clears scaler 0
                        (T + 4, NONE, 2, 2, NONE, 3, 3, NONE),        # D: tip-tip
                        (T + 1, 1, T + 4, 6, NONE, T + 3, 7, 0),      # B: reads scaler 0 with a CLV of `prev`
                        (T + 2, 2, T + 0, 8, 0, T + 1, 9, 1)],        # C: parent of A and B
                       dtype=OP_DTYPE)
        part.update_partials(ops)
        results.append(([part.get_clv(T + k).tobytes() for k in range(5)],
                        [part.get_scaler(k).copy() for k in range(3)]))
        part.destroy()
    (clv0, sc0), (clv1, sc1) = results
    assert clv0 == clv1
    for a, b in zip(sc0, sc1):
        assert np.array_equal(a, b)
    assert not sc0[0].any() and not sc0[1].any() and not sc0[2].any(), "the marker leaked into a result"
