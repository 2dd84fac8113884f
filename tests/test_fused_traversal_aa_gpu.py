"""GPU: the 20-state single-kernel walk (libpll_b200/csrc/gpu/plg_walk_aa.cu; the default for
operation lists that recycle CLV / scaler slots, PLL_GPU_FUSED_AA=1 forces it for every list of a
protein partition with 1, 2 or 4 rate categories and per-site scalers) against

  * the level-by-level tensor-core kernels (PLL_GPU_FUSED=0): both run the same DMMA chains in the
    same order, so every CLV and every scaler array must be BIT-identical - plain and
    slot-recycling lists, per-site and per-rate scalers, 1 and 2 tile-cache slots (1 slot forces
    reads back from HBM), alignment lengths that leave partial 16-pattern tiles and partial
    8-pattern groups, CLV tips and pattern tips, LG and the LG4M four-matrix mixture;
  * the reference's AVX2 path (oracle/_ref): scalers bit-exact, CLVs within 1e-12 relative
    (DMMA sums the 20 products of a row in another order than the AVX2 lanes), lnL within 1e-10.
"""
import numpy as np
import pytest

from libpll_b200 import synthetic as S
from libpll_b200.binding import (PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP,
                                 PLL_ATTRIB_RATE_SCALERS)
from test_parity_gpu import _caterpillar

pytestmark = pytest.mark.gpu

def _select(monkeypatch, fused=True):
    monkeypatch.setenv("PLL_GPU_FUSED_AA", "1" if fused else "0")


def _fused_here(rate_scalers):
    """Per-rate scalers stay on the level-by-level kernels."""
    return not rate_scalers


def _run(gpu_lib, monkeypatch, w, attrs, fused, slots=3, variant="default"):
    monkeypatch.setenv("PLL_GPU_FUSED", "1" if fused else "0")
    _select(monkeypatch, fused)
    monkeypatch.setenv("PLL_GPU_FUSED_SLOTS", str(slots))
    monkeypatch.setenv("PLL_GPU_AA_EXACT", "0")
    part, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | attrs, variant=variant)
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


def _same(got, ref, what):
    for k in ref[0]:
        assert got[0][k] == ref[0][k], f"{what}: CLV {k} differs"
    for k in ref[1]:
        assert got[1][k] == ref[1][k], f"{what}: scaler {k} differs"
    assert got[2] == ref[2], what


@pytest.mark.parametrize("tips,sites,cats", [(40, 1000, 4), (150, 4097, 4), (33, 777, 1), (64, 2050, 2), (9, 7, 4),
                                             (21, 113, 4)])
@pytest.mark.parametrize("rate_scalers", [False, True])
@pytest.mark.parametrize("pattern_tip", [True, False])
def test_fused_equals_level_by_level(gpu_lib, monkeypatch, tips, sites, cats, rate_scalers, pattern_tip):
    w = S.make_workload(tips, sites, states=20, rate_cats=cats, seed=tips + cats)
    attrs = (PLL_ATTRIB_PATTERN_TIP if pattern_tip else 0) | (PLL_ATTRIB_RATE_SCALERS if rate_scalers else 0)
    ref = _run(gpu_lib, monkeypatch, w, attrs, fused=False)
    assert ref[3] > 3, "the level-by-level path launches one kernel per level and kind"
    for slots in (1, 2, 3):
        got = _run(gpu_lib, monkeypatch, w, attrs, fused=True, slots=slots)
        assert (got[3] == 2) == _fused_here(rate_scalers), "pack + traverse"
        _same(got, ref, f"{slots} cache slots")


def test_other_category_counts_keep_the_level_by_level_kernels(gpu_lib, monkeypatch):
    w = S.make_workload(20, 500, states=20, rate_cats=8, seed=3)
    ref = _run(gpu_lib, monkeypatch, w, PLL_ATTRIB_PATTERN_TIP, fused=False)
    got = _run(gpu_lib, monkeypatch, w, PLL_ATTRIB_PATTERN_TIP, fused=True)
    assert got[3] == ref[3] > 3
    _same(got, ref, "8 categories")


def test_lg4m_mixture(gpu_lib, monkeypatch):
    """Four rate matrices, one per category (reference examples/lg4/lg4.c:295-310)."""
    w = S.make_workload(60, 1500, states=20, rate_cats=4, seed=17)
    ref = _run(gpu_lib, monkeypatch, w, PLL_ATTRIB_PATTERN_TIP, fused=False, variant="lg4m")
    got = _run(gpu_lib, monkeypatch, w, PLL_ATTRIB_PATTERN_TIP, fused=True, variant="lg4m")
    assert got[3] == 2
    _same(got, ref, "LG4M")


@pytest.mark.parametrize("rate_scalers", [False, True])
def test_fused_with_rescaling_and_recycled_slots(gpu_lib, monkeypatch, rate_scalers):
    attrs = PLL_ATTRIB_PATTERN_TIP | (PLL_ATTRIB_RATE_SCALERS if rate_scalers else 0)
    # long caterpillar: repeated rescaling, every operation hands its result to the next one
    w = _caterpillar(200, 150, 20, seed=5)
    ref = _run(gpu_lib, monkeypatch, w, attrs, fused=False)
    got = _run(gpu_lib, monkeypatch, w, attrs, fused=True)
    _same(got, ref, "caterpillar")
    assert any(np.frombuffer(v, np.uint32).any() for v in ref[1].values()), "no rescaling happened"
    # a list that recycles CLV / scaler slots is executed in its own order; with one cache slot
    # most children come back from HBM, and most stores are dead
    w = S.recycle_slots(S.make_workload(120, 900, states=20, seed=9), 9)
    ref = _run(gpu_lib, monkeypatch, w, attrs, fused=False)
    for slots in (1, 2):
        got = _run(gpu_lib, monkeypatch, w, attrs, fused=True, slots=slots)
        _same(got, ref, f"recycled, {slots} slots")


@pytest.mark.parametrize("tips,sites", [(200, 5000), (30, 333)])
def test_fused_against_the_reference(gpu_lib, ref_lib, monkeypatch, tips, sites):
    monkeypatch.setenv("PLL_GPU_FUSED", "1")
    _select(monkeypatch)
    monkeypatch.setenv("PLL_GPU_AA_EXACT", "0")
    w = S.make_workload(tips, sites, states=20, seed=tips)
    rates = ref_lib.gamma_rates(w.alpha, w.rate_cats)
    pg, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP, rates=rates)
    pr, _ = S.build_partition(ref_lib, w, PLL_ATTRIB_ARCH_AVX2 | PLL_ATTRIB_PATTERN_TIP, rates=rates)
    pg.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    pr.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    for m in range(w.prob_matrices):
        pg.set_pmatrix(m, pr.get_pmatrix(m))
    pg.reset_stats()
    pg.update_partials(w.ops)
    assert pg.stats()["kernel_launches"] == 2
    pr.update_partials(w.ops)
    for k in range(w.inner):
        np.testing.assert_array_equal(pg.get_scaler(k), pr.get_scaler(k), err_msg=f"scaler {k}")
        np.testing.assert_allclose(pg.get_clv(w.tips + k), pr.get_clv(w.tips + k), rtol=1e-12, atol=0,
                                   err_msg=f"CLV {w.tips + k}")
    ps_g, ps_r = np.zeros(sites), np.zeros(sites)
    args = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
    lg = pg.edge_loglikelihood(*args, persite=ps_g)
    lr = pr.edge_loglikelihood(*args, persite=ps_r)
    np.testing.assert_allclose(ps_g, ps_r, rtol=1e-10, atol=0)
    assert abs(lg - lr) <= 1e-10 * abs(lr)
    pg.destroy()
    pr.destroy()


def test_repeated_calls_replay_a_graph(gpu_lib, monkeypatch):
    """The second sighting of a list captures a CUDA graph (pack + traverse); replays give the same bits."""
    monkeypatch.setenv("PLL_GPU_FUSED", "1")
    _select(monkeypatch)
    monkeypatch.setenv("PLL_GPU_AA_EXACT", "0")
    w = S.make_workload(50, 2000, states=20, seed=4)
    part, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
    part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    top = w.tips + w.inner - 1
    part.update_partials(w.ops)
    first = part.get_clv(top).tobytes()
    part.reset_stats()
    for _ in range(3):
        part.update_partials(w.ops)
    assert part.stats()["graph_launches"] >= 2
    assert part.get_clv(top).tobytes() == first
    part.destroy()


def test_the_walk_is_deterministic(gpu_lib, monkeypatch):
    """Math and DMA warps hand tiles over through mbarriers only; a missed ordering would show as a
    result that moves between identical calls.  The same 300-operation list 40 times: the root CLV
    and the log-likelihood must not change by a bit."""
    monkeypatch.setenv("PLL_GPU_FUSED", "1")
    _select(monkeypatch)
    monkeypatch.setenv("PLL_GPU_AA_EXACT", "0")
    w = S.make_workload(300, 20000, states=20, seed=12)
    part, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
    part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    top = w.tips + w.inner - 1
    args = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
    seen = set()
    for _ in range(40):
        part.update_partials(w.ops)
        seen.add((part.edge_loglikelihood(*args), part.get_clv(top).tobytes()))
    assert len(seen) == 1
    part.destroy()


@pytest.mark.parametrize("walk", [False, True])
def test_benchmark_shape_against_the_reference(gpu_lib, ref_lib, monkeypatch, walk):
    """BASELINE configs[2] itself: the 500-taxon LG+G4 operations list of bench.py's C3 record (173
    tip-tip / 154 tip-inner / 171 inner-inner operations) on a 6 000-pattern window of the 200 000-
    pattern alignment, device - level-by-level kernels and the whole-list walk - against the
    reference's AVX2 path with the same P-matrices: all 498 scaler arrays bit for bit, CLVs within
    1e-12 relative (DMMA sums a row's 20 products in another order), per-pattern lnL within 1e-10."""
    monkeypatch.setenv("PLL_GPU_FUSED", "1")
    _select(monkeypatch, fused=walk)
    monkeypatch.setenv("PLL_GPU_AA_EXACT", "0")
    w = S.make_workload(500, 200_000, states=20)
    lo, hi = 97_000, 103_000
    rates = ref_lib.gamma_rates(w.alpha, w.rate_cats)
    pg, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP, lo=lo, hi=hi, rates=rates)
    pr, _ = S.build_partition(ref_lib, w, PLL_ATTRIB_ARCH_AVX2 | PLL_ATTRIB_PATTERN_TIP, lo=lo, hi=hi, rates=rates)
    pg.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    pr.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    for m in range(w.prob_matrices):
        pg.set_pmatrix(m, pr.get_pmatrix(m))
    pg.reset_stats()
    pg.update_partials(w.ops)
    assert (pg.stats()["kernel_launches"] == 2) == walk
    pr.update_partials(w.ops)
    for k in range(w.inner):
        np.testing.assert_array_equal(pg.get_scaler(k), pr.get_scaler(k), err_msg=f"scaler {k}")
        np.testing.assert_allclose(pg.get_clv(w.tips + k), pr.get_clv(w.tips + k), rtol=1e-12, atol=0,
                                   err_msg=f"CLV {w.tips + k}")
    ps_g, ps_r = np.zeros(hi - lo), np.zeros(hi - lo)
    args = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
    lg = pg.edge_loglikelihood(*args, persite=ps_g)
    lr = pr.edge_loglikelihood(*args, persite=ps_r)
    np.testing.assert_allclose(ps_g, ps_r, rtol=1e-10, atol=0)
    assert abs(lg - lr) <= 1e-10 * abs(lr)
    pg.destroy()
    pr.destroy()


def test_default_selection(gpu_lib, monkeypatch):
    """Without PLL_GPU_FUSED_AA a list that recycles slots (most of its stores are dead: the walk
    is 16-22 % faster there) runs as pack + walk, a list with one slot per node on the level-by-level
    kernels; both give the bits of the forced level-by-level run."""
    monkeypatch.setenv("PLL_GPU_FUSED", "1")
    monkeypatch.delenv("PLL_GPU_FUSED_AA", raising=False)
    monkeypatch.setenv("PLL_GPU_AA_EXACT", "0")
    plain = S.make_workload(60, 1500, states=20, seed=21)
    recycled = S.recycle_slots(plain, 9)
    for w, walk in ((plain, False), (recycled, True)):
        part, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
        part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
        part.reset_stats()
        part.update_partials(w.ops)
        assert (part.stats()["kernel_launches"] == 2) == walk
        lnl = part.edge_loglikelihood(w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b),
                                      w.root_matrix, pidx)
        part.destroy()
        ref = _run(gpu_lib, monkeypatch, w, PLL_ATTRIB_PATTERN_TIP, fused=False)
        monkeypatch.setenv("PLL_GPU_FUSED", "1")      # _run switched the whole-list kernels off
        monkeypatch.delenv("PLL_GPU_FUSED_AA", raising=False)
        assert lnl == ref[2]
