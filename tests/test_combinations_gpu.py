"""GPU parity over the CROSS PRODUCT of the path's switches: alphabet (4 / 20 / 5 states: fused, tensor-core and generic kernels) x pattern tips /
CLV tips x per-site / per-rate scalers x one device context / three pattern slices x invariant sites
x ascertainment-bias type, on trees long enough that scaler counts are not zero (a 300-taxon
caterpillar: tip-inner root edge; a 500-taxon random tree with long branches: inner-inner root
edge).  Every value a caller can obtain after a traversal - root and edge log-likelihood with the
per-pattern vectors, first and second derivative at two branch lengths - against the reference's
AVX2 path (oracle/_ref) at 1e-10.  (This sweep found the sliced per-rate root case fixed in
pll_devices.c: single switches were covered, their combinations were not.)"""
import itertools

import numpy as np
import pytest

from libpll_b200 import synthetic as S
from libpll_b200.binding import (PLL_ATTRIB_AB_FELSENSTEIN, PLL_ATTRIB_AB_FLAG, PLL_ATTRIB_AB_LEWIS,
                                 PLL_ATTRIB_AB_STAMATAKIS, PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_ARCH_CPU, PLL_ATTRIB_ARCH_GPU,
                                 PLL_ATTRIB_PATTERN_TIP, PLL_ATTRIB_RATE_SCALERS)

pytestmark = pytest.mark.gpu
RTOL = 1e-10
AB = (None, 0, PLL_ATTRIB_AB_LEWIS, PLL_ATTRIB_AB_FELSENSTEIN, PLL_ATTRIB_AB_STAMATAKIS)


def _same(got, want, scale):
    """within tolerance - or not a number on both sides (the Felsenstein term of a tree whose
    per-state site likelihoods underflow is 0 / 0 in the reference too)"""
    if np.isnan(want) or np.isnan(got):
        return bool(np.isnan(want) and np.isnan(got))
    return abs(got - want) <= RTOL * scale


def _tree(kind, states):
    from test_parity_gpu import _caterpillar

    if kind == "caterpillar":
        return _caterpillar(140 if states == 20 else 300, 150, states, seed=7)
    w = S.make_workload(120 if states == 20 else 500, 150, states=states, seed=23)
    w.branch_lengths = np.full(w.prob_matrices, 2.5)
    return w


@pytest.mark.parametrize("slices", [1, 3])
@pytest.mark.parametrize("rate_scalers", [False, True])
@pytest.mark.parametrize("pattern_tip", [True, False])
@pytest.mark.parametrize("kind", ["caterpillar", "random"])
@pytest.mark.parametrize("states", [4, 20, 5])
def test_switch_combinations(gpu_lib, ref_lib, states, kind, pattern_tip, rate_scalers, slices):
    if states == 5 and pattern_tip and rate_scalers:
        pytest.skip("the reference's plain-C tip-inner kernel ignores per-rate scalers (src/core_partials.c:461-510)")
    w = _tree(kind, states)
    ref_arch = PLL_ATTRIB_ARCH_AVX2 if states in (4, 20) else PLL_ATTRIB_ARCH_CPU   # odd alphabets: the plain-C path
    rates = ref_lib.gamma_rates(w.alpha, w.rate_cats)
    sites = w.sites
    checked = scaled = 0
    for pinv, ab in itertools.product((0.0, 0.2), AB):
        if ab and pinv:
            continue  # the reference refuses the pair (src/pll.c:1075-1087)
        if states != 4 and pattern_tip and ab is not None:
            continue  # the reference reads past its tables there (src/pll.c:885-903)
        extra = ((PLL_ATTRIB_PATTERN_TIP if pattern_tip else 0) | (PLL_ATTRIB_RATE_SCALERS if rate_scalers else 0) |
                 (PLL_ATTRIB_AB_FLAG if ab is not None else 0))
        assert gpu_lib.pll_gpu_set_devices(slices) == 1
        try:
            pg, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | extra, rates=rates)
        finally:
            gpu_lib.pll_gpu_set_devices(0)
        pr, _ = S.build_partition(ref_lib, w, ref_arch | extra, rates=rates)
        tag = (pinv, ab)
        for p in (pg, pr):
            if pinv:
                p.update_invariant_sites()
                for i in set(int(x) for x in pidx):
                    p.update_invariant_sites_proportion(i, pinv)
            if ab:
                p.set_asc_bias_type(ab)
                p.set_asc_state_weights(list(range(3, 3 + states)))
            p.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
            p.update_partials(w.ops)
        top = w.tips + w.inner - 1
        counts = np.asarray(pr.get_scaler(w.scaler_of(top)))
        assert np.array_equal(counts, np.asarray(pg.get_scaler(w.scaler_of(top)))), tag
        scaled += int(counts.sum())
        sg, sr = np.zeros(sites), np.zeros(sites)
        rg = pg.root_loglikelihood(top, w.scaler_of(top), pidx, persite=sg)
        rr = pr.root_loglikelihood(top, w.scaler_of(top), pidx, persite=sr)
        assert np.isfinite(rr) and abs(rg - rr) <= RTOL * abs(rr), ("root", tag, rg, rr)
        assert np.allclose(sg, sr, rtol=RTOL, atol=0), ("root per pattern", tag)
        a, b = w.root_a, w.root_b
        args = (a, w.scaler_of(a), b, w.scaler_of(b), w.root_matrix, pidx)
        eg, er = pg.edge_loglikelihood(*args, persite=sg), pr.edge_loglikelihood(*args, persite=sr)
        assert np.isfinite(er) and abs(eg - er) <= RTOL * abs(er), ("edge", tag, eg, er)
        assert np.allclose(sg, sr, rtol=RTOL, atol=0), ("edge per pattern", tag)
        tg, tr = pg.new_sumtable(), pr.new_sumtable()
        pg.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, tg)
        pr.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, tr)
        for t in (0.02, 0.7):
            dg = pg.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), t, pidx, tg)
            dr = pr.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), t, pidx, tr)
            scale = max(abs(dr[0]), float(w.weights.sum()) * 1e-3)
            assert _same(dg[0], dr[0], scale), ("d1", tag, t, dg, dr)
            assert _same(dg[1], dr[1], max(abs(dr[1]), scale)), ("d2", tag, t, dg, dr)
        checked += 1
        pg.destroy()
        pr.destroy()
    assert checked >= 2 and scaled > 0


@pytest.mark.parametrize("slices", [1, 3])
@pytest.mark.parametrize("rate_scalers", [False, True])
@pytest.mark.parametrize("recycled", [False, True])
@pytest.mark.parametrize("states,variant", [(4, "default"), (20, "default"), (20, "lg4m")])
def test_list_shapes_and_models(gpu_lib, ref_lib, states, variant, recycled, rate_scalers, slices):
    """The other axis: lists that recycle CLV / scaler slots (32 slots for 500 / 160 taxa: WAR / WAW
    hazards, dead stores, the 20-state whole-list kernel) x one / four rate matrices (LG4M: a
    different matrix and frequency vector per category) x scaler mode x slices x tip kind, each list
    run twice (the second time as a replayed CUDA graph, after new branch lengths)."""
    w = S.make_workload(500 if states == 4 else 160, 300, states=states, seed=31)
    w.branch_lengths = np.full(w.prob_matrices, 2.0)
    if recycled:
        w = S.recycle_slots(w, 32)
    rates = ref_lib.gamma_rates(w.alpha, w.rate_cats)
    for pattern_tip in (True, False):
        extra = (PLL_ATTRIB_PATTERN_TIP if pattern_tip else 0) | (PLL_ATTRIB_RATE_SCALERS if rate_scalers else 0)
        assert gpu_lib.pll_gpu_set_devices(slices) == 1
        try:
            pg, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | extra, variant=variant, rates=rates)
        finally:
            gpu_lib.pll_gpu_set_devices(0)
        pr, _ = S.build_partition(ref_lib, w, PLL_ATTRIB_ARCH_AVX2 | extra, variant=variant, rates=rates)
        a, b = w.root_a, w.root_b
        for rep in range(2):
            for p in (pg, pr):
                p.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths * (1.0 + 0.1 * rep))
                p.update_partials(w.ops)
            sg, sr = np.zeros(w.sites), np.zeros(w.sites)
            args = (a, w.scaler_of(a), b, w.scaler_of(b), w.root_matrix, pidx)
            eg, er = pg.edge_loglikelihood(*args, persite=sg), pr.edge_loglikelihood(*args, persite=sr)
            assert np.isfinite(er) and abs(eg - er) <= RTOL * abs(er), ("edge", pattern_tip, rep, eg, er)
            assert np.allclose(sg, sr, rtol=RTOL, atol=0), ("edge per pattern", pattern_tip, rep)
            inner = a if a >= w.tips else b
            rg = pg.root_loglikelihood(inner, w.scaler_of(inner), pidx)
            rr = pr.root_loglikelihood(inner, w.scaler_of(inner), pidx)
            assert np.isfinite(rr) and abs(rg - rr) <= RTOL * abs(rr), ("root", pattern_tip, rep, rg, rr)
            tg, tr = pg.new_sumtable(), pr.new_sumtable()
            pg.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, tg)
            pr.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, tr)
            dg = pg.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), 0.3, pidx, tg)
            dr = pr.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), 0.3, pidx, tr)
            scale = max(abs(dr[0]), float(w.weights.sum()) * 1e-3)
            assert _same(dg[0], dr[0], scale) and _same(dg[1], dr[1], max(abs(dr[1]), scale)), (pattern_tip, rep, dg, dr)
        assert sum(int(np.asarray(pr.get_scaler(w.scaler_of(x))).sum()) for x in (a, b) if x >= w.tips) > 0
        pg.destroy()
        pr.destroy()
