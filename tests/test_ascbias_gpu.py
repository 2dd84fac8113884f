"""GPU parity for ascertainment-bias correction (recipe of reference
test/src/asc-bias.c:170-270 on synthetic data - its testdata/2000.fas is not shipped): a partition
created with PLL_ATTRIB_AB_FLAG, then no correction / Lewis / Felsenstein / Stamatakis, comparing
edge and root log-likelihoods and the Newton derivatives with the reference's own path
(oracle/_ref) at 1e-10."""
import numpy as np
import pytest

from libpll_b200 import synthetic as S
from libpll_b200.binding import (PLL_ATTRIB_AB_FELSENSTEIN, PLL_ATTRIB_AB_FLAG, PLL_ATTRIB_AB_LEWIS,
                                 PLL_ATTRIB_AB_STAMATAKIS, PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_ARCH_CPU,
                                 PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP, PllError)

pytestmark = pytest.mark.gpu
RTOL = 1e-10
TYPES = [0, PLL_ATTRIB_AB_LEWIS, PLL_ATTRIB_AB_FELSENSTEIN, PLL_ATTRIB_AB_STAMATAKIS]


def _pair(gpu_lib, ref_lib, w, extra, ref_arch):
    rates = ref_lib.gamma_rates(w.alpha, w.rate_cats)
    pg, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_AB_FLAG | extra, rates=rates)
    pr, _ = S.build_partition(ref_lib, w, ref_arch | PLL_ATTRIB_AB_FLAG | extra, rates=rates)
    return pg, pr, pidx


# the reference stores ASCII characters in the dummy sites of non-DNA pattern tips (see
# libpll_b200/csrc/host/pll_partition.c), so non-DNA alphabets are compared with CLV tips
@pytest.mark.parametrize("states,cats,pattern_tip,ref_arch", [
    (4, 4, True, PLL_ATTRIB_ARCH_AVX2), (4, 4, False, PLL_ATTRIB_ARCH_AVX2), (4, 1, True, PLL_ATTRIB_ARCH_AVX2),
    (20, 4, False, PLL_ATTRIB_ARCH_AVX2), (5, 3, False, PLL_ATTRIB_ARCH_CPU), (2, 4, False, PLL_ATTRIB_ARCH_CPU)])
def test_ascbias_loglikelihood_and_derivatives(gpu_lib, ref_lib, states, cats, pattern_tip, ref_arch):
    w = S.make_workload(16, 531, states=states, rate_cats=cats, seed=40 + states)
    extra = PLL_ATTRIB_PATTERN_TIP if pattern_tip else 0
    pg, pr, pidx = _pair(gpu_lib, ref_lib, w, extra, ref_arch)
    state_weights = np.arange(3, 3 + states, dtype=np.uint32) * 7
    top = w.tips + w.inner - 1
    last = w.ops[-1]
    tip_edge = None
    for op in w.ops:  # an edge inner -> tip that is valid after the full traversal
        if int(op["child1_clv_index"]) < w.tips:
            tip_edge = (int(op["parent_clv_index"]), int(op["child1_clv_index"]), int(op["child1_matrix_index"]))
    for p in (pg, pr):
        p.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
        p.update_partials(w.ops)
    for ab in TYPES:
        for p in (pg, pr):
            p.set_asc_bias_type(ab)
            if ab in (PLL_ATTRIB_AB_FELSENSTEIN, PLL_ATTRIB_AB_STAMATAKIS):
                p.set_asc_state_weights(state_weights)
        args = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
        lg, lr = pg.edge_loglikelihood(*args), pr.edge_loglikelihood(*args)
        assert np.isfinite(lr) and abs(lg - lr) <= RTOL * abs(lr), (ab, lg, lr)
        rg = pg.root_loglikelihood(top, w.scaler_of(top), pidx)
        rr = pr.root_loglikelihood(top, w.scaler_of(top), pidx)
        assert abs(rg - rr) <= RTOL * abs(rr), (ab, rg, rr)
        edges = [(w.root_a, w.root_b)]
        if tip_edge:
            a, b, m = tip_edge
            tg = pg.edge_loglikelihood(a, w.scaler_of(a), b, w.scaler_of(b), m, pidx)
            tr = pr.edge_loglikelihood(a, w.scaler_of(a), b, w.scaler_of(b), m, pidx)
            assert abs(tg - tr) <= RTOL * abs(tr), (ab, tg, tr)
            edges.append((a, b))
        for (a, b) in edges:
            sg, sr = pg.new_sumtable(), pr.new_sumtable()
            pg.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, sg)
            pr.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, sr)
            for t in (0.01, 0.2, 1.5):
                dg = pg.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), t, pidx, sg)
                dr = pr.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), t, pidx, sr)
                scale = max(abs(dr[0]), float(w.weights.sum()) * 1e-3)
                assert abs(dg[0] - dr[0]) <= RTOL * scale, (ab, a, b, t, dg, dr)
                assert abs(dg[1] - dr[1]) <= RTOL * max(abs(dr[1]), scale), (ab, a, b, t, dg, dr)
    pg.destroy()
    pr.destroy()


def test_ascbias_with_scaling(gpu_lib, ref_lib):
    """long caterpillar: the per-state sites are rescaled too, their scaler counts enter the
    correction as powers of 2^-256"""
    from test_parity_gpu import _caterpillar

    w = _caterpillar(300, 40, 4, seed=3)
    pg, pr, pidx = _pair(gpu_lib, ref_lib, w, PLL_ATTRIB_PATTERN_TIP, PLL_ATTRIB_ARCH_AVX2)
    for p in (pg, pr):
        p.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
        p.update_partials(w.ops)
    total = sum(int(pr.get_scaler(k).sum()) for k in range(w.inner))
    assert total > 0
    for ab in TYPES:
        for p in (pg, pr):
            p.set_asc_bias_type(ab)
            p.set_asc_state_weights([5, 6, 7, 8])
        args = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
        lg, lr = pg.edge_loglikelihood(*args), pr.edge_loglikelihood(*args)
        assert np.isfinite(lr) and abs(lg - lr) <= RTOL * abs(lr), (ab, lg, lr)
    pg.destroy()
    pr.destroy()


@pytest.mark.parametrize("pattern_tip,slices,sites", [(True, 1, 42), (False, 1, 42), (True, 3, 150)])
def test_ascbias_with_per_rate_scalers(gpu_lib, ref_lib, monkeypatch, pattern_tip, slices, sites):
    """PLL_ATTRIB_RATE_SCALERS with the correction: the reference reads element `sites + n` of a
    scaler array laid out [site][rate] (src/likelihood.c:91, :378-381, src/core_derivatives.c:684-685),
    i.e. the count of pattern (sites + n) / R at rate (sites + n) % R; the drop-in returns the same
    values.  Long caterpillar so that those counts are not zero."""
    from libpll_b200.binding import PLL_ATTRIB_RATE_SCALERS
    from test_parity_gpu import _caterpillar

    if slices > 1:  # the flat elements then come from whichever pattern slices hold them
        monkeypatch.setenv("PLL_GPU_DEVICES", str(slices))
    w = _caterpillar(300, sites, 4, seed=5)
    extra = PLL_ATTRIB_RATE_SCALERS | (PLL_ATTRIB_PATTERN_TIP if pattern_tip else 0)
    pg, pr, pidx = _pair(gpu_lib, ref_lib, w, extra, PLL_ATTRIB_ARCH_AVX2)
    assert gpu_lib.pll_gpu_partition_devices(pg.ptr) == slices
    for p in (pg, pr):
        p.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
        p.update_partials(w.ops)
    top = w.tips + w.inner - 1
    picked = 0  # the flat elements sites .. sites + states - 1 the reference reads
    for k in range(w.inner):
        sr, sg = np.asarray(pr.get_scaler(k)).ravel(), np.asarray(pg.get_scaler(k)).ravel()
        assert np.array_equal(sr, sg)
        picked += int(sr[w.sites:w.sites + 4].sum())
    assert picked > 0
    a, b = w.root_a, w.root_b
    for ab in TYPES:
        for p in (pg, pr):
            p.set_asc_bias_type(ab)
            p.set_asc_state_weights([5, 6, 7, 8])
        args = (a, w.scaler_of(a), b, w.scaler_of(b), w.root_matrix, pidx)
        lg, lr = pg.edge_loglikelihood(*args), pr.edge_loglikelihood(*args)
        assert np.isfinite(lr) and abs(lg - lr) <= RTOL * abs(lr), (ab, lg, lr)
        rg = pg.root_loglikelihood(top, w.scaler_of(top), pidx)
        rr = pr.root_loglikelihood(top, w.scaler_of(top), pidx)
        assert np.isfinite(rr) and abs(rg - rr) <= RTOL * abs(rr), (ab, rg, rr)
        sg, sr = pg.new_sumtable(), pr.new_sumtable()
        pg.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, sg)
        pr.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, sr)
        for t in (0.01, 0.2, 1.5):
            dg = pg.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), t, pidx, sg)
            dr = pr.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), t, pidx, sr)
            scale = max(abs(dr[0]), float(w.weights.sum()) * 1e-3)
            assert abs(dg[0] - dr[0]) <= RTOL * scale, (ab, t, dg, dr)
            assert abs(dg[1] - dr[1]) <= RTOL * max(abs(dr[1]), scale), (ab, t, dg, dr)
    pg.destroy()
    pr.destroy()


def test_ascbias_errors(gpu_lib):
    w = S.make_workload(6, 50, states=4, seed=1)
    part, _ = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
    with pytest.raises(PllError):  # not created with AB storage (PLL_ERROR_AB_NOSUPPORT)
        part.set_asc_bias_type(PLL_ATTRIB_AB_LEWIS)
    part.destroy()
    part, _ = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP | PLL_ATTRIB_AB_FLAG)
    with pytest.raises(PllError):  # not a valid type
        part.set_asc_bias_type(1)
    part.set_asc_bias_type(PLL_ATTRIB_AB_LEWIS)
    with pytest.raises(PllError):  # p-inv is incompatible with the correction
        part.update_invariant_sites_proportion(0, 0.2)
    part.destroy()


@pytest.mark.parametrize("slices,sites", [(2, 128), (3, 200), (4, 61)])
def test_ascbias_on_a_sliced_partition(gpu_lib, ref_lib, monkeypatch, slices, sites):
    """pll_gpu_set_devices / PLL_GPU_DEVICES with the correction: the per-state sites behind the
    last pattern fall into the last slice (with 128 patterns in two slices they ARE the last slice:
    a context without a single active pattern), the reductions cover each slice's share of the real
    patterns, the host epilogue fetches the per-state CLVs / scalers / sumtable rows from wherever
    they live.  Same values as the reference (src/likelihood.c:24-119,
    src/core_derivatives.c:654-727)."""
    monkeypatch.setenv("PLL_GPU_DEVICES", str(slices))
    w = S.make_workload(9, sites, states=4, seed=sites)
    pg, pr, pidx = _pair(gpu_lib, ref_lib, w, PLL_ATTRIB_PATTERN_TIP, PLL_ATTRIB_ARCH_AVX2)
    assert gpu_lib.pll_gpu_partition_devices(pg.ptr) >= 2  # 64-pattern aligned slices of sites + states
    state_weights = np.arange(2, 6, dtype=np.uint32) * 5
    top = w.tips + w.inner - 1
    for p in (pg, pr):
        p.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
        p.update_partials(w.ops)
    a, b = w.root_a, w.root_b
    for ab in TYPES:
        for p in (pg, pr):
            p.set_asc_bias_type(ab)
            if ab in (PLL_ATTRIB_AB_FELSENSTEIN, PLL_ATTRIB_AB_STAMATAKIS):
                p.set_asc_state_weights(state_weights)
        args = (a, w.scaler_of(a), b, w.scaler_of(b), w.root_matrix, pidx)
        lg, lr = pg.edge_loglikelihood(*args), pr.edge_loglikelihood(*args)
        assert np.isfinite(lr) and abs(lg - lr) <= RTOL * abs(lr), (ab, lg, lr)
        rg = pg.root_loglikelihood(top, w.scaler_of(top), pidx)
        rr = pr.root_loglikelihood(top, w.scaler_of(top), pidx)
        assert abs(rg - rr) <= RTOL * abs(rr), (ab, rg, rr)
        sg, sr = pg.new_sumtable(), pr.new_sumtable()
        pg.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, sg)
        pr.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, sr)
        for t in (0.05, 0.7):
            dg = pg.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), t, pidx, sg)
            dr = pr.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), t, pidx, sr)
            scale = max(abs(dr[0]), float(w.weights.sum()) * 1e-3)
            assert abs(dg[0] - dr[0]) <= RTOL * scale, (ab, t, dg, dr)
            assert abs(dg[1] - dr[1]) <= RTOL * max(abs(dr[1]), scale), (ab, t, dg, dr)
    pg.destroy()
    pr.destroy()
