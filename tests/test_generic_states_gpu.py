"""GPU parity for alphabets / category counts outside the specialised kernels (anything but 4 or
20 states with 1/2/4/8/16 categories): the generic device path (libpll_b200/csrc/gpu/plg_generic.cu)
against the reference's own plain-C path (oracle/_ref with PLL_ATTRIB_ARCH_CPU) on the same seeded
inputs.

Why ARCH_CPU and not AVX2 here: with a state count that is not a multiple of 4 the reference's
generic AVX kernels walk 4 matrix rows at a time over states_padded rows, i.e. they read rows
that belong to the NEXT rate category's matrix and leave those products in the padding lanes of
the parent CLV (reference src/core_partials_avx.c:1400-1560, the `displacement` correction);
the rescaling test then looks at those lanes too, so its scaler counts depend on memory that is
not part of the model.  The likelihoods agree either way (the golden cases recorded from the
AVX2 path pass, tests/test_golden_gpu.py); scaler counts are pinned against the plain-C path,
which is the reference's definition of the computation.

Bars: scaler counts bit-exact; CLVs bit-exact when both sides are given the same P-matrices;
per-site / total lnL, d_f, dd_f relative 1e-10."""
import numpy as np
import pytest

from libpll_b200 import synthetic as S
from libpll_b200.binding import (PLL_ATTRIB_ARCH_CPU, PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP,
                                 PLL_ATTRIB_RATE_SCALERS)
from test_parity_gpu import _caterpillar

pytestmark = pytest.mark.gpu

RTOL = 1e-10
CONFIGS = [(2, 4), (3, 3), (5, 4), (7, 5), (11, 2), (16, 4), (32, 1), (4, 3), (20, 5), (4, 25)]


def _pair(gpu_lib, ref_lib, w, extra=0, ref_extra=None):
    rates = ref_lib.gamma_rates(w.alpha, w.rate_cats)
    pg, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | extra, rates=rates)
    pr, _ = S.build_partition(ref_lib, w, PLL_ATTRIB_ARCH_CPU | (extra if ref_extra is None else ref_extra),
                              rates=rates)
    return pg, pr, pidx


def _share_pmatrices(pg, pr, w, pidx):
    """Both sides compute their P-matrices (compared at 1e-10), then the reference's are pushed
    to the device (padded to its row pitch) so that CLVs can be compared bit for bit."""
    K = w.states
    pg.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    pr.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    for m in range(w.prob_matrices):
        a, b = pg.get_pmatrix(m), pr.get_pmatrix(m)
        np.testing.assert_allclose(a[..., :K], b[..., :K], rtol=RTOL, atol=1e-15)
        assert not a[..., K:].any(), "padding columns must stay zero"
        padded = np.zeros_like(a)
        padded[..., :K] = b[..., :K]
        pg.set_pmatrix(m, padded)


@pytest.mark.parametrize("states,cats", CONFIGS)
@pytest.mark.parametrize("pattern_tip", [True, False])
def test_generic_traversal(gpu_lib, ref_lib, states, cats, pattern_tip):
    tips, sites = 14, 1037
    w = S.make_workload(tips, sites, states=states, rate_cats=cats, seed=100 + states)
    pg, pr, pidx = _pair(gpu_lib, ref_lib, w, PLL_ATTRIB_PATTERN_TIP if pattern_tip else 0)
    _share_pmatrices(pg, pr, w, pidx)
    pg.update_partials(w.ops)
    pr.update_partials(w.ops)
    for k in range(w.inner):
        np.testing.assert_array_equal(pg.get_scaler(k), pr.get_scaler(k), err_msg=f"scaler {k}")
        a, b = pg.get_clv(w.tips + k), pr.get_clv(w.tips + k)
        assert a[..., :states].tobytes() == b[..., :states].tobytes(), \
            f"CLV {k} differs (max abs {np.max(np.abs(a[..., :states] - b[..., :states]))})"
        assert not a[..., states:].any(), "padding states must stay zero"
    ps_g, ps_r = np.zeros(sites), np.zeros(sites)
    args = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
    lg = pg.edge_loglikelihood(*args, persite=ps_g)
    lr = pr.edge_loglikelihood(*args, persite=ps_r)
    np.testing.assert_allclose(ps_g, ps_r, rtol=RTOL, atol=0)
    assert abs(lg - lr) <= RTOL * abs(lr)
    top = w.tips + w.inner - 1
    rg = pg.root_loglikelihood(top, w.scaler_of(top), pidx)
    rr = pr.root_loglikelihood(top, w.scaler_of(top), pidx)
    assert abs(rg - rr) <= RTOL * abs(rr), (rg, rr)
    pg.destroy()
    pr.destroy()


@pytest.mark.parametrize("states,cats", [(5, 4), (7, 3), (4, 3)])
@pytest.mark.parametrize("rate_scalers", [False, True])
def test_generic_scaling_long_tree(gpu_lib, ref_lib, states, cats, rate_scalers):
    w = _caterpillar(260, 41, states, seed=9)
    w.rate_cats = cats
    extra = PLL_ATTRIB_PATTERN_TIP | (PLL_ATTRIB_RATE_SCALERS if rate_scalers else 0)
    # the reference's plain-C tip-inner kernel has no per-rate branch (it bumps parent_scaler[n]
    # in a per-rate array, reference src/core_partials.c:461-510), so with per-rate scalers its
    # side runs tips as CLVs: the inner-inner kernel (reference src/core_partials.c:604-662)
    # implements per-rate scaling and yields the same numbers as pattern tips
    ref_extra = PLL_ATTRIB_RATE_SCALERS if rate_scalers and states != 4 else None
    pg, pr, pidx = _pair(gpu_lib, ref_lib, w, extra, ref_extra)
    _share_pmatrices(pg, pr, w, pidx)
    pg.update_partials(w.ops)
    pr.update_partials(w.ops)
    total = 0
    for k in range(w.inner):
        a, b = pg.get_scaler(k), pr.get_scaler(k)
        np.testing.assert_array_equal(a, b, err_msg=f"scaler {k}")
        total += int(b.sum())
    assert total > 0, "the test tree must actually trigger rescaling"
    args = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
    lg, lr = pg.edge_loglikelihood(*args), pr.edge_loglikelihood(*args)
    assert np.isfinite(lr)
    assert abs(lg - lr) <= RTOL * abs(lr), (lg, lr)
    pg.destroy()
    pr.destroy()


@pytest.mark.parametrize("states,cats", [(5, 4), (7, 3), (13, 2), (4, 5)])
@pytest.mark.parametrize("pinv", [0.0, 0.3])
def test_generic_derivatives(gpu_lib, ref_lib, states, cats, pinv):
    w = S.make_workload(12, 701, states=states, rate_cats=cats, seed=31)
    pg, pr, pidx = _pair(gpu_lib, ref_lib, w, PLL_ATTRIB_PATTERN_TIP)
    if pinv > 0:
        for p in (pg, pr):
            p.update_invariant_sites_proportion(0, pinv)
        np.testing.assert_array_equal(pg.get_invariant(), pr.get_invariant())
    _share_pmatrices(pg, pr, w, pidx)
    pg.update_partials(w.ops)
    pr.update_partials(w.ops)
    edges = [(w.root_a, w.root_b)]
    last = w.ops[-1]
    edges.append((int(last["parent_clv_index"]), int(last["child1_clv_index"])))
    first = w.ops[0]  # an edge down to a tip, if the first operation has one
    if int(first["child1_clv_index"]) < w.tips:
        edges.append((int(first["parent_clv_index"]), int(first["child1_clv_index"])))
    for (a, b) in edges:
        sg, sr = pg.new_sumtable(), pr.new_sumtable()
        pg.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, sg)
        pr.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, sr)
        for t in (0.001, 0.05, 0.3, 2.0):
            dg = pg.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), t, pidx, sg)
            dr = pr.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), t, pidx, sr)
            scale = max(abs(dr[0]), float(w.weights.sum()) * 1e-3)
            assert abs(dg[0] - dr[0]) <= RTOL * scale, (a, b, t, dg, dr)
            assert abs(dg[1] - dr[1]) <= RTOL * max(abs(dr[1]), scale), (a, b, t, dg, dr)
    pg.destroy()
    pr.destroy()
