"""GPU: one pll_partition_t spread over several device contexts (pll_gpu_set_devices,
libpll_b200/csrc/host/pll_devices.c).

Slices are placed round-robin over the visible devices, so the slicing logic is exercised on a
single GPU too (several contexts on one device); with two or more GPUs the same tests run across
devices.  A sliced partition must be indistinguishable from an unsliced one: CLVs, scalers,
invariant sites and per-pattern log-likelihoods bit for bit (patterns are independent), totals and
derivatives up to the order of the final sum (<= 1e-12 relative), and both within the parity
tolerance of the reference.
"""
import numpy as np
import pytest

from libpll_b200 import synthetic as S
from libpll_b200.binding import (PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP,
                                 PLL_ATTRIB_RATE_SCALERS)

pytestmark = pytest.mark.gpu
RTOL = 1e-10   # north_star tolerance against the reference


@pytest.fixture(autouse=True, params=["host-sum", "device-reduce"])
def reduction_mode(request, monkeypatch):
    """Both ways of combining the per-slice scalars: every device writes its own mapped host words
    and the host layer adds them (default), or the devices add them among themselves through
    peer-mapped slots on the first device (plg_group_*, PLL_GPU_DEVICE_REDUCE=1)."""
    monkeypatch.setenv("PLL_GPU_DEVICE_REDUCE", "1" if request.param == "device-reduce" else "0")
    return request.param


def evaluate(lib, w, attributes, slices, pinv=0.0):
    assert lib.pll_gpu_set_devices(slices) == 1
    try:
        part, pidx = S.build_partition(lib, w, attributes)
    finally:
        lib.pll_gpu_set_devices(0)
    out = {"devices": lib.pll_gpu_partition_devices(part.ptr)}
    if pinv:
        part.update_invariant_sites()
        out["invariant"] = np.ctypeslib.as_array(part.p.invariant, shape=(w.sites,)).copy()
        for i in set(int(x) for x in pidx):
            part.update_invariant_sites_proportion(i, pinv)
    out["lnl"] = S.full_evaluation(part, w, pidx)
    persite = np.zeros(w.sites)
    a, b = w.root_a, w.root_b
    args = (a, w.scaler_of(a), b, w.scaler_of(b), w.root_matrix, pidx)
    assert part.edge_loglikelihood(*args, persite=persite) == out["lnl"]
    out["persite"] = persite
    out["clv"] = part.get_clv(a if a >= w.tips else b)
    out["scalers"] = [part.get_scaler(k) for k in range(w.inner)]
    tab = part.new_sumtable()
    part.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, tab)
    out["derivs"] = [part.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), t, pidx, tab)
                     for t in (0.05, 0.3)]
    root_clv = a if a >= w.tips else b
    out["root"] = part.root_loglikelihood(root_clv, w.scaler_of(root_clv), pidx)
    part.destroy()
    return out


@pytest.mark.parametrize("states,tips,sites,slices,attrs,pinv", [
    (4, 40, 5000, 2, PLL_ATTRIB_PATTERN_TIP, 0.0),
    (4, 40, 5000, 3, PLL_ATTRIB_PATTERN_TIP, 0.2),
    (4, 150, 1000, 4, PLL_ATTRIB_PATTERN_TIP | PLL_ATTRIB_RATE_SCALERS, 0.0),
    (4, 24, 777, 5, 0, 0.0),
    (20, 12, 700, 2, PLL_ATTRIB_PATTERN_TIP, 0.1),
    (20, 10, 300, 3, 0, 0.0),
    (7, 12, 500, 2, PLL_ATTRIB_PATTERN_TIP, 0.0),
])
def test_sliced_partition_equals_unsliced(gpu_lib, ref_lib, states, tips, sites, slices, attrs, pinv):
    w = S.make_workload(tips, sites, states=states, seed=100 + slices)
    one = evaluate(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | attrs, 1, pinv)
    many = evaluate(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | attrs, slices, pinv)
    assert one["devices"] == 1
    assert many["devices"] == min(slices, -(-sites // 64))
    assert many["devices"] > 1
    assert np.array_equal(one["persite"], many["persite"])
    assert one["clv"].tobytes() == many["clv"].tobytes()
    for k, (x, y) in enumerate(zip(one["scalers"], many["scalers"])):
        assert np.array_equal(x, y), f"scaler {k}"
    if pinv:
        assert np.array_equal(one["invariant"], many["invariant"])
    assert abs(one["lnl"] - many["lnl"]) <= 1e-12 * abs(one["lnl"])
    if not attrs & PLL_ATTRIB_RATE_SCALERS:
        # the reference's root kernels index a per-rate scale buffer per SITE
        # (src/core_likelihood_avx.c:176-178, reproduced): entry n of the flattened [site][rate]
        # array, which is a different entry inside a slice - rooted evaluation with per-rate
        # scalers is not meaningful in the reference either
        assert abs(one["root"] - many["root"]) <= 1e-12 * abs(one["root"])
    for (a1, a2), (b1, b2) in zip(one["derivs"], many["derivs"]):
        assert abs(a1 - b1) <= 1e-11 * max(abs(a1), 1.0)
        assert abs(a2 - b2) <= 1e-11 * max(abs(a2), 1.0)

    # and against the reference itself (its generic kernels for 7 states are the plain-C path)
    if states in (4, 20) and not (attrs & PLL_ATTRIB_RATE_SCALERS and pinv):
        pr, pidx = S.build_partition(ref_lib, w, PLL_ATTRIB_ARCH_AVX2 | attrs)
        if pinv:
            pr.update_invariant_sites()
            for i in set(int(x) for x in pidx):
                pr.update_invariant_sites_proportion(i, pinv)
        want = S.full_evaluation(pr, w, pidx)
        pr.destroy()
        assert abs(many["lnl"] - want) <= RTOL * abs(want)


def test_slices_follow_the_environment_and_reject_asc_bias(gpu_lib, monkeypatch):
    from libpll_b200.binding import PLL_ATTRIB_AB_LEWIS
    w = S.make_workload(8, 400, states=4, seed=5)
    assert gpu_lib.pll_gpu_set_devices(0) == 1      # 0 = default: the environment decides
    monkeypatch.setenv("PLL_GPU_DEVICES", "3")
    part, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
    assert gpu_lib.pll_gpu_partition_devices(part.ptr) == 3
    first, count = np.zeros(1, np.uint32), np.zeros(1, np.uint32)
    import ctypes as C
    u = C.POINTER(C.c_uint)
    covered = 0
    for d in range(3):
        assert gpu_lib.pll_gpu_context_of(part.ptr, d, first.ctypes.data_as(u), count.ctypes.data_as(u))
        assert int(first[0]) == covered and int(first[0]) % 64 == 0
        covered += int(count[0])
    assert covered == 400
    assert not gpu_lib.pll_gpu_context_of(part.ptr, 3, None, None)
    lnl3 = S.full_evaluation(part, w, pidx)
    part.destroy()
    monkeypatch.delenv("PLL_GPU_DEVICES")
    part, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
    assert gpu_lib.pll_gpu_partition_devices(part.ptr) == 1
    lnl1 = S.full_evaluation(part, w, pidx)
    part.destroy()
    assert abs(lnl1 - lnl3) <= 1e-12 * abs(lnl1)

    assert gpu_lib.pll_gpu_set_devices(-1) == 0 and gpu_lib.pll_gpu_set_devices(17) == 0


@pytest.mark.parametrize("states,tips,sites", [(4, 300, 150), (4, 300, 192), (20, 140, 150)])
def test_root_loglikelihood_of_a_sliced_partition_with_per_rate_scalers(gpu_lib, ref_lib, states, tips, sites):
    """The reference's root kernels read element n of the per-rate scaler array [site][rate] for
    pattern n (src/core_likelihood_avx.c:176-178, src/core_likelihood.c:197-198) - the count of pattern
    n / R at rate n % R.  On a sliced partition those elements live in an earlier slice; the host
    layer gathers them (pllg_dev_root_loglikelihood).  A long caterpillar, so that the counts are not
    zero: root and edge log-likelihood and the per-pattern values against the reference."""
    from test_parity_gpu import _caterpillar

    w = _caterpillar(tips, sites, states, seed=5)
    rates = ref_lib.gamma_rates(w.alpha, w.rate_cats)
    extra = PLL_ATTRIB_PATTERN_TIP | PLL_ATTRIB_RATE_SCALERS
    pr, pidx = S.build_partition(ref_lib, w, PLL_ATTRIB_ARCH_AVX2 | extra, rates=rates)
    assert gpu_lib.pll_gpu_set_devices(3) == 1
    try:
        pg, _ = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | extra, rates=rates)
    finally:
        gpu_lib.pll_gpu_set_devices(0)
    assert gpu_lib.pll_gpu_partition_devices(pg.ptr) == 3
    for p in (pg, pr):
        p.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
        p.update_partials(w.ops)
    top = w.tips + w.inner - 1
    counts = np.asarray(pr.get_scaler(w.scaler_of(top))).ravel()
    assert np.array_equal(counts, np.asarray(pg.get_scaler(w.scaler_of(top))).ravel())
    assert counts[:sites].any() and len(set(counts[:sites].tolist())) > 1   # the flat elements differ
    sg, sr = np.zeros(sites), np.zeros(sites)
    rg = pg.root_loglikelihood(top, w.scaler_of(top), pidx, persite=sg)
    rr = pr.root_loglikelihood(top, w.scaler_of(top), pidx, persite=sr)
    assert np.isfinite(rr) and abs(rg - rr) <= RTOL * abs(rr), (rg, rr)
    assert np.allclose(sg, sr, rtol=RTOL, atol=0)
    a, b = w.root_a, w.root_b
    args = (a, w.scaler_of(a), b, w.scaler_of(b), w.root_matrix, pidx)
    eg, er = pg.edge_loglikelihood(*args), pr.edge_loglikelihood(*args)
    assert abs(eg - er) <= RTOL * abs(er), (eg, er)
    pg.destroy()
    pr.destroy()
