"""GPU parity tests: the CUDA path (through the pll.h C-ABI) against the reference's own
AVX2 path (oracle/_ref) on the same seeded inputs.

Bars (BASELINE.json north_star):
  * scaler counts                 bit-exact
  * CLVs                          bit-exact when both sides are given the same P-matrices
  * per-site / total lnL, d_f, dd_f   relative 1e-10
"""
import numpy as np
import pytest

from libpll_b200 import synthetic as S
from libpll_b200.binding import (
    PLL_ATTRIB_ARCH_AVX2,
    PLL_ATTRIB_ARCH_GPU,
    PLL_ATTRIB_PATTERN_TIP,
    PLL_ATTRIB_RATE_SCALERS,
)

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _pair(gpu_lib, ref_lib, w, extra=0, variant="default"):
    rates = ref_lib.gamma_rates(w.alpha, w.rate_cats)
    pg, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | extra, variant=variant, rates=rates)
    pr, _ = S.build_partition(ref_lib, w, PLL_ATTRIB_ARCH_AVX2 | extra, variant=variant, rates=rates)
    return pg, pr, pidx


def _share_pmatrices(pg, pr, w, pidx):
    """Both sides compute their P-matrices; the reference's are then pushed to the device so
    that everything downstream can be compared bit for bit (isolates expm1)."""
    pg.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    pr.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    worst = 0.0
    for m in range(w.prob_matrices):
        a, b = pg.get_pmatrix(m), pr.get_pmatrix(m)
        worst = max(worst, float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))))
        pg.set_pmatrix(m, b)
    return worst


@pytest.fixture(params=["exact", "dmma"])
def aa_mode(request, monkeypatch):
    """20-state CLV updates have two device implementations: the vector-pipe kernels that follow
    the reference's AVX2 lane order (bit-exact CLVs) and the default DMMA tensor-core kernels
    (CLVs to ~1e-16 relative).  The library reads PLL_GPU_AA_EXACT at partition creation."""
    monkeypatch.setenv("PLL_GPU_AA_EXACT", "1" if request.param == "exact" else "0")
    return request.param


def _assert_clv(a, b, exact, what):
    if exact:
        assert a.tobytes() == b.tobytes(), f"{what} differs (max abs {np.max(np.abs(a - b))})"
    else:
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=0, err_msg=what)


@pytest.mark.parametrize("states,tips,sites", [(4, 12, 1000), (4, 60, 4099), (20, 10, 777), (20, 40, 1031)])
@pytest.mark.parametrize("pattern_tip", [True, False])
def test_traversal_bit_exact(gpu_lib, ref_lib, states, tips, sites, pattern_tip, aa_mode):
    if states == 4 and aa_mode == "dmma":
        pytest.skip("DNA has a single implementation")
    exact = states == 4 or aa_mode == "exact"
    w = S.make_workload(tips, sites, states=states, seed=7 + tips)
    extra = PLL_ATTRIB_PATTERN_TIP if pattern_tip else 0
    pg, pr, pidx = _pair(gpu_lib, ref_lib, w, extra)
    worst = _share_pmatrices(pg, pr, w, pidx)
    assert worst < 1e-13, f"P-matrix relative difference {worst}"

    pg.update_partials(w.ops)
    pr.update_partials(w.ops)
    for k in range(w.inner):
        np.testing.assert_array_equal(pg.get_scaler(k), pr.get_scaler(k), err_msg=f"scaler {k}")
        _assert_clv(pg.get_clv(w.tips + k), pr.get_clv(w.tips + k), exact, f"CLV {w.tips + k}")

    ps_g, ps_r = np.zeros(sites), np.zeros(sites)
    args = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
    lg = pg.edge_loglikelihood(*args, persite=ps_g)
    lr = pr.edge_loglikelihood(*args, persite=ps_r)
    np.testing.assert_allclose(ps_g, ps_r, rtol=RTOL, atol=0)
    assert abs(lg - lr) <= RTOL * abs(lr)
    pg.destroy()
    pr.destroy()


@pytest.mark.parametrize("states,tips,sites", [(4, 30, 2000), (20, 16, 600)])
def test_end_to_end_device_pmatrices(gpu_lib, ref_lib, states, tips, sites):
    """No injection: device expm1 P-matrices all the way.  Scalers must still be identical,
    lnL within 1e-10."""
    w = S.make_workload(tips, sites, states=states, seed=99)
    pg, pr, pidx = _pair(gpu_lib, ref_lib, w, PLL_ATTRIB_PATTERN_TIP)
    lg = S.full_evaluation(pg, w, pidx)
    lr = S.full_evaluation(pr, w, pidx)
    for k in range(w.inner):
        np.testing.assert_array_equal(pg.get_scaler(k), pr.get_scaler(k))
    assert abs(lg - lr) <= RTOL * abs(lr), (lg, lr)
    # root log-likelihood at the last inner CLV as well
    top = w.tips + w.inner - 1
    rg = pg.root_loglikelihood(top, w.scaler_of(top), pidx)
    rr = pr.root_loglikelihood(top, w.scaler_of(top), pidx)
    assert abs(rg - rr) <= RTOL * abs(rr), (rg, rr)
    pg.destroy()
    pr.destroy()


def _caterpillar(tips, sites, states, seed):
    """Ladder tree with alternating long / tiny branches: forces repeated rescaling
    (the recipe of reference test/src/scaling.c:237-238 on a synthetic tree)."""
    w = S.make_workload(tips, sites, states=states, seed=seed)
    ops = np.zeros(tips - 2, dtype=S.OP_DTYPE)
    prev = 0
    for k in range(tips - 2):
        parent = tips + k
        ops[k] = (parent, k, prev, prev, w.scaler_of(prev), k + 1, k + 1, w.scaler_of(k + 1))
        prev = parent
    w.ops = ops
    w.root_a, w.root_b, w.root_matrix = prev, tips - 1, tips - 1
    w.branch_lengths = np.where(np.arange(w.prob_matrices) % 2 == 0, 1.0, 1e-6)
    return w


@pytest.mark.parametrize("states,rate_scalers", [(4, False), (4, True), (20, False), (20, True)])
def test_scaling_long_tree(gpu_lib, ref_lib, states, rate_scalers, aa_mode):
    if states == 4 and aa_mode == "dmma":
        pytest.skip("DNA has a single implementation")
    exact = states == 4 or aa_mode == "exact"
    tips, sites = (400, 64) if states == 4 else (160, 37)
    w = _caterpillar(tips, sites, states, seed=5)
    extra = PLL_ATTRIB_PATTERN_TIP | (PLL_ATTRIB_RATE_SCALERS if rate_scalers else 0)
    pg, pr, pidx = _pair(gpu_lib, ref_lib, w, extra)
    _share_pmatrices(pg, pr, w, pidx)
    pg.update_partials(w.ops)
    pr.update_partials(w.ops)
    total = 0
    for k in range(w.inner):
        a, b = pg.get_scaler(k), pr.get_scaler(k)
        np.testing.assert_array_equal(a, b, err_msg=f"scaler {k}")
        total += int(b.sum())
    assert total > 0, "the test tree must actually trigger rescaling"
    top = w.root_a
    _assert_clv(pg.get_clv(top), pr.get_clv(top), exact, "top CLV")
    args = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
    lg, lr = pg.edge_loglikelihood(*args), pr.edge_loglikelihood(*args)
    assert np.isfinite(lr)
    assert abs(lg - lr) <= RTOL * abs(lr), (lg, lr)
    pg.destroy()
    pr.destroy()


@pytest.mark.parametrize("states", [4, 20])
@pytest.mark.parametrize("pinv", [0.0, 0.3])
def test_derivatives(gpu_lib, ref_lib, states, pinv):
    tips, sites = (24, 3001) if states == 4 else (12, 501)
    w = S.make_workload(tips, sites, states=states, seed=21)
    pg, pr, pidx = _pair(gpu_lib, ref_lib, w, PLL_ATTRIB_PATTERN_TIP)
    if pinv > 0:
        for p in (pg, pr):
            p.update_invariant_sites_proportion(0, pinv)
        np.testing.assert_array_equal(pg.get_invariant(), pr.get_invariant())
    _share_pmatrices(pg, pr, w, pidx)
    pg.update_partials(w.ops)
    pr.update_partials(w.ops)
    # inner-inner edge (root edge if b is inner) and a tip-inner edge
    edges = [(w.root_a, w.root_b)]
    last = w.ops[-1]
    edges.append((int(last["parent_clv_index"]), int(last["child1_clv_index"])))
    for (a, b) in edges:
        sg, sr = pg.new_sumtable(), pr.new_sumtable()
        pg.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, sg)
        pr.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, sr)
        for t in (0.001, 0.05, 0.3, 2.0):
            dg = pg.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), t, pidx, sg)
            dr = pr.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), t, pidx, sr)
            # d_f is a sum of mixed-sign terms: tolerance relative to sum |w| * |term| ~ sites
            scale = max(abs(dr[0]), float(w.weights.sum()) * 1e-3)
            assert abs(dg[0] - dr[0]) <= RTOL * scale, (a, b, t, dg, dr)
            assert abs(dg[1] - dr[1]) <= RTOL * max(abs(dr[1]), scale), (a, b, t, dg, dr)
    pg.destroy()
    pr.destroy()


def test_edge_lnl_with_invariant_sites(gpu_lib, ref_lib):
    w = S.make_workload(20, 1500, states=4, seed=3)
    pg, pr, pidx = _pair(gpu_lib, ref_lib, w, PLL_ATTRIB_PATTERN_TIP)
    for p in (pg, pr):
        p.update_invariant_sites_proportion(0, 0.25)
    lg = S.full_evaluation(pg, w, pidx)
    lr = S.full_evaluation(pr, w, pidx)
    assert abs(lg - lr) <= RTOL * abs(lr), (lg, lr)
    pg.destroy()
    pr.destroy()


@pytest.mark.parametrize("states,tips,sites,slots", [(4, 150, 1500, 10), (4, 64, 700, 7), (20, 60, 300, 8)])
def test_slot_recycling_hazards(gpu_lib, ref_lib, states, tips, sites, slots):
    """CLV / scaler slots recycled by the caller (legal pll.h use, needed for 5,000 x 10M):
    later operations overwrite slots earlier ones read, so the batched scheduler has to honour
    WAR and WAW hazards, not only RAW.  Same list on the reference (strictly sequential)."""
    w = S.recycle_slots(S.make_workload(tips, sites, states=states, seed=23), slots)
    pg, pr, pidx = _pair(gpu_lib, ref_lib, w, PLL_ATTRIB_PATTERN_TIP)
    _share_pmatrices(pg, pr, w, pidx)
    for _ in range(2):  # second call replays the cached CUDA graph
        pg.update_partials(w.ops)
    pr.update_partials(w.ops)
    for k in range(slots):
        np.testing.assert_array_equal(pg.get_scaler(k), pr.get_scaler(k), err_msg=f"scaler slot {k}")
        np.testing.assert_allclose(pg.get_clv(w.tips + k), pr.get_clv(w.tips + k), rtol=1e-12, atol=0)
    args = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
    lg, lr = pg.edge_loglikelihood(*args), pr.edge_loglikelihood(*args)
    assert abs(lg - lr) <= RTOL * abs(lr), (lg, lr)
    pg.destroy()
    pr.destroy()


def test_full_width_properties(gpu_lib, ref_lib):
    """Size-independent properties at a pattern count the CPU oracle cannot cover in seconds
    (500k patterns, 25 GB of CLVs): the per-pattern lnLs add up to the total; the lnL of the
    whole alignment equals the sum over two half partitions (what site sharding relies on);
    a 3,000-pattern window is checked against the reference pattern by pattern; repeated
    evaluations are bit-reproducible."""
    tips, sites = 48, 500_000
    w = S.make_workload(tips, sites, states=4, seed=31)
    pg, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
    total = S.full_evaluation(pg, w, pidx)
    again = S.full_evaluation(pg, w, pidx)
    assert total == again
    persite = np.zeros(sites)
    args = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
    assert pg.edge_loglikelihood(*args, persite=persite) == total
    assert abs(np.sum(persite) - total) <= 1e-11 * abs(total)
    pg.destroy()

    halves = 0.0
    for lo, hi in ((0, 250_048), (250_048, sites)):
        ph, _ = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP, lo=lo, hi=hi)
        halves += S.full_evaluation(ph, w, pidx)
        ph.destroy()
    assert abs(halves - total) <= 1e-11 * abs(total)

    lo, hi = 123_456, 126_456
    pr, _ = S.build_partition(ref_lib, w, PLL_ATTRIB_ARCH_AVX2 | PLL_ATTRIB_PATTERN_TIP, lo=lo, hi=hi)
    S.full_evaluation(pr, w, pidx)
    ps_ref = np.zeros(hi - lo)
    pr.edge_loglikelihood(*args, persite=ps_ref)
    pr.destroy()
    np.testing.assert_allclose(persite[lo:hi], ps_ref, rtol=RTOL, atol=0)


def test_two_devices_in_one_process(gpu_lib):
    """Partitions on two different GPUs of one process (pll_gpu_set_device): per-device kernel
    attributes (opt-in shared memory) must be set on both; same workload, same log-likelihood."""
    if gpu_lib.pll_gpu_device_count() < 2:
        pytest.skip("needs two visible GPUs")
    results = []
    for states, tips, sites in ((4, 40, 5000), (20, 12, 700)):
        w = S.make_workload(tips, sites, states=states, seed=77)
        per_device = []
        for dev in (0, 1):
            gpu_lib.pll_gpu_set_device(dev)
            part, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
            per_device.append(S.full_evaluation(part, w, pidx))
            a, b = w.root_a, w.root_b
            tab = part.new_sumtable()
            part.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, tab)
            per_device.append(part.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), 0.1, pidx, tab))
            part.destroy()
        results.append(per_device)
        assert per_device[0] == per_device[2] and per_device[1] == per_device[3], per_device
    gpu_lib.pll_gpu_set_device(0)
