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


def test_child_clv_and_scaler_from_different_producers(gpu_lib, ref_lib, monkeypatch):
    """Slot recycling lets an operation pair a CLV with a scaler that ANOTHER operation wrote.  W
    below reads CLV c0 (written by X) together with scaler s1 (written by Y); fillers push both out
    of the one-slot tile cache, and later operations overwrite c1 / s1, so Y's stores are dead
    unless the planner notices that W reads s1 back from HBM.  Scaler s1 holds a marker before the
    call: if Y's store of zeros were dropped, W would add the marker to its own count."""
    from libpll_b200.binding import OP_DTYPE, PLL_ATTRIB_ARCH_AVX2
    import ctypes as C

    w = S.make_workload(12, 1300, states=4, seed=31)
    T, NONE = w.tips, -1
    ops = np.array([
        (T + 0, 0, 0, 0, NONE, 1, 1, NONE),            # X: tip-tip -> c0, s0
        (T + 1, 1, 2, 2, NONE, 3, 3, NONE),            # Y: tip-tip -> c1, s1 (zeroes s1)
        (T + 2, 2, 4, 4, NONE, 5, 5, NONE),            # fillers: evict X and Y from the cache
        (T + 3, 3, 6, 6, NONE, T + 2, T + 2, 2),
        (T + 5, 5, 8, 8, NONE, 9, 9, NONE),
        (T + 4, 4, 7, 7, NONE, T + 0, T + 0, 1),       # W: CLV of X with the scaler of Y
        (T + 1, 1, 10, 10, NONE, 11, 11, NONE),        # c1 / s1 overwritten: Y is not their last writer
        (T + 6, 6, T + 1, T + 1, 1, T + 4, T + 4, 4),
    ], dtype=OP_DTYPE)
    mats = np.arange(w.prob_matrices, dtype=np.uint32)
    bl = np.linspace(0.05, 0.6, w.prob_matrices)
    marker = np.full(w.sites, 7, dtype=np.uint32)

    pr, pidx = S.build_partition(ref_lib, w, PLL_ATTRIB_ARCH_AVX2 | PLL_ATTRIB_PATTERN_TIP)
    pr.update_prob_matrices(pidx, mats, bl)
    pr.update_partials(ops)
    want_sc = [pr.get_scaler(k).copy() for k in range(7)]
    want_clv = [pr.get_clv(T + k).copy() for k in range(7)]

    for fused, slots in ((False, 3), (True, 1), (True, 3)):
        monkeypatch.setenv("PLL_GPU_FUSED", "1" if fused else "0")
        monkeypatch.setenv("PLL_GPU_FUSED_SLOTS", str(slots))
        part, _ = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
        part.update_prob_matrices(pidx, mats, bl)
        for m in range(w.prob_matrices):
            part.set_pmatrix(m, pr.get_pmatrix(m))
        for s in range(7):
            assert gpu_lib.plg_set_scaler(part.ctx(), s, marker.ctypes.data_as(C.POINTER(C.c_uint))) == 0
        part.update_partials(ops)
        for k in (0, 1, 2, 3, 4, 5, 6):
            np.testing.assert_array_equal(part.get_scaler(k), want_sc[k], err_msg=f"scaler {k} fused={fused} slots={slots}")
            assert part.get_clv(T + k).tobytes() == want_clv[k].tobytes(), f"CLV {k} fused={fused} slots={slots}"
        part.destroy()
    pr.destroy()


def test_lists_longer_than_the_staging_ring(gpu_lib, monkeypatch):
    """The reference accepts any operation count.  A 45 000-operation list (a small recycled tree
    evaluated over and over with changing matrices) does not fit the 8 MB staging ring: it gets a
    device buffer of its own and gives the same bits as the same list issued in short pieces."""
    w = S.recycle_slots(S.make_workload(16, 640, states=4, seed=13), 6)
    reps = 45000 // len(w.ops) + 1
    chunks = []
    for r in range(reps):
        o = w.ops.copy()
        # rotate the matrices so that repetitions are not identical work
        o["child1_matrix_index"] = (o["child1_matrix_index"] + r) % w.prob_matrices
        o["child2_matrix_index"] = (o["child2_matrix_index"] + 2 * r) % w.prob_matrices
        chunks.append(o)
    long_list = np.concatenate(chunks)
    assert len(long_list) * 128 > 4 << 20
    results = []
    for fused in (True, False):
        monkeypatch.setenv("PLL_GPU_FUSED", "1" if fused else "0")
        for pieces in (False, True):
            part, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
            part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
            if pieces:
                for c in chunks:
                    part.update_partials(c)
            else:
                part.update_partials(long_list)
            results.append(([part.get_clv(w.tips + k).tobytes() for k in range(w.inner)],
                            [part.get_scaler(k).tobytes() for k in range(w.inner)]))
            part.destroy()
    for r in results[1:]:
        assert r == results[0]


def test_graph_cache_evicts_least_recently_used(gpu_lib, monkeypatch):
    monkeypatch.setenv("PLL_GPU_GRAPH_CACHE", "4")
    w = S.make_workload(40, 2000, states=4, seed=8)
    part, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
    part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    part.update_partials(w.ops)
    want = {int(o["parent_clv_index"]): part.get_clv(int(o["parent_clv_index"])).tobytes() for o in w.ops}
    # ten distinct lists (the full traversal minus its first j operations, whose results exist
    # already), each issued three times: capture on the second sighting, replay on the third
    part.reset_stats()
    for j in range(10):
        for _ in range(3):
            part.update_partials(w.ops[j:])
    st = part.stats()
    assert st["graph_evictions"] >= 6, st
    assert st["graph_launches"] >= 20, st
    for j in (0, 9, 3):                     # evicted lists are simply captured again
        for _ in range(2):
            part.update_partials(w.ops[j:])
    for idx, b in want.items():
        assert part.get_clv(idx).tobytes() == b
    part.destroy()


def _memo_tips(monkeypatch):
    cache = {}
    orig = S.tip_sequence

    def memo(w, tip, lo=0, hi=None):
        key = (w.tips, w.sites, w.states, w.seed, tip, lo, hi)
        if key not in cache:
            cache[key] = orig(w, tip, lo, hi)
        return cache[key]

    monkeypatch.setattr(S, "tip_sequence", memo)


def test_miss_path_at_scale(gpu_lib, monkeypatch):
    """The tile-cache miss path (children read back from HBM inside the launch that stored them) at
    the benchmark's scale: the 1 000-taxon traversal over 1 M patterns in 64 recycled slots with a
    ONE-slot cache against the level-by-level kernels, every final CLV and scaler bit for bit; and
    the plain (one slot per node) list over 262 144 patterns, all scalers, per-pattern lnL and a
    sample of CLVs."""
    _memo_tips(monkeypatch)
    attrs = PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP

    def run(w, fused, slots, sample):
        monkeypatch.setenv("PLL_GPU_FUSED", "1" if fused else "0")
        monkeypatch.setenv("PLL_GPU_FUSED_SLOTS", str(slots))
        part, pidx = S.build_partition(gpu_lib, w, attrs)
        part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
        part.update_partials(w.ops)
        ps = np.zeros(w.sites)
        lnl = part.edge_loglikelihood(w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b),
                                      w.root_matrix, pidx, persite=ps)
        import hashlib
        sc = [hashlib.sha1(part.get_scaler(k).tobytes()).hexdigest() for k in range(w.inner)]
        clv = [hashlib.sha1(part.get_clv(w.tips + k).tobytes()).hexdigest() for k in sample]
        part.destroy()
        return lnl, ps, sc, clv

    w = S.recycle_slots(S.make_workload(1000, 1_000_000, states=4), 64)
    sample = list(range(0, 64, 5))
    ref = run(w, False, 3, sample)
    got = run(w, True, 1, sample)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and got[2] == ref[2] and got[3] == ref[3]

    w = S.make_workload(1000, 262_144, states=4)
    sample = list(range(0, w.inner, 37))
    ref = run(w, False, 3, sample)
    for slots in (1, 3):
        got = run(w, True, slots, sample)
        assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and got[2] == ref[2] and got[3] == ref[3], slots


@pytest.mark.parametrize("lo", [0, 524_288 - 5_000])
def test_benchmark_shape_against_the_reference(gpu_lib, ref_lib, monkeypatch, lo):
    """BASELINE configs[1] itself: the 1 000-taxon, 27-level operations list of bench.py on two
    10 000-pattern windows of the 1 M-pattern alignment, device against the reference's AVX2 path:
    all 998 scaler arrays and CLVs bit for bit (same P-matrices), per-pattern lnL within 1e-10."""
    from libpll_b200.binding import PLL_ATTRIB_ARCH_AVX2

    w = S.make_workload(1000, 1_000_000, states=4)
    hi = lo + 10_000
    rates = ref_lib.gamma_rates(w.alpha, w.rate_cats)
    pg, pidx = S.build_partition(gpu_lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP, lo=lo, hi=hi, rates=rates)
    pr, _ = S.build_partition(ref_lib, w, PLL_ATTRIB_ARCH_AVX2 | PLL_ATTRIB_PATTERN_TIP, lo=lo, hi=hi, rates=rates)
    pg.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    pr.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    for m in range(w.prob_matrices):
        pg.set_pmatrix(m, pr.get_pmatrix(m))
    pg.update_partials(w.ops)
    pr.update_partials(w.ops)
    rescaled = 0
    for k in range(w.inner):
        a, b = pg.get_scaler(k), pr.get_scaler(k)
        np.testing.assert_array_equal(a, b, err_msg=f"scaler {k}")
        rescaled += int(b.sum())
        assert pg.get_clv(w.tips + k).tobytes() == pr.get_clv(w.tips + k).tobytes(), f"CLV {w.tips + k}"
    assert rescaled > 0, "the benchmark tree rescales"
    ps_g, ps_r = np.zeros(hi - lo), np.zeros(hi - lo)
    args = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
    lg = pg.edge_loglikelihood(*args, persite=ps_g)
    lr = pr.edge_loglikelihood(*args, persite=ps_r)
    np.testing.assert_allclose(ps_g, ps_r, rtol=1e-10, atol=0)
    assert abs(lg - lr) <= 1e-10 * abs(lr)
    pg.destroy()
    pr.destroy()
