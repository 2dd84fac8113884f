"""CPU tests that PIN THE ORACLE (no GPU needed):

  1. every orc_core_* kernel of oracle/pll_oracle.c BIT FOR BIT against the reference's own
     exported pll_core_* kernels called with PLL_ATTRIB_ARCH_AVX2 (oracle/_ref), on random
     inputs including forced-underflow CLVs, for 4 and 20 states, per-site and per-rate scaling;
  2. the assembled port pipeline against the committed golden cases (tests/golden/cases.json,
     generated from the reference and asserted there against its test/out text fixtures).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from libpll_b200.binding import PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_PATTERN_TIP, PLL_ATTRIB_RATE_SCALERS
from oracle import port
from oracle.port import bp, dp, dpp, ip, up

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AVX2 = PLL_ATTRIB_ARCH_AVX2


def az(n, dtype=np.float64):
    """zeros, 32-byte aligned: the reference's AVX kernels use aligned loads."""
    itemsize = np.dtype(dtype).itemsize
    raw = np.zeros(n * itemsize + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % 32
    return raw[off:off + n * itemsize].view(dtype)


def al(a):
    """32-byte aligned contiguous copy."""
    a = np.ascontiguousarray(a)
    out = az(a.size, a.dtype)
    out[:] = a.reshape(-1)
    return out


def _rand_model(rng, K):
    nsub = K * (K - 1) // 2
    subst = rng.uniform(0.2, 3.0, nsub)
    subst[-1] = 1.0
    freqs = rng.uniform(0.5, 1.5, K)
    freqs /= freqs.sum()
    ev, iev, val = port.eigen(subst, freqs)
    return al(ev), al(iev), al(val), al(freqs)


def _rand_pmats(rng, K, R, n, lib):
    ev, iev, val, fr = _rand_model(rng, K)
    rates = np.ascontiguousarray(port.gamma_mean_rates(0.7, R))
    pm = [az(R * K * K) for _ in range(n)]
    bl = np.ascontiguousarray(rng.uniform(0.01, 1.5, n))
    mi = np.arange(n, dtype=np.uint32)
    pi = np.zeros(R, dtype=np.uint32)
    pinv = np.zeros(1)
    rc = lib.orc_core_update_pmatrix(dpp(pm), K, R, dp(rates), dp(bl), up(mi), up(pi), dp(pinv),
                                     dpp([val]), dpp([ev]), dpp([iev]), n, 0)
    assert rc == 1
    return pm, (ev, iev, val, fr, rates)


def _rand_clv(rng, S, R, K, underflow_frac=0.3):
    clv = rng.uniform(0.0, 1.0, (S, R, K))
    tiny = rng.random(S) < underflow_frac
    clv[tiny] *= 2.0 ** -200  # products of two such children fall below 2^-256
    some = rng.random((S, R)) < 0.2
    clv[some] *= 2.0 ** -140
    return al(clv.reshape(-1))


def _tips(rng, S, K):
    if K == 4:
        return np.ascontiguousarray(rng.integers(1, 16, S, dtype=np.uint8)), np.zeros(256, np.uint32), 16
    tipmap = np.zeros(256, dtype=np.uint32)
    codes = [1 << i for i in range(20)] + [(1 << 2) | (1 << 3), (1 << 5) | (1 << 6), 0xFFFFF]
    tipmap[:len(codes)] = codes
    return np.ascontiguousarray(rng.integers(0, len(codes), S, dtype=np.uint8)), tipmap, len(codes)


@pytest.mark.parametrize("K", [4, 20])
def test_pmatrix_bit_exact(ref_lib, port_lib, K):
    rng = np.random.default_rng(K)
    R, n = 4, 9
    ev, iev, val, fr = _rand_model(rng, K)
    rates = np.ascontiguousarray(port.gamma_mean_rates(0.3, R))
    bl = np.ascontiguousarray(np.concatenate([[0.0, 1e-9, 1e-6], rng.uniform(0.01, 3.0, n - 4), [100.0]]))
    mi = np.arange(n, dtype=np.uint32)
    pi = np.zeros(R, dtype=np.uint32)
    for pinv in (0.0, 0.35):
        a = [az(R * K * K) for _ in range(n)]
        b = [az(R * K * K) for _ in range(n)]
        pv = np.array([pinv])
        port_lib.orc_core_update_pmatrix(dpp(a), K, R, dp(rates), dp(bl), up(mi), up(pi), dp(pv), dpp([val]),
                                         dpp([ev]), dpp([iev]), n, 0)
        ref_lib.dll.pll_core_update_pmatrix(dpp(b), K, R, dp(rates), dp(bl), up(mi), up(pi), dp(pv),
                                            dpp([val]), dpp([ev]), dpp([iev]), n, AVX2)
        for x, y in zip(a, b):
            assert x.tobytes() == y.tobytes()


@pytest.mark.parametrize("K", [4, 20])
@pytest.mark.parametrize("rate_scalers", [False, True])
def test_partials_bit_exact(ref_lib, port_lib, K, rate_scalers):
    rng = np.random.default_rng(100 + K + rate_scalers)
    S, R = 257, 4
    attrib = PLL_ATTRIB_RATE_SCALERS if rate_scalers else 0
    pm, _ = _rand_pmats(rng, K, R, 2, port_lib)
    left, right = _rand_clv(rng, S, R, K), _rand_clv(rng, S, R, K)
    slen = S * R if rate_scalers else S
    ls = np.ascontiguousarray(rng.integers(0, 5, slen, dtype=np.uint32))
    rs = np.ascontiguousarray(rng.integers(0, 5, slen, dtype=np.uint32))
    tl, tipmap, maxstates = _tips(rng, S, K)
    tr, _, _ = _tips(rng, S, K)

    # inner-inner, with every combination of present/absent scalers
    for use_p, use_l, use_r in [(1, 1, 1), (1, 1, 0), (1, 0, 0), (0, 1, 1)]:
        pa, pb = az(S * R * K), az(S * R * K)
        sa, sb = np.full(slen, 77, np.uint32), np.full(slen, 77, np.uint32)
        port_lib.orc_core_update_partial_ii(K, S, R, dp(pa), up(sa) if use_p else None, dp(left), dp(right),
                                            dp(pm[0]), dp(pm[1]), up(ls) if use_l else None,
                                            up(rs) if use_r else None, attrib)
        ref_lib.dll.pll_core_update_partial_ii(K, S, R, dp(pb), up(sb) if use_p else None, dp(left), dp(right),
                                               dp(pm[0]), dp(pm[1]), up(ls) if use_l else None,
                                               up(rs) if use_r else None, attrib | AVX2)
        assert pa.tobytes() == pb.tobytes()
        assert np.array_equal(sa, sb)
        if use_p and use_l and use_r:
            assert (sa > ls + rs).any(), "the inputs must actually trigger rescaling"
        if not use_p:
            assert (sa == 77).all()

    # tip-inner
    pa, pb = az(S * R * K), az(S * R * K)
    sa, sb = np.zeros(slen, np.uint32), np.zeros(slen, np.uint32)
    port_lib.orc_core_update_partial_ti(K, S, R, dp(pa), up(sa), bp(tl), dp(right), dp(pm[0]), dp(pm[1]), up(rs),
                                        up(tipmap), maxstates, attrib)
    ref_lib.dll.pll_core_update_partial_ti(K, S, R, dp(pb), up(sb), bp(tl), dp(right), dp(pm[0]), dp(pm[1]),
                                           up(rs), up(tipmap), maxstates, attrib | AVX2)
    assert pa.tobytes() == pb.tobytes()
    assert np.array_equal(sa, sb)

    # tip-tip (the reference goes through its pre-multiplied lookup table)
    pa, pb = az(S * R * K), az(S * R * K)
    sa, sb = np.full(slen, 5, np.uint32), np.full(slen, 5, np.uint32)
    port_lib.orc_core_update_partial_tt(K, S, R, dp(pa), up(sa), bp(tl), bp(tr), dp(pm[0]), dp(pm[1]), up(tipmap),
                                        maxstates, attrib)
    log2max = int(np.ceil(np.log2(maxstates)))
    lookup = az((1 << (2 * log2max)) * R * K + 4096)
    ref_lib.dll.pll_core_create_lookup(K, R, dp(lookup), dp(pm[0]), dp(pm[1]), up(tipmap), maxstates, AVX2)
    ref_lib.dll.pll_core_update_partial_tt(K, S, R, dp(pb), up(sb), bp(tl), bp(tr), up(tipmap), maxstates,
                                           dp(lookup), attrib | AVX2)
    assert pa.tobytes() == pb.tobytes()
    assert np.array_equal(sa, sb) and not sa.any()


@pytest.mark.parametrize("K", [4, 20])
@pytest.mark.parametrize("rate_scalers", [False, True])
@pytest.mark.parametrize("pinv", [0.0, 0.4])
def test_likelihood_bit_exact(ref_lib, port_lib, K, rate_scalers, pinv):
    for fn in ("pll_core_edge_loglikelihood_ii", "pll_core_edge_loglikelihood_ti",
               "pll_core_edge_loglikelihood_ti_4x4", "pll_core_root_loglikelihood"):
        getattr(ref_lib.dll, fn).restype = C.c_double
    rng = np.random.default_rng(7 + K + 2 * rate_scalers)
    S, R = 203, 4
    attrib = PLL_ATTRIB_RATE_SCALERS if rate_scalers else 0
    pm, (ev, iev, val, fr, rates) = _rand_pmats(rng, K, R, 1, port_lib)
    p, c = _rand_clv(rng, S, R, K, 0.0), _rand_clv(rng, S, R, K, 0.0)
    slen = S * R if rate_scalers else S
    ps = np.ascontiguousarray(rng.integers(0, 6, slen, dtype=np.uint32))
    cs = np.ascontiguousarray(rng.integers(0, 6, slen, dtype=np.uint32))
    w = np.ascontiguousarray(rng.integers(1, 5, S, dtype=np.uint32))
    rw = np.ascontiguousarray(rng.dirichlet(np.ones(R)))
    inv = np.ascontiguousarray(rng.integers(-1, K, S).astype(np.int32))
    pv = np.array([pinv])
    fi = np.zeros(R, dtype=np.uint32)
    tips, tipmap, maxstates = _tips(rng, S, K)
    for name, args_port, args_ref in [
        ("edge_ii",
         lambda out: port_lib.orc_core_edge_loglikelihood_ii(K, S, R, dp(p), up(ps), dp(c), up(cs), dp(pm[0]),
                                                             dpp([fr]), dp(rw), up(w), dp(pv), ip(inv), up(fi),
                                                             dp(out), attrib),
         lambda out: ref_lib.dll.pll_core_edge_loglikelihood_ii(K, S, R, dp(p), up(ps), dp(c), up(cs), dp(pm[0]),
                                                                dpp([fr]), dp(rw), up(w), dp(pv), ip(inv), up(fi),
                                                                dp(out), attrib | AVX2)),
        ("edge_ti",
         lambda out: port_lib.orc_core_edge_loglikelihood_ti(K, S, R, dp(p), up(ps), bp(tips), up(tipmap),
                                                             maxstates, dp(pm[0]), dpp([fr]), dp(rw), up(w), dp(pv),
                                                             ip(inv), up(fi), dp(out), attrib),
         (lambda out: ref_lib.dll.pll_core_edge_loglikelihood_ti_4x4(S, R, dp(p), up(ps), bp(tips), dp(pm[0]),
                                                                     dpp([fr]), dp(rw), up(w), dp(pv), ip(inv),
                                                                     up(fi), dp(out), attrib | AVX2)) if K == 4 else
         (lambda out: ref_lib.dll.pll_core_edge_loglikelihood_ti(K, S, R, dp(p), up(ps), bp(tips), up(tipmap),
                                                                 maxstates, dp(pm[0]), dpp([fr]), dp(rw), up(w),
                                                                 dp(pv), ip(inv), up(fi), dp(out), attrib | AVX2))),
        ("root",
         lambda out: port_lib.orc_core_root_loglikelihood(K, S, R, dp(p), up(ps[:S].copy()), dpp([fr]), dp(rw), up(w),
                                                          dp(pv), ip(inv), up(fi), dp(out), attrib),
         lambda out: ref_lib.dll.pll_core_root_loglikelihood(K, S, R, dp(p), up(ps[:S].copy()), dpp([fr]), dp(rw),
                                                             up(w), dp(pv), ip(inv), up(fi), dp(out), attrib | AVX2)),
    ]:
        oa, ob = az(S), az(S)
        la, lb = args_port(oa), args_ref(ob)
        assert oa.tobytes() == ob.tobytes(), name
        assert la == lb, name


@pytest.mark.parametrize("K", [4, 20])
@pytest.mark.parametrize("rate_scalers", [False, True])
def test_sumtable_and_derivatives_bit_exact(ref_lib, port_lib, K, rate_scalers):
    rng = np.random.default_rng(31 + K)
    S, R = 131, 4  # not a multiple of 4: exercises the backwards tail loop
    attrib = PLL_ATTRIB_RATE_SCALERS if rate_scalers else 0
    ev, iev, val, fr = _rand_model(rng, K)
    p, c = _rand_clv(rng, S, R, K, 0.0), _rand_clv(rng, S, R, K, 0.0)
    slen = S * R if rate_scalers else S
    ps = np.ascontiguousarray(rng.integers(0, 6, slen, dtype=np.uint32))
    cs = np.ascontiguousarray(rng.integers(0, 6, slen, dtype=np.uint32))
    tips, tipmap, maxstates = _tips(rng, S, K)
    evs, ievs, frs, vals = [ev] * R, [iev] * R, [fr] * R, [val] * R

    ta, tb = az(S * R * K), az(S * R * K)
    port_lib.orc_core_update_sumtable_ii(K, S, R, dp(p), dp(c), up(ps), up(cs), dpp(evs), dpp(ievs), dpp(frs), dp(ta), attrib)
    ref_lib.dll.pll_core_update_sumtable_ii(K, S, R, dp(p), dp(c), up(ps), up(cs), dpp(evs), dpp(ievs), dpp(frs), dp(tb),
                                            attrib | AVX2)
    assert ta.tobytes() == tb.tobytes()

    ua, ub = az(S * R * K), az(S * R * K)
    port_lib.orc_core_update_sumtable_ti(K, S, R, dp(p), bp(tips), up(ps), dpp(evs), dpp(ievs), dpp(frs), up(tipmap),
                                         maxstates, dp(ua), attrib)
    ref_lib.dll.pll_core_update_sumtable_ti(K, S, R, dp(p), bp(tips), up(ps), dpp(evs), dpp(ievs), dpp(frs), up(tipmap),
                                            maxstates, dp(ub), attrib | AVX2)
    assert ua.tobytes() == ub.tobytes()

    rates = np.ascontiguousarray(port.gamma_mean_rates(0.9, R))
    inv = np.ascontiguousarray(rng.integers(-1, K, S).astype(np.int32))
    for weights, rw, pinv in [
        (np.ones(S, np.uint32), np.full(R, 0.25), 0.0),
        (rng.integers(1, 6, S).astype(np.uint32), np.full(R, 0.25), 0.0),
        (rng.integers(1, 6, S).astype(np.uint32), rng.dirichlet(np.ones(R)), 0.3),
    ]:
        weights, rw = np.ascontiguousarray(weights), np.ascontiguousarray(rw)
        pv = np.full(R, pinv)
        for t in (0.01, 0.4, 3.0):
            a1, a2, b1, b2 = C.c_double(), C.c_double(), C.c_double(), C.c_double()
            port_lib.orc_core_likelihood_derivatives(K, S, R, dp(rw), None, None, ip(inv), up(weights), C.c_double(t),
                                                     dp(pv), dpp(frs), dp(rates), dpp(vals), dp(ta), C.byref(a1),
                                                     C.byref(a2), attrib)
            ref_lib.dll.pll_core_likelihood_derivatives(K, S, R, dp(rw), None, None, ip(inv), up(weights),
                                                        C.c_double(t), dp(pv), dpp(frs), dp(rates), dpp(vals), dp(tb),
                                                        C.byref(b1), C.byref(b2), attrib | AVX2)
            assert (a1.value, a2.value) == (b1.value, b2.value), (t, pinv)


# ---------------------------------------------------------------------------------------------
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "cases.json")))


def compare_outputs(got, want, rtol, what):
    assert len(got) == len(want), what
    for g, w in zip(got, want):
        assert g["kind"] == w["kind"], what
        if g["kind"] in ("pmatrix", "clv"):
            np.testing.assert_allclose(g["values"], w["values"], rtol=rtol, atol=1e-15, err_msg=what)  # atol: entries are probabilities <= 1
        elif g["kind"] in ("edge", "root"):
            np.testing.assert_allclose(g["persite"], w["persite"], rtol=rtol, atol=1e-12, err_msg=what)
            assert abs(g["logl"] - w["logl"]) <= rtol * abs(w["logl"]), (what, g["logl"], w["logl"])
        elif g["kind"] == "derivs":
            scale = max(abs(w["d_f"]), 1.0)
            assert abs(g["d_f"] - w["d_f"]) <= rtol * scale, (what, g, w)
            assert abs(g["dd_f"] - w["dd_f"]) <= rtol * max(abs(w["dd_f"]), 1.0), (what, g, w)
        elif g["kind"] == "newton":
            assert g["iterations"] == w["iterations"], what
            assert abs(g["final"] - w["final"]) <= 1e-9 * abs(w["final"]), what


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
@pytest.mark.parametrize("variant", ["tv", "notv"])
def test_port_reproduces_golden_cases(case, variant):
    """oracle port (independent eigen / gamma set-up) vs the reference's recorded outputs."""
    from golden_runner import execute

    if case.get("no_port"):
        pytest.skip("ascertainment-bias cases are pinned on the reference itself (oracle/_ref), not on the port")
    attr = PLL_ATTRIB_ARCH_AVX2 | (PLL_ATTRIB_PATTERN_TIP if variant == "tv" else 0)
    got = execute(port.PortAsLibrary(), case, attr)
    compare_outputs(got, case["expect"][variant], 1e-10, f"{case['name']}[{variant}]")


def test_golden_cases_match_reference_text_fixtures():
    """The fixtures carry the numbers printed in the reference's test/out files."""
    by_name = {c["name"]: c for c in GOLDEN}
    e = [o for o in by_name["test_00010_NMDU_lkcalc"]["expect"]["tv"] if o["kind"] == "edge"]
    assert all(abs(o["logl"] - (-58.887310)) < 5e-7 for o in e)
    e = [o for o in by_name["test_00011_NMAU_lkcalc"]["expect"]["notv"] if o["kind"] == "edge"]
    assert all(abs(o["logl"] - (-227.371279)) < 5e-7 for o in e)
    e = [o["logl"] for o in by_name["example_unrooted"]["expect"]["notv"] if o["kind"] == "edge"]
    np.testing.assert_allclose(e, [-33.387713, -34.550204, -36.830297], atol=5e-7)
    n = [o for o in by_name["example_newton"]["expect"]["tv"] if o["kind"] == "newton"][0]
    assert n["iterations"] == 7 and abs(n["final"] - 2.607098) < 5e-7
