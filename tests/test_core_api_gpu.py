"""GPU parity of the direct-call surface `pll_core_*` (reference src/pll.h:864-1000, 1659-1700;
libpll_b200/csrc/host/pll_core.c): plain host arrays in, host arrays out, no partition.

Inputs are the host arrays of a REFERENCE partition (oracle/_ref) after a full traversal - its CLVs,
scalers, tip characters, P-matrices and eigen-data - handed to the same `pll_core_*` entry point of
both libraries with the same `attrib`; outputs must agree (DNA CLVs / scalers bit for bit, 20 states
within the DMMA path's 1e-12, sums within 1e-10).  5 states with PLL_ATTRIB_ARCH_CPU exercises the
re-padding between the caller's layout (states_padded = 5) and the device's (8)."""
import ctypes as C

import numpy as np
import pytest

from libpll_b200 import synthetic as S
from libpll_b200.binding import (PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_ARCH_CPU, PLL_ATTRIB_ARCH_SSE,
                                 PLL_ATTRIB_PATTERN_TIP, PLL_ATTRIB_RATE_SCALERS)

pytestmark = pytest.mark.gpu

u32p, f64p, u8p, i32p = C.POINTER(C.c_uint), C.POINTER(C.c_double), C.POINTER(C.c_ubyte), C.POINTER(C.c_int)
f64pp = C.POINTER(f64p)
U = C.c_uint
SIGS = {
    "pll_core_create_lookup": (None, [U, U, f64p, f64p, f64p, u32p, U, U]),
    "pll_core_update_partial_tt": (None, [U, U, U, f64p, u32p, u8p, u8p, u32p, U, f64p, U]),
    "pll_core_update_partial_ti": (None, [U, U, U, f64p, u32p, u8p, f64p, f64p, f64p, u32p, u32p, U, U]),
    "pll_core_update_partial_ii": (None, [U, U, U, f64p, u32p, f64p, f64p, f64p, f64p, u32p, u32p, U]),
    "pll_core_update_pmatrix": (C.c_int, [f64pp, U, U, f64p, f64p, u32p, u32p, f64p, f64pp, f64pp, f64pp, U, U]),
    "pll_core_update_sumtable_ii": (C.c_int, [U, U, U, f64p, f64p, u32p, u32p, f64pp, f64pp, f64pp, f64p, U]),
    "pll_core_update_sumtable_ti": (C.c_int, [U, U, U, f64p, u8p, u32p, f64pp, f64pp, f64pp, u32p, U, f64p, U]),
    "pll_core_likelihood_derivatives": (C.c_int, [U, U, U, f64p, u32p, u32p, i32p, u32p, C.c_double, f64p, f64pp,
                                                  f64p, f64pp, f64p, f64p, f64p, U]),
    "pll_core_root_loglikelihood": (C.c_double, [U, U, U, f64p, u32p, f64pp, f64p, u32p, f64p, i32p, u32p, f64p, U]),
    "pll_core_edge_loglikelihood_ii": (C.c_double, [U, U, U, f64p, u32p, f64p, u32p, f64p, f64pp, f64p, u32p, f64p,
                                                    i32p, u32p, f64p, U]),
    "pll_core_edge_loglikelihood_ti": (C.c_double, [U, U, U, f64p, u32p, u8p, u32p, U, f64p, f64pp, f64p, u32p,
                                                    f64p, i32p, u32p, f64p, U]),
}


def _fn(lib, name):
    f = getattr(lib.dll, name)
    f.restype, f.argtypes = SIGS[name]
    return f


def _al(values, dtype=np.float64):
    """64-byte aligned copy (the reference's AVX kernels use aligned loads and stores)"""
    values = np.asarray(values, dtype=dtype)
    raw = np.empty(values.size * values.itemsize + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % 64
    out = raw[off:off + values.size * values.itemsize].view(dtype)
    out[:] = values.ravel()
    return out


def _d(a):
    return a.ctypes.data_as(f64p)


def _u(a):
    return a.ctypes.data_as(u32p) if a is not None else None


class Case:
    """A reference partition after a full traversal + views of its host arrays."""

    def __init__(self, ref_lib, states, rate_cats, arch, rate_scalers, tips=14, sites=203, pinv=0.0):
        self.attrib = arch | (PLL_ATTRIB_RATE_SCALERS if rate_scalers else 0)
        w = S.make_workload(tips, sites, states=states, rate_cats=rate_cats, seed=11 + states)
        # long branches on a third of the tree so that some scaler counts are not zero
        rng = np.random.default_rng(5)
        w.branch_lengths = np.where(rng.random(w.prob_matrices) < 0.3, 30.0, w.branch_lengths)
        self.w = w
        rates = ref_lib.gamma_rates(w.alpha, w.rate_cats)
        self.part, self.pidx = S.build_partition(ref_lib, w, self.attrib | PLL_ATTRIB_PATTERN_TIP, rates=rates)
        p = self.part
        if pinv:
            p.update_invariant_sites()
            p.update_invariant_sites_proportion(0, pinv)
        p.update_prob_matrices(self.pidx, w.matrix_indices, w.branch_lengths)
        p.update_partials(w.ops)
        self.K, self.R, self.sites = states, rate_cats, sites
        self.Kp = p.p.states_padded
        self.span = self.R * self.Kp
        self.kinds = {}
        for op in w.ops:
            a, b = int(op["child1_clv_index"]), int(op["child2_clv_index"])
            kind = "tt" if a < w.tips and b < w.tips else ("ii" if a >= w.tips and b >= w.tips else "ti")
            self.kinds.setdefault(kind, op)

    def clv(self, i):
        return np.ctypeslib.as_array(self.part.p.clv[i], shape=(self.sites * self.span,))

    def scaler(self, i):
        if i < 0:
            return None
        n = self.sites * (self.R if self.attrib & PLL_ATTRIB_RATE_SCALERS else 1)
        return np.ctypeslib.as_array(self.part.p.scale_buffer[i], shape=(n,))

    def pmatrix(self, i):
        return np.ctypeslib.as_array(self.part.p.pmatrix[i], shape=(self.R * self.K * self.Kp,))

    def tip(self, i):
        return self.part.p.tipchars[i]

    def tipmap(self):
        return self.part.p.tipmap, self.part.p.maxstates

    def ptrs(self, field):
        """per-rate pointer array (the way reference src/derivatives.c:196-207 gathers them)"""
        arr = (f64p * self.R)()
        for r in range(self.R):
            arr[r] = getattr(self.part.p, field)[int(self.pidx[r])]
        return arr


def _clv_close(got, want, states):
    if states == 4:
        assert np.array_equal(got, want)
    else:
        scale = np.maximum(np.abs(want), 1e-300)
        assert np.all(np.abs(got - want) <= 1e-12 * scale + 1e-300), float(np.max(np.abs(got - want) / scale))


CASES = [(4, 4, PLL_ATTRIB_ARCH_AVX2, False), (4, 4, PLL_ATTRIB_ARCH_AVX2, True), (20, 4, PLL_ATTRIB_ARCH_AVX2, False),
         (5, 3, PLL_ATTRIB_ARCH_CPU, False), (4, 1, PLL_ATTRIB_ARCH_AVX2, False), (20, 2, PLL_ATTRIB_ARCH_AVX2, True),
         (5, 2, PLL_ATTRIB_ARCH_SSE, False)]


@pytest.mark.parametrize("states,cats,arch,rate_scalers", CASES)
def test_core_partials(gpu_lib, ref_lib, states, cats, arch, rate_scalers):
    c = Case(ref_lib, states, cats, arch, rate_scalers)
    tm, tms = c.tipmap()
    nsc = c.sites * (c.R if rate_scalers else 1)
    seen_scaling = 0
    rng = np.random.default_rng(3)
    # every other pattern is pushed below the scaling threshold (2^-256) and the children arrive with
    # made-up scaler counts: the parent's counts must be their sum plus the new rescalings
    shrink = np.where(np.arange(c.sites) % 2 == 0, 1e-45, 1.0).repeat(c.span)
    for kind, op in c.kinds.items():
        a, b = int(op["child1_clv_index"]), int(op["child2_clv_index"])
        ma, mb = int(op["child1_matrix_index"]), int(op["child2_matrix_index"])
        if kind == "ti" and a >= c.w.tips:  # the tip goes left (reference src/partials.c:88-113)
            a, b, ma, mb = b, a, mb, ma
        clv_a = None if a < c.w.tips else _al(c.clv(a) * shrink)
        clv_b = None if b < c.w.tips else _al(c.clv(b) * (shrink if kind == "ii" else shrink * shrink))
        sc_a = _al(rng.integers(0, 3, nsc), np.uint32)
        sc_b = _al(rng.integers(0, 3, nsc), np.uint32)
        outs = []
        for lib in (ref_lib, gpu_lib):
            clv = _al(np.full(c.sites * c.span, np.nan))
            sc = _al(np.full(nsc, 77), np.uint32)
            if kind == "ii":
                _fn(lib, "pll_core_update_partial_ii")(c.K, c.sites, c.R, _d(clv), _u(sc), _d(clv_a), _d(clv_b),
                                                       _d(c.pmatrix(ma)), _d(c.pmatrix(mb)), _u(sc_a), _u(sc_b),
                                                       c.attrib)
            elif kind == "ti":
                _fn(lib, "pll_core_update_partial_ti")(c.K, c.sites, c.R, _d(clv), _u(sc), c.tip(a), _d(clv_b),
                                                       _d(c.pmatrix(ma)), _d(c.pmatrix(mb)), _u(sc_b), tm, tms,
                                                       c.attrib)
            else:
                # the reference's table: maxstates^2 (rounded up to a power of 4) x span doubles
                lookup = _al(np.zeros(max(1024 * c.R, (1 << (2 * int(np.ceil(np.log2(max(tms, 2)))))) * c.span)))
                _fn(lib, "pll_core_create_lookup")(c.K, c.R, _d(lookup), _d(c.pmatrix(ma)), _d(c.pmatrix(mb)), tm, tms,
                                                   c.attrib)
                _fn(lib, "pll_core_update_partial_tt")(c.K, c.sites, c.R, _d(clv), _u(sc), c.tip(a), c.tip(b), tm, tms,
                                                       _d(lookup), c.attrib)
            assert lib is ref_lib or lib.errno() == 0, lib.errmsg()
            outs.append((clv, sc))
        (rc, rs), (gc, gs) = outs
        assert np.array_equal(rs, gs), kind
        # pads of the caller's layout are not part of the contract
        view = lambda x: x.reshape(-1, c.Kp)[:, :c.K]
        _clv_close(view(gc), view(rc), states)
        if kind == "tt":
            assert not rs.any()      # tip-tip zeroes the parent's counts (src/core_partials_avx.c:113-116)
        else:
            seen_scaling += int((rs > (sc_b if kind == "ti" else sc_a + sc_b)).sum())
    assert set(c.kinds) == {"tt", "ti", "ii"}
    assert seen_scaling > 0
    c.part.destroy()


@pytest.mark.parametrize("states,cats,arch,rate_scalers", CASES)
def test_core_pmatrix(gpu_lib, ref_lib, states, cats, arch, rate_scalers):
    c = Case(ref_lib, states, cats, arch, rate_scalers, tips=6, sites=8)
    p = c.part.p
    count = 5
    lengths = _al([0.0, 1e-6, 0.05, 1.3, 40.0])
    which = np.array([3, 0, 4, 1, 2], dtype=np.uint32)  # scattered destinations
    params = np.ascontiguousarray(c.pidx, dtype=np.uint32)
    rates = _al(np.ctypeslib.as_array(p.rates, shape=(c.R,)))
    pinv = _al(np.ctypeslib.as_array(p.prop_invar, shape=(p.rate_matrices,)))
    outs = []
    for lib in (ref_lib, gpu_lib):
        mats = [_al(np.full(c.R * c.K * c.Kp, np.nan)) for _ in range(count)]
        arr = (f64p * count)(*[_d(m) for m in mats])
        ok = _fn(lib, "pll_core_update_pmatrix")(arr, c.K, c.R, _d(rates), _d(lengths), _u(which), _u(params),
                                                 _d(pinv), p.eigenvals, p.eigenvecs, p.inv_eigenvecs, count, c.attrib)
        assert ok, lib.errmsg()
        outs.append(mats)
    for r, g in zip(*outs):
        r, g = r.reshape(-1, c.Kp)[:, :c.K], g.reshape(-1, c.Kp)[:, :c.K]
        assert np.allclose(g, r, rtol=1e-12, atol=1e-15)
    c.part.destroy()


@pytest.mark.parametrize("pinv", [0.0, 0.25])
@pytest.mark.parametrize("states,cats,arch,rate_scalers", CASES)
def test_core_likelihood_sumtable_derivatives(gpu_lib, ref_lib, states, cats, arch, rate_scalers, pinv):
    if pinv and arch != PLL_ATTRIB_ARCH_AVX2:
        pytest.skip("one p-inv case per alphabet size is enough")
    if arch == PLL_ATTRIB_ARCH_SSE:
        pytest.skip("the reference's SSE log-likelihood of an odd alphabet is not reproducible (NaN on a repeated call)")
    c = Case(ref_lib, states, cats, arch, rate_scalers, pinv=pinv)
    p, w = c.part.p, c.w
    tm, tms = c.tipmap()
    freqs, evecs, ievecs, evals = c.ptrs("frequencies"), c.ptrs("eigenvecs"), c.ptrs("inv_eigenvecs"), c.ptrs("eigenvals")
    ident = np.arange(c.R, dtype=np.uint32)
    weights = (np.arange(c.sites, dtype=np.uint32) % 5) + 1
    rw = _al(np.ctypeslib.as_array(p.rate_weights, shape=(c.R,)))
    rates = _al(np.ctypeslib.as_array(p.rates, shape=(c.R,)))
    pinv_r = _al(np.full(c.R, pinv))
    inv = p.invariant if pinv else None
    top = w.tips + w.inner - 1
    ii = c.kinds["ii"]
    a, b = int(ii["child1_clv_index"]), int(ii["child2_clv_index"])
    sa, sb, mb = int(ii["child1_scaler_index"]), int(ii["child2_scaler_index"]), int(ii["child2_matrix_index"])
    ti = c.kinds["ti"]
    t, n = int(ti["child1_clv_index"]), int(ti["child2_clv_index"])
    mt, sn = int(ti["child1_matrix_index"]), int(ti["child2_scaler_index"])
    if t >= w.tips:
        t, n, mt, sn = n, t, int(ti["child2_matrix_index"]), int(ti["child1_scaler_index"])
    res = []
    for lib in (ref_lib, gpu_lib):
        out = {}
        site = _al(np.full(c.sites, np.nan))
        out["root"] = _fn(lib, "pll_core_root_loglikelihood")(c.K, c.sites, c.R, _d(c.clv(top)), _u(c.scaler(w.scaler_of(top))),
                                                              freqs, _d(rw), _u(weights), _d(pinv_r), inv, _u(ident),
                                                              _d(site), c.attrib)
        out["root_site"] = site.copy()
        out["edge_ii"] = _fn(lib, "pll_core_edge_loglikelihood_ii")(
            c.K, c.sites, c.R, _d(c.clv(a)), _u(c.scaler(sa)), _d(c.clv(b)), _u(c.scaler(sb)), _d(c.pmatrix(mb)), freqs,
            _d(rw), _u(weights), _d(pinv_r), inv, _u(ident), None, c.attrib)
        out["edge_ti"] = _fn(lib, "pll_core_edge_loglikelihood_ti")(
            c.K, c.sites, c.R, _d(c.clv(n)), _u(c.scaler(sn)), c.tip(t), tm, tms, _d(c.pmatrix(mt)), freqs, _d(rw),
            _u(weights), _d(pinv_r), inv, _u(ident), None, c.attrib)
        st_ii = _al(np.full(c.sites * c.span, np.nan))
        assert _fn(lib, "pll_core_update_sumtable_ii")(c.K, c.sites, c.R, _d(c.clv(a)), _d(c.clv(b)), _u(c.scaler(sa)),
                                                       _u(c.scaler(sb)), evecs, ievecs, freqs, _d(st_ii), c.attrib), lib.errmsg()
        st_ti = _al(np.full(c.sites * c.span, np.nan))
        assert _fn(lib, "pll_core_update_sumtable_ti")(c.K, c.sites, c.R, _d(c.clv(n)), c.tip(t), _u(c.scaler(sn)), evecs,
                                                       ievecs, freqs, tm, tms, _d(st_ti), c.attrib), lib.errmsg()
        out["st_ii"], out["st_ti"] = st_ii, st_ti
        res.append(out)
    r, g = res
    for key in ("root", "edge_ii", "edge_ti"):
        assert np.isfinite(r[key]) and abs(g[key] - r[key]) <= 1e-10 * abs(r[key]), (key, g[key], r[key])
    assert np.allclose(g["root_site"], r["root_site"], rtol=1e-10, atol=0)
    view = lambda x: x.reshape(-1, c.Kp)[:, :c.K]
    for key in ("st_ii", "st_ti"):
        scale = np.abs(view(r[key])).max(axis=1, keepdims=True) + 1e-300
        assert np.all(np.abs(view(g[key]) - view(r[key])) <= 1e-11 * scale), key
    # derivatives from the REFERENCE's table, by both libraries (the table is an input here)
    for key in ("st_ii", "st_ti"):
        for t_len in (0.01, 0.4, 2.0):
            d = []
            for lib in (ref_lib, gpu_lib):
                d1, d2 = C.c_double(), C.c_double()
                assert _fn(lib, "pll_core_likelihood_derivatives")(
                    c.K, c.sites, c.R, _d(rw), None, None, inv, _u(weights), t_len, _d(pinv_r), freqs, _d(rates), evals,
                    _d(r[key]), C.byref(d1), C.byref(d2), c.attrib), lib.errmsg()
                d.append((d1.value, d2.value))
            (r1, r2), (g1, g2) = d
            scale = max(abs(r1), float(weights.sum()) * 1e-3)
            assert abs(g1 - r1) <= 1e-10 * scale and abs(g2 - r2) <= 1e-10 * max(abs(r2), scale), (key, t_len, d)
    gpu_lib.dll.pll_gpu_core_release()
    c.part.destroy()
