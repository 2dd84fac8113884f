"""Python driver of the CPU oracle (TEST INFRASTRUCTURE ONLY).

oracle/pll_oracle.c restates the reference's hot-path kernels in scalar C; this module loads
it with ctypes and assembles whole evaluations from it with an independent model set-up:
  * eigendecomposition by numpy.linalg.eigh of sqrt(pi) Q sqrt(pi)^-1 (reference
    src/models.c:182-331 uses Householder + QL: same matrix, same normalisation)
  * discrete-Gamma mean rates from scipy's chi-square quantile / regularised incomplete gamma
    (the closed form Yang 1994 eq. 10 that reference src/gamma.c:262-283 evaluates)
so that it cross-checks the product's host code as well as its kernels (to ~1e-13, not bit
for bit; bit-level checks call the orc_core_* functions directly with shared inputs).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libpll_oracle.so")

c_double_p = C.POINTER(C.c_double)
c_uint_p = C.POINTER(C.c_uint)
c_ubyte_p = C.POINTER(C.c_ubyte)
c_int_p = C.POINTER(C.c_int)
c_dpp = C.POINTER(c_double_p)

ATTR_RATE_SCALERS = 1 << 9

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "pll_oracle.c")):
            subprocess.run(["make", "-C", _HERE, "port"], check=True, stdout=subprocess.DEVNULL)
        _lib = C.CDLL(_LIB)
        _lib.orc_core_edge_loglikelihood_ii.restype = C.c_double
        _lib.orc_core_edge_loglikelihood_ti.restype = C.c_double
        _lib.orc_core_root_loglikelihood.restype = C.c_double
    return _lib


# ---- ctypes helpers ----------------------------------------------------------------------
def dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def up(a):
    return None if a is None else a.ctypes.data_as(c_uint_p)


def bp(a):
    return None if a is None else a.ctypes.data_as(c_ubyte_p)


def ip(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


def dpp(arrays):
    """double** over a list of contiguous float64 arrays (kept alive by the caller)."""
    arr = (c_double_p * len(arrays))()
    for i, a in enumerate(arrays):
        arr[i] = a.ctypes.data_as(c_double_p)
    return arr


# ---- independent model set-up ---------------------------------------------------------------
def gamma_mean_rates(alpha: float, cats: int) -> np.ndarray:
    from scipy.special import gammainc
    from scipy.stats import gamma as gamma_dist

    if cats == 1:
        return np.ones(1)
    cuts = gamma_dist.ppf(np.arange(1, cats) / cats, a=alpha, scale=1.0 / alpha)
    g = gammainc(alpha + 1.0, cuts * alpha)
    g = np.concatenate([[0.0], g, [1.0]])
    return np.diff(g) * cats


def eigen(subst: np.ndarray, freqs: np.ndarray):
    """(eigenvecs V [K,K], inv_eigenvecs V^-1 [K,K], eigenvals [K]) with
    Q = V^-1 diag(l) V in the reference's row convention (reference src/models.c:293-320)."""
    K = len(freqs)
    q = np.zeros((K, K))
    p = np.asarray(subst, dtype=np.float64).copy()
    if p[-1] > 0:
        p = p / p[-1]
    k = 0
    for i in range(K):
        for j in range(i + 1, K):
            q[i, j] = q[j, i] = p[k] * np.sqrt(freqs[i] * freqs[j])
            q[i, i] -= p[k] * freqs[j]
            q[j, j] -= p[k] * freqs[i]
            k += 1
    mean = float(np.sum(freqs * -np.diag(q)))
    q /= mean
    vals, vecs = np.linalg.eigh(q)  # columns are eigenvectors of the symmetric matrix
    a = vecs.T  # row k = eigenvector k
    sq = np.sqrt(freqs)
    evecs = a * sq[None, :]
    ievecs = (a.T) / sq[:, None]
    return np.ascontiguousarray(evecs), np.ascontiguousarray(ievecs), np.ascontiguousarray(vals)


# ---- a tiny partition-like driver over the C kernels -------------------------------------------
class PortPartition:
    """Holds CLVs / scalers / P-matrices as numpy arrays and runs pll_operation_t lists
    through the orc_core_* kernels in array order (reference src/partials.c:177-213)."""

    def __init__(self, tips, clv_buffers, states, sites, rate_cats, prob_matrices, scale_buffers,
                 pattern_tip=True, rate_scalers=False):
        self.lib = load()
        self.tips, self.K, self.S, self.R = tips, states, sites, rate_cats
        self.pattern_tip = pattern_tip
        self.attrib = ATTR_RATE_SCALERS if rate_scalers else 0
        self.clv = {}
        self.n_clv = tips + clv_buffers
        self.tipchars = {}
        self.tipmap = np.zeros(256, dtype=np.uint32)
        self.maxstates = 0
        self.charmap = np.zeros(256, dtype=np.uint8)
        slen = sites * rate_cats if rate_scalers else sites
        self.scalers = [np.zeros(slen, dtype=np.uint32) for _ in range(scale_buffers)]
        self.pmat = [np.zeros(rate_cats * states * states) for _ in range(prob_matrices)]
        self.weights = np.ones(sites, dtype=np.uint32)
        self.invariant = None
        self.prop_invar = None
        self.rates = np.ones(rate_cats)
        self.rate_weights = np.full(rate_cats, 1.0 / rate_cats)
        self.models = {}  # index -> (evecs, ievecs, evals, freqs)

    # -- set-up -----------------------------------------------------------------------------
    def set_model(self, idx, subst, freqs):
        ev, iev, val = eigen(np.asarray(subst, float), np.asarray(freqs, float))
        self.models[idx] = (ev, iev, val, np.ascontiguousarray(freqs, dtype=np.float64))
        if self.prop_invar is None or len(self.prop_invar) <= idx:
            old = self.prop_invar
            self.prop_invar = np.zeros(max(idx + 1, len(self.models)))
            if old is not None:
                self.prop_invar[:len(old)] = old

    def set_tip_states(self, tip, seq: bytes, amap):
        amap = np.asarray(amap, dtype=np.uint32)
        chars = np.frombuffer(seq, dtype=np.uint8)
        masks = amap[chars]
        assert np.all(masks != 0), "illegal state"
        if self.pattern_tip:
            if self.K == 4:
                self.tipchars[tip] = masks.astype(np.uint8)
                self.maxstates = 16
            else:
                # one code per distinct mask in ASCII order (reference src/pll.c:272-397)
                for ch in range(256):
                    m = amap[ch]
                    if m == 0:
                        continue
                    known = np.nonzero(self.tipmap[:self.maxstates] == m)[0]
                    if len(known):
                        self.charmap[ch] = known[0]
                    else:
                        self.tipmap[self.maxstates] = m
                        self.charmap[ch] = self.maxstates
                        self.maxstates += 1
                self.tipchars[tip] = self.charmap[chars].copy()
        else:
            bits = ((masks[:, None] >> np.arange(self.K)[None, :]) & 1).astype(np.float64)
            self.clv[tip] = np.ascontiguousarray(np.repeat(bits[:, None, :], self.R, axis=1)).reshape(-1)

    def update_invariant(self):
        state = np.full(self.S, (1 << self.K) - 1, dtype=np.uint32)
        for t in range(self.tips):
            if self.pattern_tip:
                c = self.tipchars[t]
                state &= (c.astype(np.uint32) if self.K == 4 else self.tipmap[c])
            else:
                v = self.clv[t].reshape(self.S, self.R, self.K)[:, 0, :]
                state &= (v.astype(np.uint32) << np.arange(self.K, dtype=np.uint32)[None, :]).sum(axis=1).astype(np.uint32)
        pop = np.array([bin(int(x)).count("1") for x in state])
        inv = np.where(pop == 1, np.log2(np.maximum(state, 1)).astype(np.int32), -1).astype(np.int32)
        self.invariant = np.ascontiguousarray(inv)

    # -- kernels ----------------------------------------------------------------------------
    def _gather(self, pidx):
        ev = [self.models[int(i)][0] for i in pidx]
        iev = [self.models[int(i)][1] for i in pidx]
        val = [self.models[int(i)][2] for i in pidx]
        fr = [self.models[int(i)][3] for i in pidx]
        return ev, iev, val, fr

    def update_prob_matrices(self, pidx, matrix_indices, branch_lengths):
        n_rm = max(self.models) + 1
        ev = [self.models[i][0] for i in range(n_rm)]
        iev = [self.models[i][1] for i in range(n_rm)]
        val = [self.models[i][2] for i in range(n_rm)]
        pm = dpp(self.pmat)
        mi = np.ascontiguousarray(matrix_indices, dtype=np.uint32)
        bl = np.ascontiguousarray(branch_lengths, dtype=np.float64)
        pi = np.ascontiguousarray(pidx, dtype=np.uint32)
        pinv = np.ascontiguousarray(self.prop_invar, dtype=np.float64)
        rc = self.lib.orc_core_update_pmatrix(pm, self.K, self.R, dp(self.rates), dp(bl), up(mi), up(pi),
                                              dp(pinv), dpp(val), dpp(ev), dpp(iev), len(mi), self.attrib)
        assert rc == 1

    def _clv(self, idx):
        if idx not in self.clv:
            self.clv[idx] = np.zeros(self.S * self.R * self.K)
        return self.clv[idx]

    def update_partials(self, ops):
        for op in ops:
            p = int(op["parent_clv_index"])
            ps = int(op["parent_scaler_index"])
            c1, c2 = int(op["child1_clv_index"]), int(op["child2_clv_index"])
            m1, m2 = int(op["child1_matrix_index"]), int(op["child2_matrix_index"])
            s1, s2 = int(op["child1_scaler_index"]), int(op["child2_scaler_index"])
            pscale = self.scalers[ps] if ps >= 0 else None
            t1 = self.pattern_tip and c1 < self.tips
            t2 = self.pattern_tip and c2 < self.tips
            parent = self._clv(p)
            if t1 and t2:
                self.lib.orc_core_update_partial_tt(self.K, self.S, self.R, dp(parent), up(pscale),
                                                    bp(self.tipchars[c1]), bp(self.tipchars[c2]),
                                                    dp(self.pmat[m1]), dp(self.pmat[m2]), up(self.tipmap),
                                                    self.maxstates, self.attrib)
            elif t1 or t2:
                tip, inner = (c1, c2) if t1 else (c2, c1)
                tm, im = (m1, m2) if t1 else (m2, m1)
                isc = s2 if t1 else s1
                self.lib.orc_core_update_partial_ti(self.K, self.S, self.R, dp(parent), up(pscale),
                                                    bp(self.tipchars[tip]), dp(self._clv(inner)),
                                                    dp(self.pmat[tm]), dp(self.pmat[im]),
                                                    up(self.scalers[isc] if isc >= 0 else None),
                                                    up(self.tipmap), self.maxstates, self.attrib)
            else:
                self.lib.orc_core_update_partial_ii(self.K, self.S, self.R, dp(parent), up(pscale),
                                                    dp(self._clv(c1)), dp(self._clv(c2)),
                                                    dp(self.pmat[m1]), dp(self.pmat[m2]),
                                                    up(self.scalers[s1] if s1 >= 0 else None),
                                                    up(self.scalers[s2] if s2 >= 0 else None), self.attrib)

    def edge_loglikelihood(self, pc, ps, cc, cs, matrix, fidx, persite=None):
        fidx = np.ascontiguousarray(fidx, dtype=np.uint32)
        n_rm = max(self.models) + 1
        freqs = [self.models[i][3] for i in range(n_rm)]
        pinv = np.ascontiguousarray(self.prop_invar, dtype=np.float64)
        ptip = self.pattern_tip and pc < self.tips
        ctip = self.pattern_tip and cc < self.tips
        if ptip or ctip:
            inner, tip, isc = (cc, pc, cs) if ptip else (pc, cc, ps)
            return self.lib.orc_core_edge_loglikelihood_ti(
                self.K, self.S, self.R, dp(self._clv(inner)), up(self.scalers[isc] if isc >= 0 else None),
                bp(self.tipchars[tip]), up(self.tipmap), self.maxstates, dp(self.pmat[matrix]), dpp(freqs),
                dp(self.rate_weights), up(self.weights), dp(pinv), ip(self.invariant), up(fidx), dp(persite),
                self.attrib)
        return self.lib.orc_core_edge_loglikelihood_ii(
            self.K, self.S, self.R, dp(self._clv(pc)), up(self.scalers[ps] if ps >= 0 else None),
            dp(self._clv(cc)), up(self.scalers[cs] if cs >= 0 else None), dp(self.pmat[matrix]), dpp(freqs),
            dp(self.rate_weights), up(self.weights), dp(pinv), ip(self.invariant), up(fidx), dp(persite),
            self.attrib)

    def root_loglikelihood(self, clv, scaler, fidx, persite=None):
        fidx = np.ascontiguousarray(fidx, dtype=np.uint32)
        n_rm = max(self.models) + 1
        freqs = [self.models[i][3] for i in range(n_rm)]
        pinv = np.ascontiguousarray(self.prop_invar, dtype=np.float64)
        return self.lib.orc_core_root_loglikelihood(
            self.K, self.S, self.R, dp(self._clv(clv)), up(self.scalers[scaler] if scaler >= 0 else None),
            dpp(freqs), dp(self.rate_weights), up(self.weights), dp(pinv), ip(self.invariant), up(fidx),
            dp(persite), self.attrib)

    def sumtable(self, pc, cc, ps, cs, pidx):
        ev, iev, val, fr = self._gather(pidx)
        out = np.zeros(self.S * self.R * self.K)
        ptip = self.pattern_tip and pc < self.tips
        ctip = self.pattern_tip and cc < self.tips
        if ptip or ctip:
            inner, tip, isc = (cc, pc, cs) if ptip else (pc, cc, ps)
            self.lib.orc_core_update_sumtable_ti(self.K, self.S, self.R, dp(self._clv(inner)),
                                                 bp(self.tipchars[tip]),
                                                 up(self.scalers[isc] if isc >= 0 else None), dpp(ev), dpp(iev),
                                                 dpp(fr), up(self.tipmap), self.maxstates, dp(out), self.attrib)
        else:
            self.lib.orc_core_update_sumtable_ii(self.K, self.S, self.R, dp(self._clv(pc)), dp(self._clv(cc)),
                                                 up(self.scalers[ps] if ps >= 0 else None),
                                                 up(self.scalers[cs] if cs >= 0 else None), dpp(ev), dpp(iev),
                                                 dpp(fr), dp(out), self.attrib)
        return out

    def derivatives(self, branch_length, pidx, sumtable):
        ev, iev, val, fr = self._gather(pidx)
        pinv = np.ascontiguousarray([self.prop_invar[int(i)] for i in pidx], dtype=np.float64)
        d1, d2 = C.c_double(0), C.c_double(0)
        self.lib.orc_core_likelihood_derivatives(self.K, self.S, self.R, dp(self.rate_weights), None, None,
                                                 ip(self.invariant), up(self.weights), C.c_double(branch_length),
                                                 dp(pinv), dpp(fr), dp(self.rates), dpp(val), dp(sumtable),
                                                 C.byref(d1), C.byref(d2), self.attrib)
        return d1.value, d2.value


# ---- whole synthetic workloads --------------------------------------------------------------
def build_port_partition(w, lo=0, hi=None, variant="default", rates=None, pattern_tip=True,
                         rate_scalers=False, maps=None):
    """PortPartition over sites [lo, hi) of a libpll_b200.synthetic.Workload."""
    from libpll_b200 import synthetic as S

    hi = w.sites if hi is None else hi
    part = PortPartition(w.tips, w.inner, w.states, hi - lo, w.rate_cats, w.prob_matrices, w.inner,
                         pattern_tip=pattern_tip, rate_scalers=rate_scalers)
    if w.states == 4:
        part.set_model(0, S.GTR_RATES, S.GTR_FREQS)
        pidx = np.zeros(w.rate_cats, np.uint32)
    else:
        raise NotImplementedError("port workloads with amino-acid models need the model tables: "
                                  "pass them through PortPartition.set_model")
    part.rates = np.ascontiguousarray(gamma_mean_rates(w.alpha, w.rate_cats) if rates is None else rates)
    amap = maps if maps is not None else dna_map()
    for t in range(w.tips):
        part.set_tip_states(t, S.tip_sequence(w, t, lo, hi), amap)
    part.weights = np.ascontiguousarray(w.weights[lo:hi], dtype=np.uint32)
    return part, pidx


def dna_map() -> np.ndarray:
    """IUPAC nucleotide codes -> 4-bit masks (same table as pll_map_nt, reference src/maps.c:46)."""
    m = np.zeros(256, dtype=np.uint32)
    codes = {"A": 1, "C": 2, "G": 4, "T": 8, "U": 8, "M": 3, "R": 5, "W": 9, "S": 6, "Y": 10, "K": 12,
             "V": 7, "H": 11, "D": 13, "B": 14, "N": 15, "O": 15, "X": 15, "-": 15, "?": 15}
    for ch, v in codes.items():
        m[ord(ch)] = v
        m[ord(ch.lower())] = v
    return m


def full_evaluation(w, lo=0, hi=None, rates=None) -> float:
    part, pidx = build_port_partition(w, lo, hi, rates=rates)
    part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    part.update_partials(w.ops)
    return part.edge_loglikelihood(w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b),
                                   w.root_matrix, pidx)


def cpu_traversal_rate(w, threads, sites_per_thread, reps):
    """bench.py fallback when oracle/_ref is absent: the scalar port, one thread."""
    part, pidx = build_port_partition(w, 0, sites_per_thread)
    best = None
    lnl = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
        part.update_partials(w.ops)
        lnl = part.edge_loglikelihood(w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b),
                                      w.root_matrix, pidx)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return len(w.ops) * sites_per_thread / best, best, lnl


# ---- adapter: the oracle port behind the same Python surface as libpll_b200.binding ---------
class _PortPartitionAPI:
    """Subset of libpll_b200.binding.Partition implemented on the oracle port, so that
    tests/golden_runner.py can drive the port with the same step lists."""

    def __init__(self, tips, clv_buffers, states, sites, rate_matrices, prob_matrices, rate_cats,
                 scale_buffers, attributes):
        self.pp = PortPartition(tips, clv_buffers, states, sites, rate_cats, prob_matrices, scale_buffers,
                                pattern_tip=bool(attributes & (1 << 4)),
                                rate_scalers=bool(attributes & ATTR_RATE_SCALERS))
        self._freqs, self._subst = {}, {}
        self.pp.prop_invar = np.zeros(rate_matrices)
        self._tables = {}
        self._map = dna_map() if states == 4 else aa_map()

    def _remodel(self, idx):
        if idx in self._freqs and idx in self._subst:
            pinv = self.pp.prop_invar.copy()
            self.pp.set_model(idx, self._subst[idx], self._freqs[idx])
            self.pp.prop_invar = pinv

    def set_frequencies(self, idx, f):
        self._freqs[idx] = np.asarray(f, float)
        self._remodel(idx)

    def set_subst_params(self, idx, s):
        self._subst[idx] = np.asarray(s, float)
        self._remodel(idx)

    def set_category_rates(self, r):
        self.pp.rates = np.ascontiguousarray(r, dtype=np.float64)

    def set_category_weights(self, w):
        self.pp.rate_weights = np.ascontiguousarray(w, dtype=np.float64)

    def set_pattern_weights(self, w):
        self.pp.weights = np.ascontiguousarray(w, dtype=np.uint32)

    def set_tip_states(self, tip, seq, amap=None):
        self.pp.set_tip_states(tip, seq, self._map if amap is None else amap)

    def update_invariant_sites_proportion(self, idx, pinv):
        if self.pp.invariant is None:
            self.pp.update_invariant()
        self.pp.prop_invar[idx] = pinv

    def update_prob_matrices(self, pidx, mi, bl):
        self.pp.update_prob_matrices(pidx, mi, bl)

    def update_partials(self, ops):
        self.pp.update_partials(ops)

    def get_pmatrix(self, idx):
        return self.pp.pmat[idx].reshape(self.pp.R, self.pp.K, self.pp.K).copy()

    def get_clv(self, idx):
        return self.pp._clv(idx).reshape(self.pp.S, self.pp.R, self.pp.K).copy()

    def get_scaler(self, idx):
        return self.pp.scalers[idx].copy()

    def edge_loglikelihood(self, pc, ps, cc, cs, m, fidx, persite=None):
        return self.pp.edge_loglikelihood(pc, ps, cc, cs, m, fidx, persite)

    def root_loglikelihood(self, clv, sc, fidx, persite=None):
        return self.pp.root_loglikelihood(clv, sc, fidx, persite)

    def new_sumtable(self):
        return np.zeros(self.pp.S * self.pp.R * self.pp.K)

    def update_sumtable(self, pc, cc, ps, cs, pidx, table):
        table[:] = self.pp.sumtable(pc, cc, ps, cs, pidx)

    def likelihood_derivatives(self, ps, cs, t, pidx, table):
        return self.pp.derivatives(t, pidx, table)

    def destroy(self):
        pass


class PortAsLibrary:
    is_gpu = False
    use_case_rates = True  # see tests/golden_runner.py

    def gamma_rates(self, alpha, cats, mode=0):
        assert mode == 0
        return gamma_mean_rates(alpha, cats)

    @staticmethod
    def make_map(table):
        m = np.zeros(256, dtype=np.uint32)
        for ch, mask in table.items():
            m[ord(ch)] = mask
        return m

    def partition(self, **kw):
        return _PortPartitionAPI(**kw)


def aa_map() -> np.ndarray:
    """Amino-acid codes -> 20-bit masks (same table as pll_map_aa, reference src/maps.c:66)."""
    m = np.zeros(256, dtype=np.uint32)
    for i, ch in enumerate("ARNDCQEGHILKMFPSTWYV"):
        m[ord(ch)] = m[ord(ch.lower())] = 1 << i
    m[ord("B")] = m[ord("b")] = (1 << 2) | (1 << 3)
    m[ord("Z")] = m[ord("z")] = (1 << 5) | (1 << 6)
    for ch in "Xx*-?":
        m[ord(ch)] = 0xFFFFF
    return m
