"""Deterministic synthetic workloads: tree, branch lengths, tips,
pattern weights and models.  Pure numpy; shared by the GPU runs, the CPU baseline and the
parity tests so that all of them see exactly the same inputs.

Tree: random-join unrooted binary topology.  Start with the T tip ids, repeatedly join two
uniformly chosen live nodes into inner node T+k (CLV index T+k, scale buffer k; a child's
P-matrix index is the child's CLV index) until two nodes remain; the likelihood is evaluated
on the edge between those two.  The operations come out in creation order, which is a valid
post-order.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .binding import (
    OP_DTYPE,
    PLL_ATTRIB_PATTERN_TIP,
    PLL_SCALE_BUFFER_NONE,
    Partition,
    PllLibrary,
)

DNA_ALPHABET = np.frombuffer(b"ACGT", dtype=np.uint8)
AA_ALPHABET = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)

GENERIC_LETTERS = b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdef"  # alphabets other than DNA / amino acids


def generic_map(states: int) -> dict:
    """char -> state-set mask for a synthetic alphabet of `states` (<= 32) letters: '-' is the
    full ambiguity, '1' = first two states, '2' = last two states."""
    assert 2 <= states <= 32
    m = {chr(GENERIC_LETTERS[i]): 1 << i for i in range(states)}
    m["-"] = (1 << states) - 1
    m["1"] = 0b11
    m["2"] = 0b11 << (states - 2)
    return m


def _alphabet(states: int):
    """(letters, full-ambiguity char, two two-state ambiguity chars)"""
    if states == 4:
        return DNA_ALPHABET, ord("N"), (ord("R"), ord("Y"))
    if states == 20:
        return AA_ALPHABET, ord("X"), (ord("B"), ord("Z"))
    return np.frombuffer(GENERIC_LETTERS[:states], dtype=np.uint8), ord("-"), (ord("1"), ord("2"))


GTR_RATES = np.array([1.2, 3.1, 0.9, 1.1, 3.3, 1.0])
GTR_FREQS = np.array([0.30, 0.20, 0.25, 0.25])


@dataclass
class Workload:
    tips: int
    sites: int
    states: int
    rate_cats: int = 4
    alpha: float = 0.5
    seed: int = 42
    ops: np.ndarray = field(default=None, repr=False)
    matrix_indices: np.ndarray = field(default=None, repr=False)
    branch_lengths: np.ndarray = field(default=None, repr=False)
    root_a: int = 0  # the two nodes of the evaluation edge
    root_b: int = 0
    root_matrix: int = 0
    root_seq: np.ndarray = field(default=None, repr=False)  # [sites] state indices
    weights: np.ndarray = field(default=None, repr=False)
    n_slots: int = 0  # > 0: inner CLVs / scalers live in this many recycled slots

    # ---- derived sizes -------------------------------------------------------------
    @property
    def inner(self) -> int:
        """CLV buffers / scale buffers to allocate."""
        return self.n_slots if self.n_slots else self.tips - 2

    @property
    def prob_matrices(self) -> int:
        return 2 * self.tips - 2

    def scaler_of(self, node: int) -> int:
        return node - self.tips if node >= self.tips else PLL_SCALE_BUFFER_NONE

    def op_kinds(self):
        """(#tip-tip, #tip-inner, #inner-inner) operations with pattern tips."""
        t1 = self.ops["child1_clv_index"] < self.tips
        t2 = self.ops["child2_clv_index"] < self.tips
        tt = int(np.sum(t1 & t2))
        ti = int(np.sum(t1 ^ t2))
        return tt, ti, len(self.ops) - tt - ti

    def algorithmic_bytes_per_site(self) -> int:
        """ii = 3*span+12, ti = 2*span+1+8, tt = span+2+4 (per-site scalers)."""
        span = self.rate_cats * self.states * 8
        tt, ti, ii = self.op_kinds()
        return ii * (3 * span + 12) + ti * (2 * span + 9) + tt * (span + 6)


def make_workload(tips: int, sites: int, states: int = 4, rate_cats: int = 4, alpha: float = 0.5,
                  seed: int = 42) -> Workload:
    assert tips >= 3
    w = Workload(tips=tips, sites=sites, states=states, rate_cats=rate_cats, alpha=alpha, seed=seed)
    rng = np.random.default_rng(seed)
    live = list(range(tips))
    ops = np.zeros(tips - 2, dtype=OP_DTYPE)
    for k in range(tips - 2):
        i = int(rng.integers(0, len(live)))
        a = live.pop(i)
        j = int(rng.integers(0, len(live)))
        b = live.pop(j)
        parent = tips + k
        ops[k] = (parent, k, a, a, w.scaler_of(a), b, b, w.scaler_of(b))
        live.append(parent)
    w.ops = ops
    # an inner node plays "parent" of the evaluation edge (a tip-tip edge cannot occur: T >= 3)
    a, b = live
    if a < tips:
        a, b = b, a
    w.root_a, w.root_b = a, b
    w.root_matrix = b
    rng_b = np.random.default_rng(seed + 2)
    w.matrix_indices = np.arange(w.prob_matrices, dtype=np.uint32)
    w.branch_lengths = rng_b.uniform(0.01, 0.21, size=w.prob_matrices)
    rng_r = np.random.default_rng(seed + 1)
    w.root_seq = rng_r.integers(0, states, size=sites, dtype=np.uint8)
    rng_w = np.random.default_rng(seed + 3)
    w.weights = rng_w.integers(1, 5, size=sites, dtype=np.uint32)
    return w


TIP_BLOCK = 1 << 16  # sites per independently seeded block (makes tips site-sliceable)


def recycle_slots(w: Workload, max_slots: int) -> Workload:
    """Re-indexes the inner CLVs / scale buffers of a workload onto a pool of `max_slots`
    recycled slots (SURVEY.md section 7 "Capacity": 5,000 x 10M cannot keep every CLV).

    The tree is walked depth-first from the evaluation edge, larger subtree first
    (Sethi-Ullman order), a node's slot is taken when its operation is emitted and its
    children's slots return to the pool right after.  P-matrix indices stay node indices.
    The result is an ordinary pll_operation_t list - slot reuse is legal pll.h API use - whose
    later operations overwrite slots earlier ones read (WAR/WAW hazards the batched scheduler
    must honour).  Raises if `max_slots` is too small for the tree."""
    import sys

    T = w.tips
    children = {int(o["parent_clv_index"]): (int(o["child1_clv_index"]), int(o["child2_clv_index"])) for o in w.ops}
    size = {}

    def subtree(n):
        if n < T:
            size[n] = 1
        else:
            a, b = children[n]
            size[n] = 1 + subtree(a) + subtree(b)
        return size[n]

    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 4 * T + 100))
    try:
        for r in (w.root_a, w.root_b):
            subtree(r)
        free = list(range(max_slots - 1, -1, -1))
        slot = {}
        new_ops = []

        def emit(n):
            if n < T:
                return
            a, b = children[n]
            first, second = (a, b) if size[a] >= size[b] else (b, a)
            emit(first)
            emit(second)
            if not free:
                raise ValueError(f"max_slots={max_slots} is too small for this tree")
            slot[n] = free.pop()

            def idx(c):
                return c if c < T else T + slot[c]

            def sc(c):
                return PLL_SCALE_BUFFER_NONE if c < T else slot[c]

            new_ops.append((T + slot[n], slot[n], idx(a), a, sc(a), idx(b), b, sc(b)))
            for c in (a, b):
                if c >= T:
                    free.append(slot[c])

        emit(w.root_a)
        emit(w.root_b)
    finally:
        sys.setrecursionlimit(old)

    out = Workload(tips=w.tips, sites=w.sites, states=w.states, rate_cats=w.rate_cats, alpha=w.alpha, seed=w.seed)
    out.ops = np.array(new_ops, dtype=OP_DTYPE)
    out.matrix_indices, out.branch_lengths = w.matrix_indices, w.branch_lengths
    out.root_seq, out.weights = w.root_seq, w.weights
    out.root_a = w.root_a if w.root_a < T else T + slot[w.root_a]
    out.root_b = w.root_b if w.root_b < T else T + slot[w.root_b]
    out.root_matrix = w.root_matrix
    out.n_slots = max_slots
    return out


def tip_sequence(w: Workload, tip: int, lo: int = 0, hi: Optional[int] = None) -> bytes:
    """Characters of one tip for sites [lo, hi): root state w.p. 0.7 else uniform; 1 % full
    ambiguity (N / X), 0.5 % two-state ambiguity (R,Y / B,Z).  Generated in independently
    seeded blocks of TIP_BLOCK sites, so the value at a site does not depend on the slice
    asked for and a rank only generates the slice it owns."""
    hi = w.sites if hi is None else hi
    alphabet, amb_full, amb2 = _alphabet(w.states)
    amb_full = np.uint8(amb_full)
    amb2 = (np.uint8(amb2[0]), np.uint8(amb2[1]))
    out = np.empty(hi - lo, dtype=np.uint8)
    for blk in range(lo // TIP_BLOCK, (hi + TIP_BLOCK - 1) // TIP_BLOCK):
        b0, b1 = blk * TIP_BLOCK, min((blk + 1) * TIP_BLOCK, w.sites)
        rng = np.random.default_rng([w.seed + 1, tip, blk])
        n = b1 - b0
        u = rng.random(n, dtype=np.float32)
        alt = rng.integers(0, w.states, size=n, dtype=np.uint8)
        u2 = rng.random(n, dtype=np.float32)
        chars = alphabet[np.where(u < 0.7, w.root_seq[b0:b1], alt)]
        chars[u2 < 0.01] = amb_full
        two = (u2 >= 0.01) & (u2 < 0.015)
        chars[two & ((alt & 1) == 0)] = amb2[0]
        chars[two & ((alt & 1) == 1)] = amb2[1]
        s0, s1 = max(lo, b0), min(hi, b1)
        out[s0 - lo:s1 - lo] = chars[s0 - b0:s1 - b0]
    return out.tobytes()


def model_for(lib: PllLibrary, w: Workload, variant: str = "default"):
    """(rate_matrices, [(subst_params, freqs)], params_indices, rates, rate_weights)."""
    rates = lib.gamma_rates(w.alpha, w.rate_cats)
    weights = np.full(w.rate_cats, 1.0 / w.rate_cats)
    if w.states == 4:
        return 1, [(GTR_RATES, GTR_FREQS)], np.zeros(w.rate_cats, np.uint32), rates, weights
    if w.states != 20:
        # any other alphabet: a seeded reversible model (exchangeabilities in [0.5, 3], last one 1)
        rng = np.random.default_rng([w.seed + 2, w.states])
        r = rng.uniform(0.5, 3.0, w.states * (w.states - 1) // 2)
        r[-1] = 1.0
        f = rng.uniform(0.5, 1.5, w.states)
        return 1, [(r, f / f.sum())], np.zeros(w.rate_cats, np.uint32), rates, weights
    if variant == "lg4m":
        assert w.rate_cats == 4
        r = lib.aa_table("pll_aa_rates_lg4m", (4, 190))
        f = lib.aa_table("pll_aa_freqs_lg4m", (4, 20))
        return 4, [(r[i], f[i]) for i in range(4)], np.arange(4, dtype=np.uint32), rates, weights
    r = lib.aa_table("pll_aa_rates_lg", (190,))
    f = lib.aa_table("pll_aa_freqs_lg", (20,))
    return 1, [(r, f)], np.zeros(w.rate_cats, np.uint32), rates, weights


def build_partition(lib: PllLibrary, w: Workload, attributes: int, lo: int = 0,
                    hi: Optional[int] = None, variant: str = "default",
                    rates: Optional[np.ndarray] = None) -> tuple[Partition, np.ndarray]:
    """Creates a partition over sites [lo, hi) of the workload, sets model, tips and weights.
    Returns (partition, params_indices).  `rates` lets a caller impose category rates computed
    elsewhere (parity tests give both libraries the same doubles)."""
    hi = w.sites if hi is None else hi
    n_rm, models, pidx, g_rates, g_weights = model_for(lib, w, variant)
    pattern_tip = bool(attributes & PLL_ATTRIB_PATTERN_TIP)
    part = lib.partition(tips=w.tips, clv_buffers=w.inner, states=w.states, sites=hi - lo,
                         rate_matrices=n_rm, prob_matrices=w.prob_matrices, rate_cats=w.rate_cats,
                         scale_buffers=w.inner, attributes=attributes)
    for i, (sp, fr) in enumerate(models):
        part.set_frequencies(i, fr)
        part.set_subst_params(i, sp)
    part.set_category_rates(g_rates if rates is None else rates)
    part.set_category_weights(g_weights)
    amap = None if w.states in (4, 20) else lib.make_map(generic_map(w.states))
    for t in range(w.tips):
        part.set_tip_states(t, tip_sequence(w, t, lo, hi), amap)
    part.set_pattern_weights(w.weights[lo:hi])
    del pattern_tip
    return part, pidx


def full_evaluation(part: Partition, w: Workload, pidx: np.ndarray) -> float:
    """One full-tree log-likelihood evaluation: all P-matrices, the whole post-order
    traversal, the edge log-likelihood (SURVEY.md 8(d) metric ii)."""
    part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    part.update_partials(w.ops)
    return part.edge_loglikelihood(w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b),
                                   w.root_matrix, pidx)


# ---- counter-based tips (generated on the device for the 10 M-pattern config) ----
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser over uint64 (wrap-around arithmetic), the function
    libpll_b200/csrc/gpu/plg_synth.cu evaluates per character."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def hash_tip_sequence(seed: int, tip: int, lo: int, hi: int) -> bytes:
    """Host restatement of plg_generate_tipchars (DNA): ASCII characters of tip `tip` for alignment
    columns [lo, hi) - root state w.p. 0.7 else uniform, 1 % N, 0.5 % R / Y."""
    site = np.arange(lo, hi, dtype=np.uint64)
    with np.errstate(over="ignore"):
        root = _splitmix64(np.uint64(seed) ^ ((site * np.uint64(0xD1342543DE82EF95)) & _M64)) & np.uint64(3)
        tip_key = _splitmix64(np.array([(seed + 0x632BE59BD9B4E019 * (tip + 1)) & 0xFFFFFFFFFFFFFFFF],
                                       dtype=np.uint64))[0]
        h = _splitmix64(tip_key ^ site)
    u = h & np.uint64(0xFFFF)
    alt = (h >> np.uint64(16)) & np.uint64(3)
    u2 = (h >> np.uint64(32)) & np.uint64(0xFFFF)
    state = np.where(u < 45875, root, alt).astype(np.intp)
    chars = DNA_ALPHABET[state].copy()
    two = (u2 >= 655) & (u2 < 983)
    chars[two & ((alt & np.uint64(1)) == 0)] = ord("R")
    chars[two & ((alt & np.uint64(1)) == 1)] = ord("Y")
    chars[u2 < 655] = ord("N")
    return chars.tobytes()
