"""GPU: the recipe of reference test/src/partial-traversal.c (its testdata/246x4465 files are not
shipped, so a seeded random tree and alignment stand in): after a full traversal the virtual
root is moved to random inner nodes; each move traverses only the subtrees whose CLVs are not yet
oriented towards the new root (pll_utree_traverse with a pruning callback), updates only the
listed P-matrices and CLVs, and evaluates the edge log-likelihood.  The device library and the
reference execute the same partial operation lists; every log-likelihood must agree at 1e-10
and stay equal to the first one (moving the root of an unrooted tree does not change it)."""
import ctypes as C

import numpy as np
import pytest

from libpll_b200 import trees as T
from libpll_b200.binding import (OP_DTYPE, PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP,
                                 PLL_ATTRIB_RATE_SCALERS)
from test_utree_cpu import random_newick

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _partition(lib, arch, tree, seqs, sites, extra=0):
    part = lib.partition(tips=tree.tips, clv_buffers=tree.inner, states=4, sites=sites, rate_matrices=1,
                         prob_matrices=2 * tree.tips - 3, rate_cats=4, scale_buffers=tree.inner,
                         attributes=arch | PLL_ATTRIB_PATTERN_TIP | extra)
    part.set_frequencies(0, [0.3, 0.2, 0.25, 0.25])
    part.set_subst_params(0, [1.2, 3.1, 0.9, 1.1, 3.3, 1.0])
    part.set_category_rates(lib.gamma_rates(0.7, 4))
    for i, label in enumerate(tree.tip_labels()):
        part.set_tip_states(i, seqs[label].encode())
    return part


@pytest.mark.parametrize("tips,sites,rate_scalers,slices", [(90, 1200, False, 1), (600, 200, False, 1), (600, 200, True, 1),
                                                            (600, 300, False, 3), (600, 300, True, 3)])
def test_partial_traversals_follow_the_reference(gpu_lib, ref_lib, tips, sites, rate_scalers, slices):
    """600 tips: deep enough that every re-rooting also re-accumulates scaler counts (per site or per
    rate) along the re-oriented path; 3 slices: one partition over three device contexts"""
    lib = T.bind(gpu_lib)
    tree = T.Tree(lib, newick=random_newick(tips, 17))
    rng = np.random.default_rng(4)
    seqs = {f"t{i}": "".join(rng.choice(list("ACGT-R"), sites, p=[.28, .22, .24, .22, .03, .01])) for i in range(tips)}
    extra = PLL_ATTRIB_RATE_SCALERS if rate_scalers else 0
    assert gpu_lib.pll_gpu_set_devices(slices) == 1
    try:
        pg = _partition(gpu_lib, PLL_ATTRIB_ARCH_GPU, tree, seqs, sites, extra)
    finally:
        gpu_lib.pll_gpu_set_devices(0)
    assert gpu_lib.pll_gpu_partition_devices(pg.ptr) == slices
    pr = _partition(ref_lib, PLL_ATTRIB_ARCH_AVX2, tree, seqs, sites, extra)
    pidx = np.zeros(4, np.uint32)

    oriented = {}  # record address -> is the CLV of its node oriented along this record?

    @T.TRAV_CB
    def partial(node):
        n = node.contents
        if not n.next:
            return 1
        me = C.addressof(n)
        if oriented.get(me):
            return 0
        oriented[me] = True
        oriented[C.addressof(n.next.contents)] = False
        oriented[C.addressof(n.next.contents.next.contents)] = False
        return 1

    def evaluate(root):
        buf, n = tree.traverse(lib, root, cb=partial)
        branches = np.zeros(2 * tips - 3)
        matrices = np.zeros(2 * tips - 3, dtype=np.uint32)
        ops = np.zeros(tree.inner, dtype=OP_DTYPE)
        nm, no = C.c_uint(0), C.c_uint(0)
        lib.pll_utree_create_operations(buf, n, branches.ctypes.data_as(C.POINTER(C.c_double)),
                                        matrices.ctypes.data_as(C.POINTER(C.c_uint)), ops.ctypes.data,
                                        C.byref(nm), C.byref(no))
        r = root.contents
        out = []
        for p in (pg, pr):
            p.update_prob_matrices(pidx, matrices[:nm.value], branches[:nm.value])
            p.update_partials(ops[:no.value])
            out.append(p.edge_loglikelihood(r.clv_index, r.scaler_index, r.back.contents.clv_index,
                                            r.back.contents.scaler_index, r.pmatrix_index, pidx))
        return out[0], out[1], no.value

    g0, r0, n_ops = evaluate(tree.root)
    assert n_ops == tree.inner
    assert abs(g0 - r0) <= RTOL * abs(r0)
    sizes = []
    for _ in range(40):
        node = tree.node(tips + int(rng.integers(0, tree.inner)))
        for _ in range(int(rng.integers(0, 3))):  # any of the three ring records
            node = node.contents.next
        g, r, n_ops = evaluate(node)
        sizes.append(n_ops)
        assert abs(g - r) <= RTOL * abs(r), (g, r)
        assert abs(g - g0) <= 1e-9 * abs(g0), "moving the virtual root must not change the likelihood"
    assert 0 < np.mean(sizes) < tree.inner / 2, "the traversals should really be partial"
    if tips >= 600:
        assert sum(int(np.asarray(pr.get_scaler(k)).sum()) for k in range(tree.inner)) > 0
    pg.destroy()
    pr.destroy()
    tree.destroy()
