"""GPU: the inner loop of a tree search, on both libraries - Newton optimisation of every branch
length (recipe of reference examples/newton/newton.c:60-100 applied edge by edge): for each
inner-node record the virtual root is moved there with a pruned traversal, the sumtable of the edge
is built and the branch is optimised with the derivatives; two sweeps over the tree.  Every call
on the path (partial CLV updates, P-matrices, sumtable, derivatives, edge log-likelihood) runs on
the device library and on the reference (oracle/_ref); they must walk the same sequence of branch
lengths and end at the same log-likelihood, which must not decrease from sweep to sweep."""
import ctypes as C

import numpy as np
import pytest

from libpll_b200 import trees as T
from libpll_b200.binding import OP_DTYPE, PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP
from test_utree_cpu import random_newick

pytestmark = pytest.mark.gpu


def _partition(lib, arch, tree, seqs, sites):
    part = lib.partition(tips=tree.tips, clv_buffers=tree.inner, states=4, sites=sites, rate_matrices=1,
                         prob_matrices=2 * tree.tips - 3, rate_cats=4, scale_buffers=tree.inner,
                         attributes=arch | PLL_ATTRIB_PATTERN_TIP)
    part.set_frequencies(0, [0.28, 0.22, 0.26, 0.24])
    part.set_subst_params(0, [1.0, 2.9, 0.8, 1.2, 3.4, 1.0])
    part.set_category_rates(lib.gamma_rates(0.8, 4))
    for i, label in enumerate(tree.tip_labels()):
        part.set_tip_states(i, seqs[label].encode())
    return part


def test_newton_sweeps_follow_the_reference(gpu_lib, ref_lib):
    lib = T.bind(gpu_lib)
    tips, sites = 24, 900
    rng = np.random.default_rng(12)
    # sequences evolved crudely along nothing in particular: enough signal for finite optima
    base = rng.choice(list("ACGT"), sites)
    seqs = {}
    for i in range(tips):
        s = base.copy()
        mut = rng.random(sites) < 0.08 + 0.02 * (i % 5)
        s[mut] = rng.choice(list("ACGT"), int(mut.sum()))
        seqs[f"t{i}"] = "".join(s)
    pidx = np.zeros(4, np.uint32)
    final = {}
    histories = {}
    for name, plib, arch in (("gpu", gpu_lib, PLL_ATTRIB_ARCH_GPU), ("ref", ref_lib, PLL_ATTRIB_ARCH_AVX2)):
        tree = T.Tree(lib, newick=random_newick(tips, 5))
        part = _partition(plib, arch, tree, seqs, sites)
        oriented = {}

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

        def move_root(root):
            buf, n = tree.traverse(lib, root, cb=partial)
            branches = np.zeros(2 * tips - 3)
            matrices = np.zeros(2 * tips - 3, dtype=np.uint32)
            ops = np.zeros(tree.inner, dtype=OP_DTYPE)
            nm, no = C.c_uint(0), C.c_uint(0)
            lib.pll_utree_create_operations(buf, n, branches.ctypes.data_as(C.POINTER(C.c_double)),
                                            matrices.ctypes.data_as(C.POINTER(C.c_uint)), ops.ctypes.data,
                                            C.byref(nm), C.byref(no))
            part.update_prob_matrices(pidx, matrices[:nm.value], branches[:nm.value])
            part.update_partials(ops[:no.value])

        def lnl(root):
            r = root.contents
            return part.edge_loglikelihood(r.clv_index, r.scaler_index, r.back.contents.clv_index,
                                           r.back.contents.scaler_index, r.pmatrix_index, pidx)

        history, sweep_lnl = [], []
        table = part.new_sumtable()
        for sweep in range(2):
            for i in range(tree.inner):
                node = tree.node(tips + i)
                for _ in range(3):  # the three edges of this inner node
                    move_root(node)
                    r = node.contents
                    b = r.back.contents
                    part.update_sumtable(r.clv_index, b.clv_index, r.scaler_index, b.scaler_index, pidx, table)
                    length = max(r.length, 1e-4)
                    for _it in range(12):
                        d1, d2 = part.likelihood_derivatives(r.scaler_index, b.scaler_index, length, pidx, table)
                        if abs(d1) < 1e-6 or d2 <= 0:
                            break
                        length = min(max(length - d1 / d2, 1e-6), 10.0)
                    node.contents.length = length
                    node.contents.back.contents.length = length
                    part.update_prob_matrices(pidx, [r.pmatrix_index], [length])
                    history.append(length)
                    node = node.contents.next
            move_root(tree.root)
            sweep_lnl.append(lnl(tree.root))
        final[name] = sweep_lnl
        histories[name] = np.array(history)
        part.destroy()
        tree.destroy()

    assert final["ref"][1] >= final["ref"][0] - 1e-6, "a Newton sweep must not make the tree worse"
    np.testing.assert_allclose(histories["gpu"], histories["ref"], rtol=1e-6, atol=1e-9)
    for a, b in zip(final["gpu"], final["ref"]):
        assert abs(a - b) <= 1e-9 * abs(b), (final["gpu"], final["ref"])
