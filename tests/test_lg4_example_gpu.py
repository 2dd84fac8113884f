"""GPU: the reference's data-driven example (examples/lg4/lg4.c on examples/lg4/data, 21 taxa x
113 amino-acid sites, LG4M then LG4X) run end to end on this library - Newick reader, FASTA
reader, traversal -> operations, P-matrices, CLV updates and edge log-likelihoods on the device -
against the values recorded from the reference (tests/golden/lg4_example.json, generator
tests/golden/make_lg4_golden.py).  Also the recycled-slot operation list on the device."""
import json
import os

import numpy as np
import pytest

from libpll_b200 import trees as T
from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LG4 = json.load(open(os.path.join(ROOT, "tests", "golden", "lg4_example.json")))
RTOL = 1e-10


def _load(gpu_lib, tmp_path):
    lib = T.bind(gpu_lib)
    tree_file, fasta_file = tmp_path / "example.tree", tmp_path / "example.fas"
    tree_file.write_text(LG4["newick"])
    fasta_file.write_text(LG4["fasta_text"])
    tree = T.Tree(lib, path=str(tree_file))
    records, err = T.read_fasta(lib, str(fasta_file))
    assert err == 102 and len(records) == tree.tips
    tip_of = {label: i for i, label in enumerate(tree.tip_labels())}
    return lib, tree, records, tip_of


def _partition(lib, tree, records, tip_of, attr, clv_buffers):
    part = lib.partition(tips=tree.tips, clv_buffers=clv_buffers, states=20, sites=len(records[0][1]),
                         rate_matrices=4, prob_matrices=2 * tree.tips - 3, rate_cats=4,
                         scale_buffers=clv_buffers, attributes=attr)
    for head, seq, _ in records:
        part.set_tip_states(tip_of[head], seq.encode())
    return part


def _lg4(lib, part, which):
    for i in range(4):
        part.set_frequencies(i, lib.aa_table(f"pll_aa_freqs_{which}", (4, 20))[i])
        part.set_subst_params(i, lib.aa_table(f"pll_aa_rates_{which}", (4, 190))[i])


@pytest.mark.parametrize("variant", ["tv", "notv"])
def test_lg4_example_end_to_end(gpu_lib, tmp_path, variant):
    lib, tree, records, tip_of = _load(gpu_lib, tmp_path)
    attr = PLL_ATTRIB_ARCH_GPU | (PLL_ATTRIB_PATTERN_TIP if variant == "tv" else 0)
    part = _partition(lib, tree, records, tip_of, attr, tree.inner)
    ops, mi, bl = tree.operations()
    r = tree.root.contents
    edge = (r.clv_index, r.scaler_index, r.back.contents.clv_index, r.back.contents.scaler_index, r.pmatrix_index)
    pidx = np.arange(4, dtype=np.uint32)
    want = LG4["expect"][variant]

    part.set_category_rates(lib.gamma_rates(1.0, 4))
    _lg4(lib, part, "lg4m")
    part.update_prob_matrices(pidx, mi, bl)
    part.update_partials(ops)
    got = part.edge_loglikelihood(*edge, pidx)
    assert abs(got - want["lg4m"]) <= RTOL * abs(want["lg4m"]), (got, want["lg4m"])

    _lg4(lib, part, "lg4x")
    part.set_category_rates([0.498991136, 0.563680734, 0.808264032, 1.887769458])
    part.set_category_weights([0.209224645, 0.224707726, 0.277599198, 0.288468431])
    part.update_prob_matrices(pidx, mi, bl)
    part.update_partials(ops)
    got = part.edge_loglikelihood(*edge, pidx)
    assert abs(got - want["lg4x"]) <= RTOL * abs(want["lg4x"]), (got, want["lg4x"])
    got = part.edge_loglikelihood(edge[2], edge[3], edge[0], edge[1], edge[4], pidx)
    assert abs(got - want["lg4x_swapped"]) <= RTOL * abs(want["lg4x_swapped"])
    part.destroy()
    tree.destroy()


def test_lg4_example_with_recycled_slots(gpu_lib, tmp_path):
    """3 CLV / scaler slots instead of 19: same log-likelihood, bit for bit, on the device."""
    lib, tree, records, tip_of = _load(gpu_lib, tmp_path)
    attr = PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP
    pidx = np.arange(4, dtype=np.uint32)
    r = tree.root.contents

    def run(part, ops, mi, bl, edge):
        part.set_category_rates(lib.gamma_rates(1.0, 4))
        _lg4(lib, part, "lg4m")
        part.update_prob_matrices(pidx, mi, bl)
        part.update_partials(ops)
        return part.edge_loglikelihood(*edge, pidx)

    full = _partition(lib, tree, records, tip_of, attr, tree.inner)
    ops, mi, bl = tree.operations()
    plain = run(full, ops, mi, bl, (r.clv_index, r.scaler_index, r.back.contents.clv_index,
                                    r.back.contents.scaler_index, r.pmatrix_index))
    full.destroy()
    ops, mi, bl, eclv, esc, used = tree.operations_recycled(8)
    assert used == 3
    small = _partition(lib, tree, records, tip_of, attr, used)
    recycled = run(small, ops, mi, bl, (eclv[0], esc[0], eclv[1], esc[1], r.pmatrix_index))
    small.destroy()
    assert plain == recycled
    assert abs(plain - LG4["expect"]["tv"]["lg4m"]) <= RTOL * abs(plain)
    tree.destroy()
