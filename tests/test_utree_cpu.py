"""CPU tests of the caller side of the path: Newick reader, index
template, traversal, traversal -> operations, slot recycling, FASTA reader - the host C code of
libpll_b200/csrc/host/pll_utree.c and pll_fasta.c.  No GPU is needed (no partition is created
on the product library here).

Pins: (1) tests/golden/lg4_example.json, recorded from the reference on its own example data;
(2) the reference's pll_utree_traverse / pll_utree_create_operations / pll_utree_export_newick /
FASTA reader (oracle/_ref) run IN PROCESS on trees parsed by this library - the struct layouts
are identical, so operation lists must match bit for bit; (3) the reference's likelihood of
recycled-slot operation lists, which must be bit-identical to the plain list's."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

import libpll_b200
from libpll_b200 import trees as T
from libpll_b200.binding import PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_PATTERN_TIP, PllError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LG4 = json.load(open(os.path.join(ROOT, "tests", "golden", "lg4_example.json")))


@pytest.fixture(scope="module")
def lib():
    return T.bind(libpll_b200.load())


@pytest.fixture(scope="module")
def ref(ref_lib):
    return T.bind(ref_lib)


from libpll_b200.trees import random_newick  # noqa: E402,F401  (other test modules import it from here)


# ---- golden: the reference's lg4 example ----------------------------------------------------
def test_newick_reader_and_template_match_golden(lib):
    t = T.Tree(lib, newick=LG4["newick"])
    assert t.tips == LG4["tip_count"] and t.inner == t.tips - 2 and t.t.edge_count == 2 * t.tips - 3
    got = t.records()
    assert len(got) == len(LG4["records"])
    for g, w in zip(got, LG4["records"]):
        assert g[0] == w[0] and g[1] == w[1] and list(g[2:]) == list(w[2:]), (g, w)
    t.destroy()


def test_operations_match_golden(lib):
    t = T.Tree(lib, newick=LG4["newick"])
    ops, mi, bl = t.operations()
    assert [list(int(x) for x in o) for o in ops.tolist()] == LG4["ops"]
    assert mi.tolist() == LG4["matrices"]
    assert bl.tolist() == LG4["branches"]
    r = t.root.contents
    assert [r.clv_index, r.scaler_index, r.back.contents.clv_index, r.back.contents.scaler_index,
            r.pmatrix_index] == LG4["edge"]
    t.destroy()


def test_fasta_reader_matches_golden(lib, tmp_path):
    p = tmp_path / "example.fas"
    p.write_text(LG4["fasta_text"])
    got, err = T.read_fasta(lib, str(p))
    assert err == 102  # PLL_ERROR_FILE_EOF ends the loop, as in examples/lg4/lg4.c:172-174
    assert [list(x) for x in got] == LG4["fasta"]


# ---- in-process parity with the reference's tree functions ----------------------------------
@pytest.mark.parametrize("tips,seed,caterpillar", [(3, 1, False), (4, 2, False), (17, 3, False), (200, 4, False),
                                                   (1000, 5, False), (300, 6, True)])
def test_traversal_and_operations_match_reference(lib, ref, tips, seed, caterpillar):
    t = T.Tree(lib, newick=random_newick(tips, seed, caterpillar))
    for root_index in {tips + t.inner - 1, tips, tips + t.inner // 2}:
        root = t.node(root_index)
        for order in (T.PLL_TREE_TRAVERSE_POSTORDER, T.PLL_TREE_TRAVERSE_PREORDER):
            a, na = t.traverse(lib, root, order)
            b, nb = t.traverse(ref, root, order)
            assert na == nb == 2 * tips - 2
            assert [C.addressof(a[i].contents) for i in range(na)] == [C.addressof(b[i].contents) for i in range(nb)]
        o1, m1, b1 = t.operations(lib, root)
        o2, m2, b2 = t.operations(ref, root)
        assert o1.tobytes() == o2.tobytes() and m1.tobytes() == m2.tobytes() and b1.tobytes() == b2.tobytes()
    assert t.export_newick(lib) == t.export_newick(ref)
    t.destroy()


def test_partial_traversal_callback(lib, ref):
    """cbtrav prunes subtrees (reference examples/partial-traversal/partial.c:60-100)."""
    t = T.Tree(lib, newick=random_newick(60, 11))

    @T.TRAV_CB
    def only_small_clv(node):
        n = node.contents
        return 1 if (not n.next) or n.clv_index % 5 != 3 else 0

    for order in (T.PLL_TREE_TRAVERSE_POSTORDER, T.PLL_TREE_TRAVERSE_PREORDER):
        a, na = t.traverse(lib, order=order, cb=only_small_clv)
        b, nb = t.traverse(ref, order=order, cb=only_small_clv)
        assert na == nb and 0 < na < 2 * 60 - 2
        assert [C.addressof(a[i].contents) for i in range(na)] == [C.addressof(b[i].contents) for i in range(nb)]
    t.destroy()


def test_export_parse_round_trip(lib):
    nw = random_newick(80, 21)
    t = T.Tree(lib, newick=nw)
    out = t.export_newick()
    t2 = T.Tree(lib, newick=out)
    assert t2.export_newick() == out
    assert t2.tip_labels() == t.tip_labels()
    assert [r[2:] for r in t2.records()] == [r[2:] for r in t.records()]
    t.destroy()
    t2.destroy()


def test_clone_and_integrity_match_reference(lib, ref):
    t = T.Tree(lib, newick=random_newick(150, 31))
    assert t.check_integrity(lib) and t.check_integrity(ref)
    ours, theirs = t.clone(), t.clone(graph_lib=ref)
    for c in (ours, theirs):
        assert c.check_integrity(lib) and c.check_integrity(ref)
    assert ours.records() == theirs.records() == t.records()
    assert ours.export_newick() == t.export_newick()
    o1, m1, b1 = ours.operations()
    o0, m0, b0 = t.operations()
    assert o1.tobytes() == o0.tobytes() and m1.tobytes() == m0.tobytes() and b1.tobytes() == b0.tobytes()
    # a clone is independent of the original
    t.node(0).contents.length = 123.0
    assert ours.node(0).contents.length != 123.0
    assert not t.check_integrity(lib) and not t.check_integrity(ref)  # the edge's two ends now disagree
    ours.destroy()
    theirs.destroy()
    t.destroy()


def _captured_stdout(fn, tmp_path, name):
    """Runs fn() with the C-level stdout (fd 1) redirected to a file; returns what was printed."""
    libc = C.CDLL(None)
    path = str(tmp_path / name)
    libc.fflush(None)
    saved = os.dup(1)
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
    os.dup2(fd, 1)
    try:
        fn()
        libc.fflush(None)
    finally:
        os.dup2(saved, 1)
        os.close(fd)
        os.close(saved)
    return open(path).read()


@pytest.mark.parametrize("tips,seed,caterpillar", [(3, 1, False), (4, 2, False), (9, 3, False), (60, 4, False),
                                                   (40, 5, True)])
def test_ascii_drawing_matches_reference(lib, ref, tmp_path, tips, seed, caterpillar):
    """pll_utree_show_ascii (reference src/utree.c:122-157): same drawing, character for character,
    from the parse root, from a tip and from another inner node, for every combination of fields."""
    t = T.Tree(lib, newick=random_newick(tips, seed, caterpillar))
    starts = [t.root, t.node(0), t.node(tips)]
    for k, start in enumerate(starts):
        for options in (0, 1, 2, 31, 5, 24):
            ours = _captured_stdout(lambda: lib.pll_utree_show_ascii(start, options), tmp_path, "ours.txt")
            theirs = _captured_stdout(lambda: ref.pll_utree_show_ascii(start, options), tmp_path, "ref.txt")
            assert ours == theirs, (k, options)
            assert ours.count("\n") == 2 * (2 * tips - 3)
    t.destroy()


def test_ascii_drawing_of_a_deep_tree(lib, tmp_path):
    tips = 1_500      # the drawing is quadratic in depth: 6 000 lines of up to 6 000 characters
    t = T.Tree(lib, newick=random_newick(tips, 3, caterpillar=True))
    text = _captured_stdout(lambda: lib.pll_utree_show_ascii(t.root, 1), tmp_path, "deep.txt")
    assert text.count("\n") == 2 * (2 * tips - 3)
    t.destroy()


def test_deep_tree_needs_no_recursion(lib):
    """A 200 000-taxon caterpillar: the reference's recursive walks would need ~200 000 stack
    frames; the reader, template, traversal, operations and export here are all iterative."""
    tips = 200_000
    t = T.Tree(lib, newick=random_newick(tips, 3, caterpillar=True))
    assert t.tips == tips
    ops, mi, bl = t.operations()
    assert len(ops) == tips - 2 and len(mi) == 2 * tips - 3
    assert len(t.export_newick()) > tips * 10
    t.destroy()


def test_lexer_details(lib):
    # quoted labels, numeric labels, labels that start like numbers, inner labels, whitespace
    nw = "( 'a b':0.1 ,\n\"c,d\":2e-1,\t(12:1,1e5x:.5)in1:0.25 )root:7 ;"
    t = T.Tree(lib, newick=nw)
    assert t.tip_labels() == ["a b", "c,d", "12", "1e5x"]
    recs = t.records()
    assert [r[1] for r in recs[:4]] == [0.1, 0.2, 1.0, 0.5]
    assert recs[4][0] == "in1" and recs[4][1] == 0.25
    assert recs[7][0] == "root"
    t.destroy()
    # missing lengths are 0 (examples set them afterwards, lg4.c:37-62)
    t = T.Tree(lib, newick="(a,b,(c,d));")
    assert all(r[1] == 0.0 for r in t.records())
    t.destroy()


@pytest.mark.parametrize("bad", ["(a,b);", "(a,b,c,d);", "(a,b,(c,d,e));", "(a,b,c)", "(a,b,(c,d);", "a;",
                                 "(a,b,c):;", "(a:x,b,c);", "(a,b,[c]);", "((a,b),(c,d));", "", "(a,,b);",
                                 "(a,b,'c);"])
def test_syntax_errors(lib, bad):
    with pytest.raises(PllError) as e:
        T.Tree(lib, newick=bad)
    assert "[111]" in str(e.value)  # PLL_ERROR_NEWICK_SYNTAX


def test_fasta_reader_matches_reference(lib, ref, tmp_path):
    text = (">seq one\r\nACGT acgt\r\nNN--??\r\n"
            ">two|x\nAC*GT!!\nj o J O 12\n\n"
            ">three\n>four\nA\n")
    p = tmp_path / "x.fas"
    p.write_bytes(text.encode())
    a, ea = T.read_fasta(lib, str(p))
    b, eb = T.read_fasta(ref, str(p))
    assert a == b and ea == eb == 102
    assert a[1][1] == "ACGTJO12" and a[2][1] == ""
    # fatal character and bad header: same error codes as the reference
    for content, code in ((">a\nAC.GT\n", 103), (">a\nAC\x01GT\n", 104), ("ACGT\n", 105)):
        p.write_bytes(content.encode())
        _, ea = T.read_fasta(lib, str(p))
        _, eb = T.read_fasta(ref, str(p))
        assert ea == eb == code
    with pytest.raises(PllError):
        T.read_fasta(lib, str(tmp_path / "missing.fas"))
    for name in ("pll_map_fasta", "pll_map_phylip"):
        ours = list((C.c_uint * 256).in_dll(lib.dll, name))
        theirs = list((C.c_uint * 256).in_dll(ref.dll, name))
        assert ours == theirs


# ---- PHYLIP reader -------------------------------------------------------------------------
SEQUENTIAL = """ 4 12
taxon_one  ACGTACGT
  ACGT
second   AC GT AC GT AC GT

third\tACGTAC
GTACGT
fourth acgtnn--??AC
"""
INTERLEAVED = """4 14
one      ACGTA CGT
two      ACGTT CGA
three
  ACGTA CGC
four     NNNNN NNN

CCCCCC
GGGGGG

TTTTTT
AAAAAA
"""


@pytest.mark.parametrize("text,interleaved", [(SEQUENTIAL, False), (INTERLEAVED, True),
                                              (SEQUENTIAL.replace("\n", "\r\n"), False),
                                              (INTERLEAVED.replace("\n", "\r\n"), True)])
def test_phylip_reader_matches_reference(lib, ref, tmp_path, text, interleaved):
    p = tmp_path / "a.phy"
    p.write_bytes(text.encode())
    ours, e1 = T.read_phylip(lib, str(p), interleaved)
    theirs, e2 = T.read_phylip(ref, str(p), interleaved)
    assert ours and ours == theirs
    again, _ = T.read_phylip(lib, str(p), interleaved, twice=True)
    assert again == ours


@pytest.mark.parametrize("text,interleaved,code", [
    ("x 12\na ACGT\n", False, 106),                      # bad header
    ("2 4 extra\na ACGT\nb ACGT\n", False, 106),         # junk after the dimensions
    ("2 4\na ACGT\n", False, 106),                       # too few sequences
    ("1 4\na ACGT\nb ACGT\n", False, 106),               # too many sequences
    ("2 4\na ACGTA\nb ACGT\n", False, 107),              # sequence too long
    ("2 4\na AC\n", False, 106),                         # input ends inside a sequence
    ("2 4\na AC.T\nb ACGT\n", False, 109),               # fatal character
    ("2 4\na AC\x01T\nb ACGT\n", False, 110),            # unprintable character
    ("2 6\na ACG\nb AC\n\nTTT\nGGG\n", True, 108),       # block out of alignment
    ("2 6\na ACG\nb ACG\n\nTTT\n", True, 106),           # incomplete last block
    ("2 8\na ACG\nb ACG\n\nTTT\nGGG\n", True, 0),        # total length differs from the header
])
def test_phylip_errors_match_reference(lib, ref, tmp_path, text, interleaved, code):
    p = tmp_path / "bad.phy"
    p.write_bytes(text.encode())
    ours, e1 = T.read_phylip(lib, str(p), interleaved)
    theirs, e2 = T.read_phylip(ref, str(p), interleaved)
    assert ours == [] and theirs == []
    if code:
        assert e1 == code, (e1, e2)
        # the reference rejects junk after the header dimensions without setting pll_errno
        assert e2 == code or "extra" in text, (e1, e2)


def test_phylip_large_random_files_match_reference(lib, ref, tmp_path):
    rng = np.random.default_rng(8)
    for trial in range(6):
        count, length = int(rng.integers(1, 30)), int(rng.integers(1, 3000))
        seqs = ["".join(rng.choice(list("ACGTN-"), length)) for _ in range(count)]
        labels = [f"t{i}_{'x' * int(rng.integers(0, 12))}" for i in range(count)]
        width = int(rng.integers(5, 200))
        seq_text = f"{count} {length}\n" + "".join(
            f"{l} " + "\n".join(s[k:k + width] for k in range(0, length, width)) + "\n" for l, s in zip(labels, seqs))
        blocks = [f"{count} {length}\n"]
        for k in range(0, length, width):
            blocks.append("".join((f"{l}  " if k == 0 else "") + " ".join(s[k:k + width][q:q + 10] for q in range(0, width, 10))
                                  + "\n" for l, s in zip(labels, seqs)) + "\n")
        for text, inter in ((seq_text, False), ("".join(blocks), True)):
            p = tmp_path / f"r{trial}_{inter}.phy"
            p.write_text(text)
            ours, _ = T.read_phylip(lib, str(p), inter)
            theirs, _ = T.read_phylip(ref, str(p), inter)
            assert ours == theirs == list(zip(labels, seqs))


# ---- slot recycling -------------------------------------------------------------------------
def _check_recycled(ops, tips, slots):
    """every operation reads tips or slots written before and not yet overwritten"""
    holder = {}
    for op in ops:
        for c in (int(op["child1_clv_index"]), int(op["child2_clv_index"])):
            if c >= tips:
                assert holder.get(c) == "live", "child slot was not written or already recycled"
        p = int(op["parent_clv_index"])
        assert tips <= p < tips + slots
        assert int(op["parent_scaler_index"]) == p - tips
        for c in (int(op["child1_clv_index"]), int(op["child2_clv_index"])):
            assert c != p
        holder[p] = "live"


@pytest.mark.parametrize("tips,seed,caterpillar", [(5, 1, False), (64, 2, False), (1000, 3, False), (500, 4, True)])
def test_recycled_operations_structure(lib, tips, seed, caterpillar):
    t = T.Tree(lib, newick=random_newick(tips, seed, caterpillar))
    ops, mi, bl, eclv, esc, used = t.operations_recycled(64)
    assert used <= math.floor(math.log2(tips)) + 2
    if caterpillar:
        assert used <= 3
    assert len(ops) == tips - 2 and len(mi) == 2 * tips - 3
    _check_recycled(ops, tips, used)
    # the same branches as the plain list, possibly in another order
    _, mi0, bl0 = t.operations()
    assert sorted(zip(mi.tolist(), bl.tolist())) == sorted(zip(mi0.tolist(), bl0.tolist()))
    # exactly `used` slots suffice, one fewer does not
    ops2 = t.operations_recycled(used)[0]
    assert ops2.tobytes() == ops.tobytes()
    with pytest.raises(PllError):
        t.operations_recycled(used - 1)
    t.destroy()


def test_recycled_operations_give_identical_likelihood_on_reference(lib, ref):
    """The reference evaluates the plain list (its own pll_utree_create_operations) and the
    recycled list of the same tree: identical log-likelihoods, bit for bit."""
    tips, sites = 120, 300
    t = T.Tree(lib, newick=random_newick(tips, 77))
    rng = np.random.default_rng(5)
    seqs = ["".join(rng.choice(list("ACGT-"), sites)) for _ in range(tips)]

    def evaluate(clv_buffers, ops, mi, bl, edge):
        part = ref.partition(tips=tips, clv_buffers=clv_buffers, states=4, sites=sites, rate_matrices=1,
                             prob_matrices=2 * tips - 3, rate_cats=4, scale_buffers=clv_buffers,
                             attributes=PLL_ATTRIB_ARCH_AVX2 | PLL_ATTRIB_PATTERN_TIP)
        part.set_frequencies(0, [0.3, 0.2, 0.25, 0.25])
        part.set_subst_params(0, [1.2, 3.1, 0.9, 1.1, 3.3, 1.0])
        part.set_category_rates(ref.gamma_rates(0.5, 4))
        for i, s in enumerate(seqs):
            part.set_tip_states(i, s.encode())
        pidx = np.zeros(4, np.uint32)
        part.update_prob_matrices(pidx, mi, bl)
        part.update_partials(ops)
        lnl = part.edge_loglikelihood(edge[0], edge[1], edge[2], edge[3], edge[4], pidx)
        part.destroy()
        return lnl

    r = t.root.contents
    ops, mi, bl = t.operations(ref)
    plain = evaluate(tips - 2, ops, mi, bl, (r.clv_index, r.scaler_index, r.back.contents.clv_index,
                                            r.back.contents.scaler_index, r.pmatrix_index))
    ops, mi, bl, eclv, esc, used = t.operations_recycled(16)
    recycled = evaluate(used, ops, mi, bl, (eclv[0], esc[0], eclv[1], esc[1], r.pmatrix_index))
    assert np.isfinite(plain) and plain == recycled
    t.destroy()


# ---- rooted trees (pll_rtree_*, reference src/parse_rtree.y + src/rtree.c) ------------------
class RNode(C.Structure):
    pass


RNode._fields_ = [("label", C.c_char_p), ("length", C.c_double), ("node_index", C.c_uint), ("clv_index", C.c_uint),
                  ("scaler_index", C.c_int), ("pmatrix_index", C.c_uint), ("left", C.POINTER(RNode)),
                  ("right", C.POINTER(RNode)), ("parent", C.POINTER(RNode)), ("data", C.c_void_p)]


class RTree(C.Structure):
    _fields_ = [("tip_count", C.c_uint), ("inner_count", C.c_uint), ("edge_count", C.c_uint),
                ("nodes", C.POINTER(C.POINTER(RNode))), ("root", C.POINTER(RNode))]


RTRAV_CB = C.CFUNCTYPE(C.c_int, C.POINTER(RNode))


def _bind_rtree(dll, with_parser):
    dll.pll_rtree_traverse.restype = C.c_int
    dll.pll_rtree_traverse.argtypes = [C.POINTER(RNode), C.c_int, RTRAV_CB, C.POINTER(C.POINTER(RNode)),
                                       C.POINTER(C.c_uint)]
    dll.pll_rtree_create_operations.restype = None
    dll.pll_rtree_create_operations.argtypes = [C.POINTER(C.POINTER(RNode)), C.c_uint, C.POINTER(C.c_double),
                                                C.POINTER(C.c_uint), C.c_void_p, C.POINTER(C.c_uint),
                                                C.POINTER(C.c_uint)]
    dll.pll_rtree_export_newick.restype = C.c_void_p
    dll.pll_rtree_export_newick.argtypes = [C.POINTER(RNode), C.c_void_p]
    dll.pll_rtree_show_ascii.restype = None
    dll.pll_rtree_show_ascii.argtypes = [C.POINTER(RNode), C.c_int]
    if with_parser:
        dll.pll_rtree_parse_newick_string.restype = C.POINTER(RTree)
        dll.pll_rtree_parse_newick_string.argtypes = [C.c_char_p]
        dll.pll_rtree_parse_newick.restype = C.POINTER(RTree)
        dll.pll_rtree_parse_newick.argtypes = [C.c_char_p]
        dll.pll_rtree_destroy.restype = None
        dll.pll_rtree_destroy.argtypes = [C.POINTER(RTree), C.c_void_p]
        dll.pll_rtree_wraptree.restype = C.POINTER(RTree)
        dll.pll_rtree_wraptree.argtypes = [C.POINTER(RNode), C.c_uint]
    return dll


def random_rooted_newick(tips, seed, caterpillar=False):
    rng = np.random.default_rng(seed)
    nodes = [f"t{i}:{rng.uniform(0.01, 0.3):.6f}" for i in range(tips)]
    k = 0
    while len(nodes) > 1:
        if caterpillar:
            a, b = nodes.pop(0), nodes.pop(0)
        else:
            a = nodes.pop(int(rng.integers(0, len(nodes))))
            b = nodes.pop(int(rng.integers(0, len(nodes))))
        last = len(nodes) == 0
        label = f"n{k}" if k % 3 == 0 else ""
        k += 1
        new = f"({a},{b}){label}" + ("" if last else f":{rng.uniform(0.01, 0.3):.6f}")
        nodes.insert(0, new) if caterpillar else nodes.append(new)
    return nodes[0] + ";"


def _rtree_views(dll, root, tips, traversal, prune=None):
    """(node_index list of the traversal, ops bytes, matrix indices, branches) computed by `dll`."""
    n = 2 * tips - 1
    buf = (C.POINTER(RNode) * n)()
    size = C.c_uint(0)
    cb = RTRAV_CB(lambda node: 0 if prune is not None and node.contents.node_index in prune else 1)
    assert dll.pll_rtree_traverse(root, traversal, cb, buf, C.byref(size)) == 1
    order = [buf[i].contents.node_index for i in range(size.value)]
    from libpll_b200.binding import OP_DTYPE
    ops = np.zeros(n, dtype=OP_DTYPE)
    branches = np.zeros(n)
    mats = np.zeros(n, dtype=np.uint32)
    mc, oc = C.c_uint(0), C.c_uint(0)
    dll.pll_rtree_create_operations(buf, size.value, branches.ctypes.data_as(C.POINTER(C.c_double)),
                                    mats.ctypes.data_as(C.POINTER(C.c_uint)), ops.ctypes.data, C.byref(mc), C.byref(oc))
    return order, ops[:oc.value].tobytes(), mats[:mc.value].tobytes(), branches[:mc.value].tobytes()


@pytest.mark.parametrize("tips,seed,caterpillar", [(2, 1, False), (3, 2, False), (17, 3, False), (120, 4, False),
                                                   (60, 5, True)])
def test_rooted_tree_functions_match_reference(lib, ref_lib, tmp_path, tips, seed, caterpillar):
    """Trees parsed by this library's rooted reader (the reference's is bison code that is not
    built) handed to the reference's pll_rtree_traverse / _create_operations / _export_newick /
    _show_ascii in process (same struct layout): identical traversals (full and pruned, post- and
    pre-order), operation lists, matrix lists, Newick text and drawings."""
    ours = _bind_rtree(lib.dll, True)
    theirs = _bind_rtree(ref_lib.dll, False)
    text = random_rooted_newick(tips, seed, caterpillar)
    tp = ours.pll_rtree_parse_newick_string(text.encode())
    assert tp, lib.errmsg()
    t = tp.contents
    assert (t.tip_count, t.inner_count, t.edge_count) == (tips, tips - 1, 2 * tips - 2)
    # index template (reference src/parse_rtree.y:167-230): tips 0.. left to right, inner nodes
    # T.. in post-order with scalers 0.., nodes[] in the same numbering, the root last
    for i in range(2 * tips - 1):
        node = t.nodes[i].contents
        assert node.node_index == i == node.clv_index
        assert node.scaler_index == (-1 if i < tips else i - tips)
        assert bool(node.left) == (i >= tips)
        if i < 2 * tips - 2:
            assert node.pmatrix_index == i and node.parent
    assert t.root.contents.node_index == 2 * tips - 2 and t.root.contents.pmatrix_index == 0 and not t.root.contents.parent
    rng = np.random.default_rng(seed)
    for traversal in (1, 2):
        for prune in (None, set(int(x) for x in rng.integers(0, 2 * tips - 2, size=max(1, tips // 5)))):
            assert _rtree_views(ours, t.root, tips, traversal, prune) == _rtree_views(theirs, t.root, tips, traversal, prune)
    a, b = ours.pll_rtree_export_newick(t.root, None), theirs.pll_rtree_export_newick(t.root, None)
    assert C.string_at(a) == C.string_at(b)
    again = ours.pll_rtree_parse_newick_string(C.string_at(a))
    assert again and again.contents.tip_count == tips
    ours.pll_rtree_destroy(again, None)
    for options in (0, 1, 31):
        x = _captured_stdout(lambda: ours.pll_rtree_show_ascii(t.root, options), tmp_path, "ours.txt")
        y = _captured_stdout(lambda: theirs.pll_rtree_show_ascii(t.root, options), tmp_path, "ref.txt")
        assert x == y, options
    # from a file; wraptree with a counted tip number
    path = tmp_path / "rooted.tree"
    path.write_text(text)
    tf = ours.pll_rtree_parse_newick(str(path).encode())
    assert tf and tf.contents.tip_count == tips
    ours.pll_rtree_destroy(tf, None)
    ours.pll_rtree_destroy(tp, None)


@pytest.mark.parametrize("bad", ["(a,b,c);", "(a);", "((a,b),c)", "(a,b));", "(a,,b);", "a;", "(a:x,b);", "((a,b,c),d);",
                                 "", "(a,b)", "(a,(b,c);"])
def test_rooted_syntax_errors(lib, bad):
    ours = _bind_rtree(lib.dll, True)
    assert not ours.pll_rtree_parse_newick_string(bad.encode())
    assert lib.errno() == 111          # PLL_ERROR_NEWICK_SYNTAX


def test_rooted_deep_tree(lib):
    ours = _bind_rtree(lib.dll, True)
    tips = 20_000
    tp = ours.pll_rtree_parse_newick_string(random_rooted_newick(tips, 9, caterpillar=True).encode())
    assert tp and tp.contents.tip_count == tips
    order, ops, mats, _ = _rtree_views(ours, tp.contents.root, tips, 1)
    assert len(order) == 2 * tips - 1 and len(ops) == 32 * (tips - 1) and len(mats) == 4 * (2 * tips - 2)
    ours.pll_rtree_destroy(tp, None)
