"""ctypes mirror of the tree / file front-end of pll.h (include/pll.h: pll_unode_t, pll_utree_t,
pll_utree_*, pll_fasta_*).  Works with any library exporting that API with the reference's struct
layouts: a tree parsed by one such library can be handed to another's traversal functions in the
same process - that is how tests/test_utree_cpu.py pins parity with the reference."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .binding import OP_DTYPE, PllError, PllLibrary

PLL_TREE_TRAVERSE_POSTORDER = 1
PLL_TREE_TRAVERSE_PREORDER = 2


class UNode(C.Structure):
    pass


UNode._fields_ = [
    ("label", C.c_char_p),
    ("length", C.c_double),
    ("node_index", C.c_uint),
    ("clv_index", C.c_uint),
    ("scaler_index", C.c_int),
    ("pmatrix_index", C.c_uint),
    ("next", C.POINTER(UNode)),
    ("back", C.POINTER(UNode)),
    ("data", C.c_void_p),
]
UNODE_P = C.POINTER(UNode)


class UTree(C.Structure):
    _fields_ = [
        ("tip_count", C.c_uint),
        ("inner_count", C.c_uint),
        ("edge_count", C.c_uint),
        ("nodes", C.POINTER(UNODE_P)),
    ]


UTREE_P = C.POINTER(UTree)
TRAV_CB = C.CFUNCTYPE(C.c_int, UNODE_P)

_TREE_API = {
    "pll_utree_parse_newick": (UTREE_P, [C.c_char_p]),
    "pll_utree_parse_newick_string": (UTREE_P, [C.c_char_p]),
    "pll_utree_destroy": (None, [UTREE_P, C.c_void_p]),
    "pll_utree_reset_template_indices": (None, [UNODE_P, C.c_uint]),
    "pll_utree_wraptree": (UTREE_P, [UNODE_P, C.c_uint]),
    "pll_utree_export_newick": (C.c_void_p, [UNODE_P, C.c_void_p]),
    "pll_utree_traverse": (C.c_int, [UNODE_P, C.c_int, TRAV_CB, C.POINTER(UNODE_P), C.POINTER(C.c_uint)]),
    "pll_utree_create_operations": (None, [C.POINTER(UNODE_P), C.c_uint, C.POINTER(C.c_double),
                                           C.POINTER(C.c_uint), C.c_void_p, C.POINTER(C.c_uint),
                                           C.POINTER(C.c_uint)]),
    "pll_utree_create_operations_recycled": (C.c_int, [UNODE_P, C.c_uint, C.c_uint, C.POINTER(C.c_double),
                                                       C.POINTER(C.c_uint), C.c_void_p, C.POINTER(C.c_uint),
                                                       C.POINTER(C.c_uint), C.POINTER(C.c_uint),
                                                       C.POINTER(C.c_int), C.POINTER(C.c_uint)]),
    "pll_utree_check_integrity": (C.c_int, [UTREE_P]),
    "pll_utree_show_ascii": (None, [UNODE_P, C.c_int]),
    "pll_utree_clone": (UTREE_P, [UTREE_P]),
    "pll_utree_graph_clone": (UNODE_P, [UNODE_P]),
    "pll_fasta_open": (C.c_void_p, [C.c_char_p, C.POINTER(C.c_uint)]),
    "pll_fasta_getnext": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_long),
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_long), C.POINTER(C.c_long)]),
    "pll_fasta_close": (None, [C.c_void_p]),
}

class Msa(C.Structure):
    _fields_ = [("count", C.c_int), ("length", C.c_int), ("sequence", C.POINTER(C.c_char_p)),
                ("label", C.POINTER(C.c_char_p))]


_TREE_API.update({
    "pll_phylip_open": (C.c_void_p, [C.c_char_p, C.POINTER(C.c_uint)]),
    "pll_phylip_close": (None, [C.c_void_p]),
    "pll_phylip_rewind": (C.c_int, [C.c_void_p]),
    "pll_phylip_parse_sequential": (C.POINTER(Msa), [C.c_void_p]),
    "pll_phylip_parse_interleaved": (C.POINTER(Msa), [C.c_void_p]),
    "pll_msa_destroy": (None, [C.POINTER(Msa)]),
})

_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def bind(lib: PllLibrary) -> PllLibrary:
    """Adds the tree / file functions to a loaded PllLibrary (symbols a library lacks are listed
    in lib.missing; the reference built without bison has no Newick reader)."""
    if getattr(lib, "_trees_bound", False):
        return lib
    for name, (res, args) in _TREE_API.items():
        lib._bind(name, res, args)
    lib._trees_bound = True
    return lib


@TRAV_CB
def full_traversal(_node):
    return 1


class Tree:
    """An unrooted tree owned by `lib` (pll_utree_parse_newick_string .. pll_utree_destroy)."""

    def __init__(self, lib: PllLibrary, newick: str | None = None, path: str | None = None):
        self.lib = bind(lib)
        if newick is not None:
            self.ptr = lib.pll_utree_parse_newick_string(newick.encode())
        else:
            self.ptr = lib.pll_utree_parse_newick(path.encode())
        if not self.ptr:
            raise PllError(f"[{lib.errno()}] {lib.errmsg()}")
        self.t = self.ptr.contents

    def destroy(self):
        if self.ptr:
            self.lib.pll_utree_destroy(self.ptr, None)
            self.ptr = None

    @property
    def tips(self) -> int:
        return self.t.tip_count

    @property
    def inner(self) -> int:
        return self.t.inner_count

    def node(self, i: int):
        return self.t.nodes[i]

    @property
    def root(self):
        """The inner node the Newick string was rooted at (last entry of nodes[])."""
        return self.t.nodes[self.t.tip_count + self.t.inner_count - 1]

    def tip_labels(self) -> list[str]:
        return [self.t.nodes[i].contents.label.decode() for i in range(self.tips)]

    def records(self):
        """(label, length, node_index, clv_index, scaler_index, pmatrix_index) of every record:
        tips, then the three ring records of each inner node."""
        out = []
        for i in range(self.tips + self.inner):
            n = self.t.nodes[i]
            ring = [n] if not n.contents.next else [n, n.contents.next, n.contents.next.contents.next]
            for r in ring:
                c = r.contents
                out.append((c.label.decode() if c.label else None, c.length, c.node_index, c.clv_index,
                            c.scaler_index, c.pmatrix_index))
        return out

    def traverse(self, lib: PllLibrary | None = None, root=None, order=PLL_TREE_TRAVERSE_POSTORDER, cb=full_traversal):
        """Traversal buffer computed by `lib` (default: the owner) - a ctypes array of node pointers."""
        lib = bind(lib or self.lib)
        buf = (UNODE_P * (self.tips + self.inner))()
        n = C.c_uint(0)
        if not lib.pll_utree_traverse(root or self.root, order, cb, buf, C.byref(n)):
            raise PllError(lib.errmsg())
        return buf, n.value

    def operations(self, lib: PllLibrary | None = None, root=None):
        """(ops, matrix_indices, branch_lengths) of a full post-order traversal, from `lib`."""
        lib = bind(lib or self.lib)
        buf, n = self.traverse(lib, root)
        branches = np.zeros(2 * self.tips - 3)
        matrices = np.zeros(2 * self.tips - 3, dtype=np.uint32)
        ops = np.zeros(self.inner, dtype=OP_DTYPE)
        nm, no = C.c_uint(0), C.c_uint(0)
        lib.pll_utree_create_operations(buf, n, branches.ctypes.data_as(C.POINTER(C.c_double)),
                                        matrices.ctypes.data_as(C.POINTER(C.c_uint)), ops.ctypes.data,
                                        C.byref(nm), C.byref(no))
        return ops[:no.value], matrices[:nm.value], branches[:nm.value]

    def operations_recycled(self, max_slots: int, root=None):
        """(ops, matrix_indices, branch_lengths, edge_clv[2], edge_scaler[2], slots_used)."""
        lib = self.lib
        branches = np.zeros(2 * self.tips - 3)
        matrices = np.zeros(2 * self.tips - 3, dtype=np.uint32)
        ops = np.zeros(self.inner, dtype=OP_DTYPE)
        nm, no, used = C.c_uint(0), C.c_uint(0), C.c_uint(0)
        eclv, esc = (C.c_uint * 2)(), (C.c_int * 2)()
        ok = lib.pll_utree_create_operations_recycled(root or self.root, self.tips, max_slots,
                                                      branches.ctypes.data_as(C.POINTER(C.c_double)),
                                                      matrices.ctypes.data_as(C.POINTER(C.c_uint)), ops.ctypes.data,
                                                      C.byref(nm), C.byref(no), eclv, esc, C.byref(used))
        if not ok:
            raise PllError(f"[{lib.errno()}] {lib.errmsg()} (needs {used.value} slots)")
        return ops[:no.value], matrices[:nm.value], branches[:nm.value], list(eclv), list(esc), used.value

    def clone(self, graph_lib: PllLibrary | None = None) -> "Tree":
        """A deep copy.  With `graph_lib` the node graph is copied by THAT library's
        pll_utree_graph_clone and wrapped by the owner's pll_utree_wraptree (the reference built
        without bison has no wraptree); both libraries allocate with the same malloc."""
        if graph_lib is None:
            ptr = self.lib.pll_utree_clone(self.ptr)
        else:
            root = bind(graph_lib).pll_utree_graph_clone(self.root)
            ptr = self.lib.pll_utree_wraptree(root, self.tips) if root else None
        if not ptr:
            raise PllError(self.lib.errmsg())
        t = Tree.__new__(Tree)
        t.lib, t.ptr, t.t = self.lib, ptr, ptr.contents
        return t

    def check_integrity(self, lib: PllLibrary | None = None) -> bool:
        return bool(bind(lib or self.lib).pll_utree_check_integrity(self.ptr))

    def export_newick(self, lib: PllLibrary | None = None) -> str:
        lib = bind(lib or self.lib)
        p = lib.pll_utree_export_newick(self.root, None)
        if not p:
            raise PllError(lib.errmsg())
        s = C.string_at(p).decode()
        _libc.free(p)
        return s


def read_fasta(lib: PllLibrary, path: str, status_table=None):
    """[(header, sequence)] read by `lib`'s FASTA reader, plus (stripped_count, errno at the end)."""
    bind(lib)
    table = status_table if status_table is not None else (C.c_uint * 256).in_dll(lib.dll, "pll_map_fasta")
    fd = lib.pll_fasta_open(path.encode(), table)
    if not fd:
        raise PllError(f"[{lib.errno()}] {lib.errmsg()}")
    out = []
    head, seq = C.c_void_p(), C.c_void_p()
    hl, sl, no = C.c_long(), C.c_long(), C.c_long()
    while lib.pll_fasta_getnext(fd, C.byref(head), C.byref(hl), C.byref(seq), C.byref(sl), C.byref(no)):
        out.append((C.string_at(head, hl.value).decode(), C.string_at(seq, sl.value).decode(), no.value))
        _libc.free(head)
        _libc.free(seq)
    err = lib.errno()
    lib.pll_fasta_close(fd)
    return out, err


def read_phylip(lib: PllLibrary, path: str, interleaved: bool = False, twice: bool = False):
    """([(label, sequence)], errno) read by `lib`'s PHYLIP reader; ([], errno) if it rejects the
    file.  `twice` parses, rewinds and parses again (exercises pll_phylip_rewind)."""
    bind(lib)
    table = (C.c_uint * 256).in_dll(lib.dll, "pll_map_phylip")
    C.c_int.in_dll(lib.dll, "pll_errno").value = 0
    fd = lib.pll_phylip_open(path.encode(), table)
    if not fd:
        return [], lib.errno()
    parse = lib.pll_phylip_parse_interleaved if interleaved else lib.pll_phylip_parse_sequential
    msa = parse(fd)
    if twice and msa:
        lib.pll_msa_destroy(msa)
        assert lib.pll_phylip_rewind(fd)
        msa = parse(fd)
    out = []
    if msa:
        m = msa.contents
        out = [(m.label[i].decode(), m.sequence[i].decode()) for i in range(m.count)]
        assert all(len(s) == m.length for _, s in out)
        lib.pll_msa_destroy(msa)
    err = lib.errno()
    lib.pll_phylip_close(fd)
    return out, err


# ---- synthetic trees and a moving virtual root (benchmarks, tests) -----------------------------
def random_newick(tips: int, seed: int, caterpillar: bool = False, labels: bool = True) -> str:
    """Random-join (Yule-like) unrooted binary tree as a Newick string with tips t0..t{tips-1} and
    branch lengths U(0.01, 0.3); `caterpillar` gives the ladder instead."""
    rng = np.random.default_rng(seed)
    if caterpillar and tips > 3:
        lens = rng.uniform(0.01, 0.3, 2 * tips)
        inner = ("(" * (tips - 3) + f"t0:{lens[0]:.6f},t1:{lens[1]:.6f}"
                 + "".join(f"):{lens[tips + i]:.6f},t{i}:{lens[i]:.6f}" for i in range(2, tips - 2))
                 + f"):{lens[tips]:.6f}")
        return f"({inner},t{tips - 2}:{lens[tips - 2]:.6f},t{tips - 1}:{lens[tips - 1]:.6f});"
    nodes = [f"t{i}:{rng.uniform(0.01, 0.3):.6f}" for i in range(tips)]
    if not labels:
        nodes = [f"{i}:{rng.uniform(0.01, 0.3):.6f}" for i in range(tips)]
    while len(nodes) > 3:
        if caterpillar:
            a, b = nodes.pop(0), nodes.pop(0)
            nodes.insert(0, f"({a},{b}):{rng.uniform(0.01, 0.3):.6f}")
        else:
            i = int(rng.integers(0, len(nodes)))
            a = nodes.pop(i)
            j = int(rng.integers(0, len(nodes)))
            b = nodes.pop(j)
            nodes.append(f"({a},{b}):{rng.uniform(0.01, 0.3):.6f}")
    return f"({nodes[0]},{nodes[1]},{nodes[2]});"


class RootWalker:
    """Moves the virtual root of an unrooted tree the way a tree search does (reference
    test/src/partial-traversal.c, examples/partial-traversal/partial.c:373-431): every move
    traverses only the subtrees whose CLVs are not yet oriented towards the new root
    (pll_utree_traverse with a pruning callback) and returns that PARTIAL operations list with the
    P-matrices it needs.  `edge(root)` names the evaluation edge of a root record."""

    def __init__(self, lib: PllLibrary, tree: Tree):
        self.lib, self.tree = bind(lib), tree
        self.oriented: dict[int, bool] = {}
        n = 2 * tree.tips - 3
        self._branches = np.zeros(n)
        self._matrices = np.zeros(n, dtype=np.uint32)
        self._ops = np.zeros(tree.inner, dtype=OP_DTYPE)

        @TRAV_CB
        def partial(node):
            rec = node.contents
            if not rec.next:
                return 1
            me = C.addressof(rec)
            if self.oriented.get(me):
                return 0
            self.oriented[me] = True
            self.oriented[C.addressof(rec.next.contents)] = False
            self.oriented[C.addressof(rec.next.contents.next.contents)] = False
            return 1

        self._cb = partial

    def move(self, root):
        """(ops, matrix_indices, branch_lengths) that re-orient the tree towards `root`."""
        buf, n = self.tree.traverse(self.lib, root, cb=self._cb)
        nm, no = C.c_uint(0), C.c_uint(0)
        self.lib.pll_utree_create_operations(buf, n, self._branches.ctypes.data_as(C.POINTER(C.c_double)),
                                             self._matrices.ctypes.data_as(C.POINTER(C.c_uint)),
                                             self._ops.ctypes.data, C.byref(nm), C.byref(no))
        return self._ops[:no.value].copy(), self._matrices[:nm.value].copy(), self._branches[:nm.value].copy()

    def full(self, root):
        self.oriented.clear()
        return self.move(root)

    @staticmethod
    def edge(root):
        """(parent_clv, parent_scaler, child_clv, child_scaler, matrix) of the edge at `root`."""
        r = root.contents
        b = r.back.contents
        return r.clv_index, r.scaler_index, b.clv_index, b.scaler_index, r.pmatrix_index
