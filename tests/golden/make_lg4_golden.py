#!/usr/bin/env python3
"""Generates tests/golden/lg4_example.json: the reference's data-driven example
(examples/lg4/lg4.c:64-430 on examples/lg4/data/example.{tree,fas}: 21 taxa x 113 amino-acid
sites, LG4M then LG4X) recorded from the REFERENCE itself (oracle/_ref).  Run in the build
container; the JSON travels to the GPU box, where /root/reference does not exist.

The reference's Newick reader is flex/bison generated and cannot be built here, so the tree is
read by the small independent parser below (plain recursive descent + the index template of
reference src/parse_utree.y:250-360), turned into a pll_unode_t graph through ctypes, and
handed to the reference's own pll_utree_traverse / pll_utree_create_operations
(src/utree.c:284-442) and FASTA reader (src/fasta.c).  Everything downstream (P-matrices, CLV
updates, edge log-likelihoods) is the reference's AVX2 path through the pll.h API.

The fixture pins, for this repository's library: the Newick reader + index template
(records), traversal -> operations, the FASTA reader, and the LG4M / LG4X log-likelihoods.

usage: python tests/golden/make_lg4_golden.py
"""
import ctypes as C
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from libpll_b200 import trees as T  # noqa: E402
from libpll_b200.binding import (OP_DTYPE, PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_PATTERN_TIP, PllLibrary)  # noqa: E402

DATA = "/root/reference/examples/lg4/data"
ref = T.bind(PllLibrary(os.path.join(ROOT, "oracle", "_ref", "libpll_ref.so"), is_gpu=False))


# ---- independent Newick reader (strictly binary, ternary root) --------------------------------
def parse_newick(text):
    toks = re.findall(r"[(),:;]|[^\s(),:;]+", text)
    pos = 0

    def subtree():
        nonlocal pos
        if toks[pos] == "(":
            pos += 1
            kids = [subtree()]
            while toks[pos] == ",":
                pos += 1
                kids.append(subtree())
            assert toks[pos] == ")"
            pos += 1
            node = {"kids": kids, "label": None, "length": 0.0}
        else:
            node = {"kids": [], "label": toks[pos], "length": 0.0}
            pos += 1
            if toks[pos] == ":":
                node["length"] = float(toks[pos + 1])
                pos += 2
            return node
        if toks[pos] not in ",):;":
            node["label"] = toks[pos]
            pos += 1
        if toks[pos] == ":":
            node["length"] = float(toks[pos + 1])
            pos += 2
        return node

    root = subtree()
    assert toks[pos] == ";" and len(root["kids"]) == 3
    return root


def template(root):
    """records in the order of pll_utree_t.nodes: (label, length, node_index, clv, scaler, pmatrix)
    per record - tips, then three ring records per inner node, the root last."""
    tips = []

    def count(n):
        if not n["kids"]:
            tips.append(n)
        for k in n["kids"]:
            count(k)

    count(root)
    T_ = len(tips)
    state = {"tip": 0, "clv": T_, "scaler": 0, "node": T_}
    tip_recs, inner_recs = [], []

    def assign(n):
        if not n["kids"]:
            n["clv"] = n["pm"] = state["tip"]
            tip_recs.append((n["label"], n["length"], state["tip"], state["tip"], -1, state["tip"]))
            state["tip"] += 1
            return
        assert len(n["kids"]) == 2
        for k in n["kids"]:
            assign(k)
        n["clv"] = n["pm"] = state["clv"]
        inner_recs.append((n["label"], n["length"], state["node"], state["clv"], state["scaler"], state["clv"]))
        for j, k in enumerate(n["kids"]):
            inner_recs.append((n["label"], k["length"], state["node"] + 1 + j, state["clv"], state["scaler"], k["pm"]))
        state["clv"] += 1
        state["scaler"] += 1
        state["node"] += 3

    for k in root["kids"]:
        assign(k)
    for j, k in enumerate(root["kids"]):
        inner_recs.append((root["label"], k["length"], state["node"] + j, state["clv"], state["scaler"], k["pm"]))
    return T_, tip_recs + inner_recs


# ---- the same tree as a pll_unode_t graph for the reference's traversal code -------------------
def build_graph(root, keep):
    def rec():
        n = T.UNode()
        keep.append(n)
        return n

    def link(a, b, length):
        a.back = C.pointer(b)
        b.back = C.pointer(a)
        a.length = b.length = length

    def build(n, counters):
        if not n["kids"]:
            r = rec()
            r.label = n["label"].encode()
            r.node_index = r.clv_index = r.pmatrix_index = n["clv"]
            r.scaler_index = -1
            return r
        kids = [build(k, counters) for k in n["kids"]]
        ring = [rec(), rec(), rec()]
        for i in range(3):
            ring[i].next = C.pointer(ring[(i + 1) % 3])
            ring[i].clv_index = n["clv"]
            ring[i].scaler_index = n["clv"] - counters["T"]
        ring[0].pmatrix_index = n["pm"]
        for j, k in enumerate(kids):
            link(ring[1 + j], k, n["kids"][j]["length"])
            ring[1 + j].pmatrix_index = n["kids"][j]["pm"]
        return ring[0]

    return build


def main():
    newick = open(os.path.join(DATA, "example.tree")).read()
    root = parse_newick(newick)
    tip_count, records = template(root)

    keep = []
    counters = {"T": tip_count}
    kids = [build_graph(root, keep)(k, counters) for k in root["kids"]]
    ring = [T.UNode(), T.UNode(), T.UNode()]
    keep.extend(ring)
    root_clv = tip_count + tip_count - 3
    for i in range(3):
        ring[i].next = C.pointer(ring[(i + 1) % 3])
        ring[i].clv_index = root_clv
        ring[i].scaler_index = tip_count - 3
        ring[i].back = C.pointer(kids[i])
        kids[i].back = C.pointer(ring[i])
        ring[i].length = kids[i].length = root["kids"][i]["length"]
        ring[i].pmatrix_index = root["kids"][i]["pm"]

    n_nodes = 2 * tip_count - 2
    buf = (T.UNODE_P * n_nodes)()
    n = C.c_uint(0)
    assert ref.pll_utree_traverse(C.pointer(ring[0]), T.PLL_TREE_TRAVERSE_POSTORDER, T.full_traversal, buf, C.byref(n))
    branches = np.zeros(2 * tip_count - 3)
    matrices = np.zeros(2 * tip_count - 3, dtype=np.uint32)
    ops = np.zeros(tip_count - 2, dtype=OP_DTYPE)
    nm, no = C.c_uint(0), C.c_uint(0)
    ref.pll_utree_create_operations(buf, n, branches.ctypes.data_as(C.POINTER(C.c_double)),
                                    matrices.ctypes.data_as(C.POINTER(C.c_uint)), ops.ctypes.data,
                                    C.byref(nm), C.byref(no))
    assert nm.value == 2 * tip_count - 3 and no.value == tip_count - 2

    fasta, err = T.read_fasta(ref, os.path.join(DATA, "example.fas"))
    assert err == 102 and len(fasta) == tip_count  # PLL_ERROR_FILE_EOF
    label_to_tip = {r[0]: r[3] for r in records[:tip_count]}

    # ---- the likelihood part of examples/lg4/lg4.c:209-423 on the reference -----------------
    expect = {}
    for tag, attr in (("tv", PLL_ATTRIB_ARCH_AVX2 | PLL_ATTRIB_PATTERN_TIP), ("notv", PLL_ATTRIB_ARCH_AVX2)):
        part = ref.partition(tips=tip_count, clv_buffers=tip_count - 2, states=20, sites=len(fasta[0][1]),
                             rate_matrices=4, prob_matrices=2 * tip_count - 3, rate_cats=4,
                             scale_buffers=tip_count - 2, attributes=attr)
        for head, seq, _ in fasta:
            part.set_tip_states(label_to_tip[head], seq.encode())
        pidx = np.arange(4, dtype=np.uint32)
        edge = (root_clv, tip_count - 3, int(kids[0].clv_index), int(kids[0].scaler_index), int(ring[0].pmatrix_index))
        out = {}
        part.set_category_rates(ref.gamma_rates(1.0, 4))
        for i in range(4):
            part.set_frequencies(i, ref.aa_table("pll_aa_freqs_lg4m", (4, 20))[i])
            part.set_subst_params(i, ref.aa_table("pll_aa_rates_lg4m", (4, 190))[i])
        part.update_prob_matrices(pidx, matrices, branches)
        part.update_partials(ops)
        out["lg4m"] = part.edge_loglikelihood(*edge, pidx)
        for i in range(4):
            part.set_frequencies(i, ref.aa_table("pll_aa_freqs_lg4x", (4, 20))[i])
            part.set_subst_params(i, ref.aa_table("pll_aa_rates_lg4x", (4, 190))[i])
        part.set_category_rates([0.498991136, 0.563680734, 0.808264032, 1.887769458])
        part.set_category_weights([0.209224645, 0.224707726, 0.277599198, 0.288468431])
        part.update_prob_matrices(pidx, matrices, branches)
        part.update_partials(ops)
        out["lg4x"] = part.edge_loglikelihood(*edge, pidx)
        out["lg4x_swapped"] = part.edge_loglikelihood(edge[2], edge[3], edge[0], edge[1], edge[4], pidx)
        expect[tag] = out
        part.destroy()
        print(tag, out)

    golden = dict(newick=newick, fasta_text=open(os.path.join(DATA, "example.fas")).read(),
                  tip_count=tip_count, records=records, fasta=fasta,
                  ops=[[int(x) for x in o] for o in ops.tolist()], matrices=matrices.tolist(),
                  branches=branches.tolist(), edge=list(edge), expect=expect)
    path = os.path.join(ROOT, "tests", "golden", "lg4_example.json")
    json.dump(golden, open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
