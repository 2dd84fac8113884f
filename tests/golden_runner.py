"""Interpreter for the golden cases of tests/golden/cases.json: drives ANY library exposing the
pll.h API (the GPU product, the reference) or the oracle port through the same step list.

Step vocabulary (see tests/golden/make_golden.py): pmatrix, partials, get_pmatrix, get_clv,
edge, root, pinv, sumtable, derivs, newton.  Outputs are plain lists/dicts (JSON-able)."""
from __future__ import annotations

import numpy as np

from libpll_b200.binding import OP_DTYPE, PLL_ATTRIB_PATTERN_TIP


def _ops(rows):
    a = np.zeros(len(rows), dtype=OP_DTYPE)
    for i, r in enumerate(rows):
        a[i] = tuple(r)
    return a


def execute(lib, case, attributes):
    part = lib.partition(tips=case["tips"], clv_buffers=case["clv_buffers"], states=case["states"],
                         sites=case["sites"], rate_matrices=case["rate_matrices"],
                         prob_matrices=case["prob_matrices"], rate_cats=case["rate_cats"],
                         scale_buffers=case["scale_buffers"], attributes=attributes | case.get("extra_attributes", 0))
    for i, (f, s) in enumerate(zip(case["freqs"], case["subst"])):
        part.set_frequencies(i, f)
        part.set_subst_params(i, s)
    # category rates are INPUTS of the hot path.  Libraries with their own C implementation of
    # pll_compute_gamma_cats compute them (that code is under test too); the oracle port, whose
    # scipy-based discretisation is exact where the reference's AS 91 quantiles are accurate to
    # ~1e-7, is fed the doubles the reference produced so that P-matrices are comparable to 1e-10
    if getattr(lib, "use_case_rates", False):
        part.set_category_rates(case["rates"])
    else:
        part.set_category_rates(lib.gamma_rates(case["alpha"], case["rate_cats"]))
    amap = None
    if "map" in case:  # alphabets other than DNA / amino acids carry their own char -> state-set map
        amap = lib.make_map(case["map"])
    for t, seq in enumerate(case["seqs"]):
        if amap is None:
            part.set_tip_states(t, seq.encode())
        else:
            part.set_tip_states(t, seq.encode(), amap)
    out = _run_steps(part, case)
    part.destroy()
    return out


def _run_steps(part, case):
    out = []
    tables = {}
    for st in case["steps"]:
        do = st["do"]
        if do == "pmatrix":
            part.update_prob_matrices(st["params"], st["matrices"], st["lengths"])
        elif do == "partials":
            part.update_partials(_ops(st["ops"]))
        elif do == "get_pmatrix":
            out.append(dict(kind="pmatrix", index=st["index"], values=part.get_pmatrix(st["index"])[..., :case["states"]].reshape(-1).tolist()))
        elif do == "get_clv":
            out.append(dict(kind="clv", index=st["index"], values=part.get_clv(st["index"])[..., :case["states"]].reshape(-1).tolist()))
        elif do == "edge":
            ps = np.zeros(case["sites"])
            a = st["args"]
            logl = part.edge_loglikelihood(a[0], a[1], a[2], a[3], a[4], st["freqs_indices"], persite=ps)
            out.append(dict(kind="edge", tag=st.get("tag"), logl=logl, persite=ps.tolist()))
        elif do == "root":
            ps = np.zeros(case["sites"])
            logl = part.root_loglikelihood(st["clv"], st["scaler"], st["freqs_indices"], persite=ps)
            out.append(dict(kind="root", tag=st.get("tag"), logl=logl, persite=ps.tolist()))
        elif do == "asc_type":
            part.set_asc_bias_type(st["value"])
        elif do == "asc_weights":
            part.set_asc_state_weights(st["value"])
        elif do == "pinv":
            part.update_invariant_sites_proportion(st["index"], st["value"])
        elif do == "sumtable":
            tables[st["key"]] = part.new_sumtable()
            e = st["edge"]
            part.update_sumtable(e[0], e[1], e[2], e[3], st["params"], tables[st["key"]])
            tables[st["key"] + "_edge"] = e
        elif do == "derivs":
            e = tables[st["key"] + "_edge"]
            d1, d2 = part.likelihood_derivatives(e[2], e[3], st["t"], st["params"], tables[st["key"]])
            out.append(dict(kind="derivs", key=st["key"], t=st["t"], d_f=d1, dd_f=d2))
        elif do == "newton":
            # reference examples/newton/newton.c:31-100
            e = st["edge"]
            table = part.new_sumtable()
            part.update_sumtable(e[0], e[1], e[2], e[3], st["params"], table)
            length, its, trace = st["start"], 0, []
            for _ in range(st["max_iter"]):
                d1, d2 = part.likelihood_derivatives(e[2], e[3], length, st["params"], table)
                its += 1
                trace.append([length, d1, d2])
                if abs(d1) < st["eps"]:
                    break
                length -= d1 / d2
            out.append(dict(kind="newton", final=length, iterations=its, trace=trace))
        else:
            raise ValueError(do)
    return out
