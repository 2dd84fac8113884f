#!/usr/bin/env python3
"""Generates tests/golden/cases.json from the REFERENCE itself (oracle/_ref/libpll_ref.so, built
from /root/reference by oracle/Makefile).  Run here, in the build container; the JSON it writes
is committed and travels to the GPU box, where /root/reference does not exist.

Each case restates, through the pll.h API, one of the reference's own self-contained tests or
examples (inputs are embedded in their sources), executes it on the reference with
PLL_ATTRIB_ARCH_AVX2 (with and without PLL_ATTRIB_PATTERN_TIP) and records every output at
full double precision.  Where the reference ships an expected value in text form
(test/out/*.out, examples' documented output) the script ASSERTS the freshly computed number
against that text, so the fixture is pinned to the reference's own golden files:

  case                     restates                                    pinned to
  test_00010_NMDU_lkcalc   test/src/00010_NMDU_lkcalc.c:33-210         test/out/00010_NMDU_lkcalc.out
  test_00011_NMAU_lkcalc   test/src/00011_NMAU_lkcalc.c                test/out/00011_NMAU_lkcalc.out
  example_unrooted         examples/unrooted/unrooted.c:32-212         lnL -33.387713 / -34.550204 / -36.830297
  example_newton           examples/newton/newton.c:31-243             0.6 -> 2.607098 in 7 iterations
  derivatives_grid         test/src/derivatives.c recipe (alpha x pinv x cats x branch lengths)
  test_00012_NMOU_lkcalc   test/src/00012_NMOU_lkcalc.c:33-235 (7 states) test/out/00012_NMOU_lkcalc.out
  derivatives_oddstates    test/src/derivatives-oddstates.c:44-330 (5 st.) test/out/derivatives-oddstates.out

usage: python tests/golden/make_golden.py
"""
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from libpll_b200.binding import (PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_PATTERN_TIP, OP_DTYPE, PllLibrary)  # noqa: E402

REF = "/root/reference"
ref = PllLibrary(os.path.join(ROOT, "oracle", "_ref", "libpll_ref.so"), is_gpu=False)
NONE = -1


def aa(name, shape):
    return ref.aa_table(name, shape).tolist()


def op(parent, ps, c1, m1, s1, c2, m2, s2):
    return [parent, ps, c1, m1, s1, c2, m2, s2]


CASES = []

# ---- test 00010: DNA, 5 taxa x 12 sites, HKY-like GTR, Gamma4 alpha 0.5 -----------------------
CASES.append(dict(
    name="test_00010_NMDU_lkcalc", states=4, tips=5, clv_buffers=4, sites=12, rate_matrices=1,
    prob_matrices=7, rate_cats=4, scale_buffers=0, alpha=0.5,
    freqs=[[0.3, 0.4, 0.1, 0.2]], subst=[[1, 2.5, 1, 1, 2.5, 1]],
    seqs=["WAC-CTA-ATCT", "CCC-TTA-ATGT", "A-C-TAG-CTCT", "CTCTTAA-A-CG", "CAC-TCA-A-TG"],
    steps=[
        dict(do="pmatrix", params=[0, 0, 0, 0], matrices=[0, 1, 2, 3], lengths=[0.1, 0.2, 1, 1]),
        dict(do="partials", ops=[op(5, NONE, 0, 1, NONE, 1, 1, NONE), op(6, NONE, 5, 0, NONE, 2, 1, NONE),
                                 op(7, NONE, 3, 1, NONE, 4, 1, NONE)]),
        dict(do="get_pmatrix", index=0), dict(do="get_pmatrix", index=1),
        dict(do="get_clv", index=5), dict(do="get_clv", index=6), dict(do="get_clv", index=7),
        dict(do="edge", args=[6, NONE, 7, NONE, 0], freqs_indices=[0, 0, 0, 0], tag="inner-inner"),
        dict(do="partials", ops=[op(7, NONE, 6, 0, NONE, 3, 1, NONE)]),
        dict(do="edge", args=[7, NONE, 4, NONE, 1], freqs_indices=[0, 0, 0, 0], tag="tip-inner"),
    ],
    printed={"inner-inner": -58.887310, "tip-inner": -58.887310}, out_file="test/out/00010_NMDU_lkcalc.out"))

# ---- test 00011: amino acids (Dayhoff), 5 taxa x 15 sites ------------------------------------
dayhoff_r = ref.aa_table("pll_aa_rates_dayhoff", (190,)).tolist()
dayhoff_f = ref.aa_table("pll_aa_freqs_dayhoff", (20,)).tolist()
CASES.append(dict(
    name="test_00011_NMAU_lkcalc", states=20, tips=5, clv_buffers=4, sites=15, rate_matrices=1,
    prob_matrices=7, rate_cats=4, scale_buffers=0, alpha=0.5,
    freqs=[dayhoff_f], subst=[dayhoff_r],
    seqs=["PIGLRVTLRRDRMWI", "IQGMDITIVT-----", "--AFALLQKIGMPFE", "MDISIVT------TA", "GLSEQTVFHEIDQDK"],
    steps=None, out_file="test/out/00011_NMAU_lkcalc.out"))

# ---- examples/unrooted and examples/newton: 4 taxa x 6 sites, JC-rates GTR, Gamma4 alpha 1 ------
EX = dict(states=4, tips=4, clv_buffers=2, sites=6, rate_matrices=1, prob_matrices=5, rate_cats=4,
          scale_buffers=2, alpha=1.0, freqs=[[0.17, 0.19, 0.25, 0.39]], subst=[[1, 1, 1, 1, 1, 1]],
          seqs=["WAAAAB", "CACACD", "AGGACA", "CGTAGT"])
EX_PM = dict(do="pmatrix", params=[0, 0, 0, 0], matrices=[0, 1, 2, 3, 4], lengths=[0.2, 0.4, 0.3, 0.5, 0.6])
EX_OPS = dict(do="partials", ops=[op(4, 0, 0, 0, NONE, 1, 1, NONE), op(5, 1, 2, 2, NONE, 3, 3, NONE)])
EX_EDGE = dict(do="edge", args=[4, 0, 5, 1, 4], freqs_indices=[0, 0, 0, 0])
CASES.append(dict(name="example_unrooted", **EX, steps=[
    EX_PM, EX_OPS, dict(do="get_clv", index=4), dict(do="get_clv", index=5),
    dict(EX_EDGE, tag="Log-L"),
    dict(do="pinv", index=0, value=0.5), EX_PM, EX_OPS, dict(EX_EDGE, tag="Log-L (Inv+Gamma 0.5)"),
    dict(do="pinv", index=0, value=0.75), EX_PM, EX_OPS, dict(EX_EDGE, tag="Log-L (Inv+Gamma 0.75)"),
], printed={"Log-L": -33.387713, "Log-L (Inv+Gamma 0.5)": -34.550204, "Log-L (Inv+Gamma 0.75)": -36.830297}))

CASES.append(dict(name="example_newton", **EX, steps=[
    EX_PM, EX_OPS,
    dict(do="newton", edge=[4, 5, 0, 1], params=[0, 0, 0, 0], start=0.6, max_iter=32, eps=1e-5,
         printed_final=2.607098, printed_iterations=7),
]))

# ---- derivatives grid (recipe of reference test/src/derivatives.c:44-48 on the 00010 data) ------
deriv_steps = [
    dict(do="pmatrix", params=[0, 0, 0, 0], matrices=[0, 1, 2, 3], lengths=[0.1, 0.2, 1, 1]),
    dict(do="partials", ops=[op(5, NONE, 0, 1, NONE, 1, 1, NONE), op(6, NONE, 5, 0, NONE, 2, 1, NONE),
                             op(7, NONE, 3, 1, NONE, 4, 1, NONE)]),
    dict(do="sumtable", key="ii", edge=[6, 7, NONE, NONE], params=[0, 0, 0, 0]),
    dict(do="sumtable", key="ti", edge=[6, 3, NONE, NONE], params=[0, 0, 0, 0]),
]
for t in (0.1, 0.5, 0.9, 1.2, 1.5, 1.8, 2.1, 5.0, 25.0):
    deriv_steps.append(dict(do="derivs", key="ii", t=t, params=[0, 0, 0, 0]))
    deriv_steps.append(dict(do="derivs", key="ti", t=t, params=[0, 0, 0, 0]))
for alpha in (0.1, 0.75, 1.5):
    for pinv in (0.0, 0.3, 0.6):
        CASES.append(dict(
            name=f"derivatives_grid_a{alpha}_p{pinv}", states=4, tips=5, clv_buffers=4, sites=12,
            rate_matrices=1, prob_matrices=7, rate_cats=4, scale_buffers=0, alpha=alpha,
            freqs=[[0.3, 0.4, 0.1, 0.2]], subst=[[1, 2.5, 1, 1, 2.5, 1]],
            seqs=["WAC-CTA-ATCT", "CCC-TTA-ATGT", "A-C-TAG-CTCT", "CTCTTAA-A-CG", "CAC-TCA-A-TG"],
            steps=([dict(do="pinv", index=0, value=pinv)] if pinv > 0 else []) + deriv_steps))

# ---- test 00012: 7 states ("odd" alphabet A..G, E ambiguous), generic kernels ------------------
def odd_map(n_states):
    """reference test/src/00012_NMOU_lkcalc.c:32-44 (7 states) and test/src/common.c:8-19 (5 states)"""
    gap = (1 << n_states) - 1 if n_states == 5 else 0x3f
    m = {"*": gap, "-": gap, "?": gap}
    for ch, v in zip("ABCDEFG", (1, 2, 4, 8, 0x0c, 0x10, 0x20)):
        if n_states == 5 and ch in "FG":
            continue
        m[ch] = v
        m[ch.lower()] = v
    return m


CASES.append(dict(
    name="test_00012_NMOU_lkcalc", states=7, tips=5, clv_buffers=4, sites=12, rate_matrices=1,
    prob_matrices=7, rate_cats=4, scale_buffers=0, alpha=0.5, map=odd_map(7),
    freqs=[[0.12, 0.14, 0.13, 0.11, 0.15, 0.13, 0.12]],
    subst=[[0.5, 2.0, 3.0, 4.0, 5.0, 1.1, 1.2, 1.3, 1.4, 1.5, 2.1, 2.2, 2.3, 2.4, 2.5, 3.1, 3.2, 3.3, 3.4, 3.5, 1.0]],
    seqs=["AAB-CCD-EFAA", "ACC-FBA-ABGG", "A-C-GAG-GCCF", "ADCFCAA-A-CG", "ABC-BCA-A-BG"],
    steps=[dict(s) for s in CASES[0]["steps"]],
    printed={"inner-inner": -95.791417, "tip-inner": -95.791417}, out_file="test/out/00012_NMOU_lkcalc.out"))

# ---- derivatives with 5 states: 1, 2 and 4 categories, alpha x pinv x 9 branch lengths ---------
ODD_BRANCHES = (0.1, 0.2, 0.5, 0.9, 1.5, 5, 10, 50, 90)
for cats in (1, 2, 4):
    for alpha in (0.1, 0.75, 1.5):
        for pinv in (0.0, 0.3, 0.9):
            pr = [0] * cats
            steps = ([dict(do="pinv", index=0, value=pinv)] if pinv > 0 else []) + [
                dict(do="pmatrix", params=pr, matrices=[0, 1, 2, 3], lengths=[0.1, 0.2, 0.3, 0.4]),
                dict(do="partials", ops=[op(5, NONE, 0, 1, NONE, 1, 1, NONE), op(6, NONE, 5, 0, NONE, 2, 1, NONE),
                                         op(7, NONE, 3, 1, NONE, 4, 1, NONE)]),
                dict(do="edge", args=[6, NONE, 7, NONE, 0], freqs_indices=pr),
                dict(do="sumtable", key="ii", edge=[6, 7, NONE, NONE], params=pr)]
            for t in ODD_BRANCHES:
                steps += [dict(do="derivs", key="ii", t=t, params=pr),
                          dict(do="pmatrix", params=pr, matrices=[0], lengths=[t]),
                          dict(do="edge", args=[6, NONE, 7, NONE, 0], freqs_indices=pr, tag=f"Branch {t}")]
            steps += [dict(do="pmatrix", params=pr, matrices=[0], lengths=[0.1]),
                      dict(do="partials", ops=[op(7, NONE, 6, 0, NONE, 3, 0, NONE)]),
                      dict(do="edge", args=[4, NONE, 7, NONE, 1], freqs_indices=pr),
                      dict(do="sumtable", key="ti", edge=[4, 7, NONE, NONE], params=pr)]
            for t in ODD_BRANCHES:
                steps += [dict(do="derivs", key="ti", t=t, params=pr),
                          dict(do="pmatrix", params=pr, matrices=[1], lengths=[t]),
                          dict(do="edge", args=[4, NONE, 7, NONE, 1], freqs_indices=pr, tag=f"Branch(Tip) {t}")]
            CASES.append(dict(
                name=f"derivatives_oddstates_c{cats}_a{alpha}_p{pinv}", states=5, tips=5, clv_buffers=4, sites=20,
                rate_matrices=1, prob_matrices=7, rate_cats=cats, scale_buffers=0, alpha=alpha, map=odd_map(5),
                freqs=[[0.3, 0.25, 0.1, 0.2, 0.15]],
                subst=[[1.452176, 0.937951, 0.462880, 0.617729, 1.745312, 0.937951, 0.462880, 0.617729, 1.745312, 1.0]],
                seqs=["DAACBCECBA--ABBCBAAB", "CACCABECBA--ABBEBCBB", "AE-C-BECAE--CBBCBACB",
                      "CEBCBBECAA--AB-C-AAE", "CEACBBECCA--AB-B-AAE"],
                steps=steps, odd_block=dict(alpha=alpha, cats=cats, pinv=pinv)))

# ---- ascertainment-bias correction: recipe of reference test/src/asc-bias.c:206-265 (no
# correction, Lewis, Felsenstein, Stamatakis; lnL and derivatives) on the 00010 data - the test's
# own alignment (testdata/2000.fas) is not shipped with the reference ---------------------------
AB_LEWIS, AB_FELSENSTEIN, AB_STAMATAKIS, AB_FLAG = 1 << 5, 2 << 5, 3 << 5, 1 << 8
asc_steps = [
    dict(do="pmatrix", params=[0, 0, 0, 0], matrices=[0, 1, 2, 3], lengths=[0.1, 0.2, 1, 1]),
    dict(do="partials", ops=[op(5, NONE, 0, 1, NONE, 1, 1, NONE), op(6, NONE, 5, 0, NONE, 2, 1, NONE),
                             op(7, NONE, 3, 1, NONE, 4, 1, NONE)]),
]
for asc_type in (0, AB_LEWIS, AB_FELSENSTEIN, AB_STAMATAKIS):
    asc_steps.append(dict(do="asc_type", value=asc_type))
    if asc_type in (AB_FELSENSTEIN, AB_STAMATAKIS):
        asc_steps.append(dict(do="asc_weights", value=[5, 7, 3, 9]))
    asc_steps.append(dict(do="edge", args=[6, NONE, 7, NONE, 0], freqs_indices=[0, 0, 0, 0], tag=f"asc {asc_type} ii"))
    asc_steps.append(dict(do="edge", args=[6, NONE, 2, NONE, 1], freqs_indices=[0, 0, 0, 0], tag=f"asc {asc_type} ti"))
    asc_steps.append(dict(do="root", clv=6, scaler=NONE, freqs_indices=[0, 0, 0, 0], tag=f"asc {asc_type} root"))
    asc_steps.append(dict(do="sumtable", key=f"ii{asc_type}", edge=[6, 7, NONE, NONE], params=[0, 0, 0, 0]))
    asc_steps.append(dict(do="sumtable", key=f"ti{asc_type}", edge=[6, 2, NONE, NONE], params=[0, 0, 0, 0]))
    for t in (0.0001, 0.01, 0.1, 1.0, 10.0):
        asc_steps.append(dict(do="derivs", key=f"ii{asc_type}", t=t, params=[0, 0, 0, 0]))
        asc_steps.append(dict(do="derivs", key=f"ti{asc_type}", t=t, params=[0, 0, 0, 0]))
CASES.append(dict(
    name="asc_bias_recipe", states=4, tips=5, clv_buffers=4, sites=12, rate_matrices=1, prob_matrices=7,
    rate_cats=4, scale_buffers=0, alpha=0.5, extra_attributes=AB_FLAG, no_port=True,
    freqs=[[0.3, 0.4, 0.1, 0.2]], subst=[[1, 2.5, 1, 1, 2.5, 1]],
    seqs=["WAC-CTA-ATCT", "CCC-TTA-ATGT", "A-C-TAG-CTCT", "CTCTTAA-A-CG", "CAC-TCA-A-TG"],
    steps=asc_steps))

# test 00011 shares the step list of 00010
CASES[1]["steps"] = [dict(s) for s in CASES[0]["steps"]]
CASES[1]["printed"] = {"inner-inner": -227.371279, "tip-inner": -227.371279}


# ------------------------------------------------------------------------------------------
def run_case(lib, case, attributes):
    """Executes a case through the pll.h API of `lib`; returns the list of outputs, one entry
    per step that produces something.  (tests/golden_runner.py holds the same interpreter for
    the libraries under test.)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_runner import execute

    return execute(lib, case, attributes)


def parse_printed(path, tags):
    text = open(os.path.join(REF, path)).read()
    out = {}
    for tag in tags:
        m = re.search(re.escape(tag) + r"\s+logL:\s+(-?\d+\.\d+)", text)
        if m:
            out[tag] = float(m.group(1))
    return out


def parse_odd_block(alpha, cats, pinv):
    """(lnL, d_f, dd_f) rows printed by the reference for one block of derivatives-oddstates.out"""
    text = open(os.path.join(REF, "test/out/derivatives-oddstates.out")).read()
    head = " TEST alpha(ncats) = %6.2f(%2d) ; pinv = %.2f" % (alpha, cats, pinv)
    block = text[text.index(head):].split(" TEST alpha")[1]
    rows = re.findall(r"Branch(?:\(Tip\))?\s+[\d.]+ :\s+(-?[\d.]+)\s+(-?[\d.e+-]+)\s+(-?[\d.e+-]+)", block)
    assert len(rows) == 18, (head, len(rows))
    return [tuple(float(x) for x in r) for r in rows]


def check_odd_block(case, outs):
    want = parse_odd_block(**case["odd_block"])
    d = [o for o in outs if o["kind"] == "derivs"]
    e = [o for o in outs if o["kind"] == "edge" and o.get("tag")]
    assert len(d) == 18 and len(e) == 18
    for (lnl, d1, d2), od, oe in zip(want, d, e):
        assert abs(oe["logl"] - lnl) < 5e-7, (case["name"], oe["logl"], lnl)
        assert abs(od["d_f"] - d1) <= 6e-5 * abs(d1) + 1e-18, (case["name"], od, d1)
        assert abs(od["dd_f"] - d2) <= 6e-5 * abs(d2) + 1e-40, (case["name"], od, d2)


def main():
    golden = []
    for case in CASES:
        entry = dict(case)
        entry["expect"] = {}
        entry["rates"] = ref.gamma_rates(case["alpha"], case["rate_cats"]).tolist()
        for label, attr in (("tv", PLL_ATTRIB_ARCH_AVX2 | PLL_ATTRIB_PATTERN_TIP), ("notv", PLL_ATTRIB_ARCH_AVX2)):
            entry["expect"][label] = run_case(ref, case, attr)
        # pin to the reference's text fixtures
        printed = dict(case.get("printed", {}))
        if case.get("out_file"):
            from_file = parse_printed(case["out_file"], list(printed))
            for tag, v in from_file.items():
                assert abs(v - printed[tag]) < 1e-9, (case["name"], tag, v, printed[tag])
            assert from_file, f"no lnL lines found in {case['out_file']}"
        for label in ("tv", "notv"):
            for out in entry["expect"][label]:
                if out.get("tag") in printed:
                    assert abs(out["logl"] - printed[out["tag"]]) < 5e-7, (case["name"], out["tag"], out["logl"])
                if out.get("kind") == "newton":
                    step = [s for s in case["steps"] if s["do"] == "newton"][0]
                    assert abs(out["final"] - step["printed_final"]) < 5e-7, out
                    assert out["iterations"] == step["printed_iterations"], out
            if "odd_block" in case:
                check_odd_block(case, entry["expect"][label])
        golden.append(entry)
        print("ok", case["name"])
    path = os.path.join(ROOT, "tests", "golden", "cases.json")
    json.dump(golden, open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
