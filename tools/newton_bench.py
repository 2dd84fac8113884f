#!/usr/bin/env python3
"""Developer tool: C4 (Newton branch-length optimisation) timings - sumtable + derivative passes."""
import argparse, sys, time
sys.path.insert(0, ".")
import numpy as np
import libpll_b200
from libpll_b200 import synthetic as S
from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP
ap = argparse.ArgumentParser()
ap.add_argument("--tips", type=int, default=100)
ap.add_argument("--sites", type=int, default=1000000)
ap.add_argument("--states", type=int, default=4)
a = ap.parse_args()
lib = libpll_b200.load()
w = S.make_workload(a.tips, a.sites, states=a.states)
seqs = [S.tip_sequence(w, t) for t in range(8)]
S.tip_sequence = lambda w_, t, lo=0, hi=None: seqs[t % 8]
part, pidx = S.build_partition(lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
S.full_evaluation(part, w, pidx)
span = w.rate_cats * w.states * 8
edges = {"ii": (w.root_a, w.root_b)}
last = w.ops[-1]
for c in ("child1_clv_index", "child2_clv_index"):
    if int(last[c]) < w.tips:
        edges["ti"] = (int(last["parent_clv_index"]), int(last[c]))
for name, (pa, ch) in edges.items():
    tab = part.new_sumtable()
    sa, sb = w.scaler_of(pa), w.scaler_of(ch)
    for _ in range(3): part.update_sumtable(pa, ch, sa, sb, pidx, tab)
    part.timer_start()
    for _ in range(10): part.update_sumtable(pa, ch, sa, sb, pidx, tab)
    ms = part.timer_stop() / 10
    nbytes = (3 if name == "ii" else 2) * span * a.sites
    print(f"sumtable {name}: {ms*1e3:.1f} us  {nbytes/ms/1e6:.0f} GB/s")
    for _ in range(3): part.likelihood_derivatives(sa, sb, 0.1, pidx, tab)
    t0 = time.perf_counter()
    n = 50
    for i in range(n): d = part.likelihood_derivatives(sa, sb, 0.1 + 0.001 * i, pidx, tab)
    wall = (time.perf_counter() - t0) / n
    print(f"derivative pass {name}: {wall*1e6:.1f} us wall per call  ({(span+4)*a.sites/wall/1e9:.0f} GB/s incl. sync)  d={d}")
    # Newton loop like reference examples/newton/newton.c
    t0 = time.perf_counter(); length = 0.1; its = 0
    for _ in range(32):
        d1, d2 = part.likelihood_derivatives(sa, sb, length, pidx, tab); its += 1
        if abs(d1) < 1e-5: break
        length -= d1 / d2
    print(f"newton {name}: {its} iterations to t={length:.6f} in {(time.perf_counter()-t0)*1e3:.3f} ms")
part.destroy()
