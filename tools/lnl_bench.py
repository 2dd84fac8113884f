#!/usr/bin/env python3
"""Developer tool: timings of the value-returning calls (edge / root lnL) on a resident partition."""
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
last = w.ops[-1]
cases = {"edge ii": (w.root_a, w.root_b, w.root_matrix, 2 * span + 12)}
for c in ("child1", "child2"):
    if int(last[c + "_clv_index"]) < w.tips:
        cases["edge ti"] = (int(last["parent_clv_index"]), int(last[c + "_clv_index"]), int(last[c + "_matrix_index"]), span + 9)
for name, (pa, ch, m, bytes_per_site) in cases.items():
    args = (pa, w.scaler_of(pa), ch, w.scaler_of(ch), m, pidx)
    for _ in range(3): part.edge_loglikelihood(*args)
    t0 = time.perf_counter(); n = 50
    for _ in range(n): v = part.edge_loglikelihood(*args)
    wall = (time.perf_counter() - t0) / n
    print(f"{name}: {wall*1e6:.1f} us wall per call ({bytes_per_site*a.sites/wall/1e9:.0f} GB/s incl. sync) lnl={v}")
top = w.tips + w.inner - 1
for _ in range(3): part.root_loglikelihood(top, w.scaler_of(top), pidx)
t0 = time.perf_counter()
for _ in range(50): v = part.root_loglikelihood(top, w.scaler_of(top), pidx)
wall = (time.perf_counter() - t0) / 50
print(f"root: {wall*1e6:.1f} us wall per call ({(span+8)*a.sites/wall/1e9:.0f} GB/s incl. sync) lnl={v}")
part.destroy()
