#!/usr/bin/env python3
"""Developer timing loop (not the judged benchmark - see bench.py): times the batched
traversal and the edge lnL of a synthetic workload with CUDA events on the library's stream."""
import argparse
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import libpll_b200
from libpll_b200 import synthetic as S
from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP

ap = argparse.ArgumentParser()
ap.add_argument("--tips", type=int, default=200)
ap.add_argument("--sites", type=int, default=200000)
ap.add_argument("--states", type=int, default=4)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--fast-tips", action="store_true", help="reuse 8 distinct tip sequences")
a = ap.parse_args()

import os
if os.environ.get("PLL_B200_LIB"):  # developer experiments: another build of the library
    libpll_b200.LIB_PATH = os.path.abspath(os.environ["PLL_B200_LIB"])
lib = libpll_b200.load()
w = S.make_workload(a.tips, a.sites, states=a.states)
t0 = time.time()
if a.fast_tips:
    seqs = [S.tip_sequence(w, t) for t in range(8)]
    orig = S.tip_sequence
    S.tip_sequence = lambda w_, t, lo=0, hi=None: seqs[t % 8]
part, pidx = S.build_partition(lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
print(f"setup {time.time()-t0:.1f}s  ops tt/ti/ii={w.op_kinds()}  bytes/site={w.algorithmic_bytes_per_site()}",
      flush=True)
part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
res = {}
for it in range(a.iters):
    part.reset_stats()
    part.timer_start()
    part.update_partials(w.ops)
    ms = part.timer_stop()
    st = part.stats()
    part.timer_start()
    lnl = part.edge_loglikelihood(w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b),
                                  w.root_matrix, pidx)
    ms_l = part.timer_stop()
    gbs = st["algorithmic_bytes"] / ms / 1e6
    print(f"iter {it}: traversal {ms:.3f} ms  {gbs:.0f} GB/s algorithmic  "
          f"{len(w.ops)*a.sites/(ms*1e-3):.3e} site-updates/s  kernels={st['kernel_launches']} "
          f"levels={st['partial_levels']}  edge lnL {ms_l:.3f} ms  lnL={lnl:.6f}", flush=True)
part.destroy()
