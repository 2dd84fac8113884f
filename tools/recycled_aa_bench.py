#!/usr/bin/env python3
"""Developer tool: 20-state traversal of a list that recycles CLV / scaler slots (the memory-saving
mode of large analyses): level-by-level kernels (every CLV is written) against the whole-list walk
(only the last value of a buffer is stored)."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import libpll_b200
from libpll_b200 import synthetic as S
from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP

tips, sites, slots = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
lib = libpll_b200.load()
w = S.recycle_slots(S.make_workload(tips, sites, states=20), slots)
for walk in ("0", "1"):
    os.environ["PLL_GPU_FUSED_AA"] = walk
    part, pidx = S.build_partition(lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
    part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    root = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
    for _ in range(3):
        part.update_partials(w.ops); lnl = part.edge_loglikelihood(*root)
    part.reset_stats(); part.timer_start()
    for _ in range(5):
        part.update_partials(w.ops)
    ms = part.timer_stop() / 5
    st = part.stats()
    print(f"PLL_GPU_FUSED_AA={walk}: {tips} taxa x {sites} patterns, {slots} recycled slots: {ms:.2f} ms per traversal, "
          f"{len(w.ops) * sites / ms / 1e3:.3e} site-updates/s, kernels per traversal {st['kernel_launches'] // 5}, "
          f"compulsory GB {st['compulsory_bytes'] / 5 / 1e9:.1f}, lnL {lnl:.6f}", flush=True)
    part.destroy()
