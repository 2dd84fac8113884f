#!/usr/bin/env python3
"""Developer tool: where does BASELINE configs[4] lose strong-scaling efficiency?  Runs the C5
operations list (5,000 taxa, 64 recycled slots, device-generated tips) on ONE GPU for the slice
sizes a rank owns at N = 1, 2, 4, 8 GPUs.  If slice time x N stays flat, the kernel scales and what
an N-GPU run loses is outside it (clocks under a shared power budget, the scalar all-reduce,
the slowest rank); if it grows, the loss is per launch (fixed costs, tail of the tile walk)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import libpll_b200
from libpll_b200 import synthetic as S
from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP

import os
if os.environ.get("PLL_B200_LIB"):  # developer experiments: another build of the library
    libpll_b200.LIB_PATH = os.path.abspath(os.environ["PLL_B200_LIB"])
lib = libpll_b200.load()
tips, total, slots = 5000, 10_000_000, 64
w = S.recycle_slots(S.make_workload(tips, 64, states=4), slots)
base = None
for n in ((8,) if os.environ.get("PLL_B200_LIB") else (8, 4, 2, 1)):
    per = (total // n + 63) // 64 * 64
    part = lib.partition(tips=tips, clv_buffers=w.inner, states=4, sites=per, rate_matrices=1,
                         prob_matrices=w.prob_matrices, rate_cats=4, scale_buffers=w.inner,
                         attributes=PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
    part.set_frequencies(0, S.GTR_FREQS)
    part.set_subst_params(0, S.GTR_RATES)
    part.set_category_rates(lib.gamma_rates(w.alpha, 4))
    part.set_category_weights(np.full(4, 0.25))
    for t in range(tips):
        assert lib.pll_gpu_generate_tip_states(part.ptr, t, 43, 0) == 1, lib.errmsg()
    pidx = np.zeros(4, np.uint32)
    part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    root = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
    for _ in range(2):
        part.update_partials(w.ops); part.edge_loglikelihood(*root)
    steps = 3
    part.timer_start()
    for _ in range(steps):
        part.update_partials(w.ops); part.edge_loglikelihood(*root)
    ms = part.timer_stop() / steps
    t0 = time.perf_counter()
    for _ in range(steps):
        part.update_partials(w.ops); part.edge_loglikelihood(*root)
    wall = (time.perf_counter() - t0) * 1e3 / steps
    part.destroy()
    print(f"slice of N={n}: {per} patterns  device {ms:8.2f} ms  wall {wall:8.2f} ms  x N = {ms * n:8.1f} ms  "
          f"{len(w.ops) * per / (ms * 1e-3):.3e} site-updates/s per GPU", flush=True)
