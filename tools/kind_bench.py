#!/usr/bin/env python3
"""Developer tool: per-kind (tt/ti/ii) event-timed bandwidth of the CLV traversal."""
import argparse, sys
sys.path.insert(0, ".")
import libpll_b200
from libpll_b200 import synthetic as S
from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP
ap = argparse.ArgumentParser()
ap.add_argument("--tips", type=int, default=300)
ap.add_argument("--sites", type=int, default=1000000)
ap.add_argument("--states", type=int, default=4)
a = ap.parse_args()
lib = libpll_b200.load()
w = S.make_workload(a.tips, a.sites, states=a.states)
seqs = [S.tip_sequence(w, t) for t in range(8)]
S.tip_sequence = lambda w_, t, lo=0, hi=None: seqs[t % 8]
part, pidx = S.build_partition(lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
for _ in range(2): part.update_partials(w.ops)
part.timer_start()
for _ in range(3): part.update_partials(w.ops)
ms = part.timer_stop() / 3
part.set_profiling(True); part.reset_stats()
for _ in range(3): part.update_partials(w.ops)
st = part.stats()
print(f"traversal {ms:.3f} ms  {w.algorithmic_bytes_per_site()*a.sites/ms/1e6:.0f} GB/s algorithmic")
for i, n in enumerate(["tt", "ti", "ii"]):
    if st["kind_launches"][i]:
        print(f"  {n}: {st['kind_bytes'][i]/st['kind_ns'][i]:.0f} GB/s  share {st['kind_ns'][i]/sum(st['kind_ns']):.3f}  launches {st['kind_launches'][i]//3}")
part.destroy()
