#!/usr/bin/env python3
"""One process, one pll_partition_t, N GPUs (pll_gpu_set_devices; not the judged benchmark, which
runs one rank per GPU under torchrun - see bench.py).  Times the C2 step through the plain pll.h
calls: all P-matrices, the full traversal and the edge log-likelihood, wall clock around K steps
(the log-likelihood call returns only when every device has delivered its partial sum).

    python tools/inprocess_devices_bench.py --devices 2 --sites-per-device 1000000
"""
import argparse
import json
import sys
import time

sys.path.insert(0, ".")
import libpll_b200
from libpll_b200 import synthetic as S
from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP

ap = argparse.ArgumentParser()
ap.add_argument("--devices", type=int, default=2)
ap.add_argument("--tips", type=int, default=1000)
ap.add_argument("--sites-per-device", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--distinct-tips", type=int, default=16,
                help="tip sequences generated on the host and reused round-robin (set-up time only)")
a = ap.parse_args()

lib = libpll_b200.load()
visible = lib.pll_gpu_device_count()
sites = a.sites_per_device * a.devices
w = S.make_workload(a.tips, sites, states=4)
seqs = [S.tip_sequence(w, t) for t in range(a.distinct_tips)]
S.tip_sequence = lambda w_, t, lo=0, hi=None: seqs[t % len(seqs)]
out = {"visible_devices": visible, "tips": a.tips, "sites": sites, "ops": len(w.ops)}
for n in sorted({1, a.devices}):
    if n == 1 and a.devices > 1:
        # same per-device load on one device for comparison: the first slice only
        w1 = S.make_workload(a.tips, a.sites_per_device, states=4)
        seq1 = [s[:a.sites_per_device] for s in seqs]
        S.tip_sequence = lambda w_, t, lo=0, hi=None: seq1[t % len(seq1)]
        wl = w1
    else:
        S.tip_sequence = lambda w_, t, lo=0, hi=None: seqs[t % len(seqs)]
        wl = w
    assert lib.pll_gpu_set_devices(n) == 1
    t0 = time.time()
    part, pidx = S.build_partition(lib, wl, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
    lib.pll_gpu_set_devices(0)
    used = lib.pll_gpu_partition_devices(part.ptr)
    print(f"[{n} slice(s)] set-up {time.time() - t0:.1f}s, {wl.sites} patterns on {used} context(s)", flush=True)
    root = (wl.root_a, wl.scaler_of(wl.root_a), wl.root_b, wl.scaler_of(wl.root_b), wl.root_matrix, pidx)

    def step(i):
        bl = wl.branch_lengths * (1.0 + 1e-3 * ((i % 7) - 3))
        part.update_prob_matrices(pidx, wl.matrix_indices, bl)
        part.update_partials(wl.ops)
        return part.edge_loglikelihood(*root)

    for i in range(a.warmup):
        lnl = step(i)
    t0 = time.perf_counter()
    for i in range(a.steps):
        lnl = step(i)
    dt = (time.perf_counter() - t0) / a.steps
    rate = len(wl.ops) * wl.sites / dt
    # Newton-style calls on the evaluation edge
    tab = part.new_sumtable()
    part.update_sumtable(root[0], root[2], root[1], root[3], pidx, tab)
    part.likelihood_derivatives(root[1], root[3], 0.1, pidx, tab)
    t0 = time.perf_counter()
    for i in range(32):
        part.likelihood_derivatives(root[1], root[3], 0.1 + 0.01 * i, pidx, tab)
    dt_d = (time.perf_counter() - t0) / 32
    out[f"slices_{n}"] = {"contexts": used, "patterns": wl.sites, "ms_per_step": dt * 1e3,
                          "site_updates_per_s": rate, "lnl": lnl, "derivative_call_us": dt_d * 1e6}
    print(f"[{n} slice(s)] {dt * 1e3:.2f} ms/step  {rate:.3e} site-updates/s  lnL={lnl:.6f}  "
          f"derivative call {dt_d * 1e6:.0f} us", flush=True)
    part.destroy()
print(json.dumps(out))
