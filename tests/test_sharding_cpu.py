"""Multi-GPU host logic on CPU: world_size-2 gloo run of the site-sharded evaluation
(libpll_b200/sharding.py).  The per-rank compute is done by the reference library (no GPU
here); what is under test is the slicing, the scalar all-reduce and that the sharded lnL and
derivatives equal the unsharded ones."""
import os
import socket
import sys

import numpy as np
import pytest

from libpll_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slice_bounds_cover_everything_once():
    for sites in (1, 63, 64, 65, 1000, 4099, 1_000_000, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            prev_hi = 0
            for r in range(world):
                lo, hi = sharding.slice_bounds(sites, world, r)
                assert lo == prev_hi and lo <= hi
                assert lo % sharding.ALIGN == 0 or lo == sites
                covered += hi - lo
                prev_hi = hi
            assert covered == sites and prev_hi == sites


def _worker(rank, world, port, sites, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from libpll_b200 import synthetic as S
    from libpll_b200.binding import PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_PATTERN_TIP, PllLibrary

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ref = PllLibrary(os.path.join(ROOT, "oracle", "_ref", "libpll_ref.so"), is_gpu=False)
    w = S.make_workload(16, sites, states=4, seed=17)
    lo, hi = sharding.slice_bounds(sites, world, rank)
    part, pidx = S.build_partition(ref, w, PLL_ATTRIB_ARCH_AVX2 | PLL_ATTRIB_PATTERN_TIP, lo=lo, hi=hi)
    part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    part.update_partials(w.ops)
    edge = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix)
    lnl = sharding.sharded_edge_loglikelihood(part, edge, pidx)
    table = part.new_sumtable()
    part.update_sumtable(w.root_a, w.root_b, w.scaler_of(w.root_a), w.scaler_of(w.root_b), pidx, table)
    d1, d2 = sharding.sharded_derivatives(part, w.scaler_of(w.root_a), w.scaler_of(w.root_b), 0.13, pidx, table)
    if rank == 0:
        out.put((lnl, d1, d2))
    part.destroy()
    dist.destroy_process_group()


def test_sharded_evaluation_world_size_2(ref_lib):
    import torch.multiprocessing as mp

    from libpll_b200 import synthetic as S
    from libpll_b200.binding import PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_PATTERN_TIP

    sites, world = 1000, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, sites, out)) for r in range(world)]
    [p.start() for p in procs]
    lnl, d1, d2 = out.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)

    w = S.make_workload(16, sites, states=4, seed=17)
    part, pidx = S.build_partition(ref_lib, w, PLL_ATTRIB_ARCH_AVX2 | PLL_ATTRIB_PATTERN_TIP)
    full = S.full_evaluation(part, w, pidx)
    table = part.new_sumtable()
    part.update_sumtable(w.root_a, w.root_b, w.scaler_of(w.root_a), w.scaler_of(w.root_b), pidx, table)
    f1, f2 = part.likelihood_derivatives(w.scaler_of(w.root_a), w.scaler_of(w.root_b), 0.13, pidx, table)
    part.destroy()
    assert abs(lnl - full) <= 1e-12 * abs(full)
    assert abs(d1 - f1) <= 1e-10 * max(abs(f1), 1.0) and abs(d2 - f2) <= 1e-10 * max(abs(f2), 1.0)
