"""Site-pattern sharding across GPUs.

Every kernel of the likelihood path treats alignment patterns independently; only the final
sums (lnL, d_f, dd_f) couple them.  One process per GPU therefore owns a contiguous slice of
the patterns of EVERY CLV / scale buffer / tip row / weight array, replicates the KB-sized model
state, and the only exchange is a scalar all-reduce (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  This module holds that host-side logic; it is backend-agnostic (any object with
the `binding.Partition` surface works), which is what the world_size-2 gloo tests exercise.
"""
from __future__ import annotations

from typing import Optional, Tuple

ALIGN = 64  # patterns; keeps every rank's slice 256-byte aligned in all per-site arrays


def slice_bounds(sites: int, world: int, rank: int, align: int = ALIGN) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of rank `rank`: boundaries at multiples of `align`, sizes
    differing by at most one block, all patterns covered exactly once."""
    assert 0 <= rank < world and sites > 0
    blocks = (sites + align - 1) // align
    base, extra = divmod(blocks, world)
    lo_b = rank * base + min(rank, extra)
    hi_b = lo_b + base + (1 if rank < extra else 0)
    return min(lo_b * align, sites), min(hi_b * align, sites)


def allreduce_sum(values, group=None):
    """Sum a short list of Python floats over the ranks (double precision).  Without an
    initialised process group this is the identity (single GPU)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return list(values)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(list(values), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(x) for x in t.tolist()]


def sharded_edge_loglikelihood(part, args, freqs_indices, group=None) -> float:
    """lnL of the whole alignment from per-rank partitions (each built over slice_bounds)."""
    local = part.edge_loglikelihood(*args, freqs_indices)
    return allreduce_sum([local], group)[0]


def sharded_derivatives(part, parent_scaler, child_scaler, branch_length, params_indices, sumtable,
                        group=None) -> Tuple[float, float]:
    d1, d2 = part.likelihood_derivatives(parent_scaler, child_scaler, branch_length, params_indices, sumtable)
    s = allreduce_sum([d1, d2], group)
    return s[0], s[1]
