#!/usr/bin/env python3
"""Headline benchmark: full post-order CLV traversal + edge log-likelihood of a synthetic
1,000-taxon x 1M-pattern GTR+G4 DNA partition per B200 (BASELINE.json configs[1]), with the other
BASELINE configurations as sub-records of the same JSON line.

    python bench.py --gpus N --steps K --warmup W          # this repository's CUDA path
    python bench.py --impl reference ...                   # the reference's own AVX2 CPU path

One "step" = one full-tree evaluation: the whole traversal (`pll_update_partials`
over T-2 operations) and `pll_compute_edge_loglikelihood`.  With N > 1 (torchrun, one rank per GPU)
the site patterns are sharded: every rank owns a contiguous slice of 1M patterns of EVERY CLV (weak
scaling: the alignment has N x 1M patterns) and only the per-rank partial lnL crosses NVLink (one
scalar NCCL all-reduce per step).

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  value        CLV site-updates/s, device-timed (CUDA events on the library's own stream) with
               everything resident in HBM: traversal + edge lnL, P-matrices already on device
  e2e          the same metric through the public pll.h API from HOST buffers, wall-clocked:
               each step uploads fresh branch lengths / matrix indices / the operations array
               (pinned staging -> HBM), recomputes all P-matrices, traverses, reads lnL back
  roofline     dominant kernel (k_traverse_dna, the whole list in one launch): COMPULSORY DRAM
               bytes of its plan (every observable parent CLV / scaler written once, tip characters
               and tile-cache misses read; plg_stats.compulsory_bytes) / CUDA-event time per launch
               against MEASURED_PEAKS.json; SURVEY 8d's algorithmic figure under `algorithmic`
  cpu_baseline the reference (oracle/_ref) timed on this box's host cores on the same workload
  lnl_check    the GPU lnL against the sum of the CPU arm's slice lnLs (rel <= 1e-10 or the run fails)
  also         the other BASELINE configurations, each bounded to a few seconds:
                 lists  e2e with a DIFFERENT partial traversal every step (moving virtual root)
                 c4     Newton branch-length optimisation: sumtable + 32 derivative passes on the
                        evaluation edge and 9 re-rooted edges (configs[3])
                 c3     500-taxon x 200k-pattern protein LG+G4 traversal + edge lnL (configs[2])
                 c5     5,000-taxon x 10M-pattern DNA sharded over the N GPUs, strong scaling, 64
                        recycled CLV slots, tips generated on the device (configs[4])
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

print_json = print

METRIC = "CLV site-updates/sec (full post-order traversal + edge logL, GTR+G4 DNA)"
UNIT = "site-updates/s"
DMMA_PEAK_TFLOPS = 37.1  # measured on this pool's B200 (tools/ubench/dmma_peak.cu, profiles/)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if "hbm_gbs" in peaks:
            return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        pass
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def measured_traffic(key: str, **must_match):
    """DRAM bytes per launch from an ncu capture of the SAME command (profiles/r02_traffic.json);
    None unless the capture matches this run's configuration."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(key)
    except Exception:
        return None, None
    if not t or any(t.get(k) != v for k, v in must_match.items()):
        return None, None
    return t["dram_bytes_per_launch"], t.get("source")


# ----------------------------------------------------------------------------------------
# tips are generated once per process and shared by every partition built from them
# ----------------------------------------------------------------------------------------
def memoize_tips():
    from libpll_b200 import synthetic as S

    cache = {}
    orig = S.tip_sequence

    def memo(w, tip, lo=0, hi=None):
        hi = w.sites if hi is None else hi
        if w.sites > 2_000_000:          # a rank of a multi-GPU run only ever asks for its own slice
            return orig(w, tip, lo, hi)
        key = (w.tips, w.sites, w.states, w.seed, tip)
        if key not in cache:
            cache[key] = orig(w, tip, 0, w.sites)
        return cache[key][lo:hi]

    S.tip_sequence = memo
    return cache


# ----------------------------------------------------------------------------------------
# CPU baseline = the reference's own AVX2 path on host cores
# ----------------------------------------------------------------------------------------
def _ref_library():
    from libpll_b200.binding import PllLibrary

    path = os.path.join(ROOT, "oracle", "_ref", "libpll_ref.so")
    if os.path.exists(path):
        return PllLibrary(path, is_gpu=False), "reference"
    return None, "port"


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_traversal(w_full, slices, reps: int, warm: int, use_mean: bool = False):
    """Downstream convention for the single-threaded reference: one host thread
    per entry of `slices` = [(lo, hi)], each owning an independent partition over that contiguous
    pattern slice of the SAME workload; time = slowest thread.  Returns (site-updates/s, seconds
    per step, [lnL per slice], kind)."""
    from libpll_b200 import synthetic as S
    from libpll_b200.binding import PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_PATTERN_TIP

    ref, kind = _ref_library()
    threads = len(slices)
    total_sites = sum(hi - lo for lo, hi in slices)
    if ref is None:
        from oracle import port as oracle_port

        spt = slices[0][1] - slices[0][0]
        rate, t, lnl = oracle_port.cpu_traversal_rate(w_full, threads, spt, reps)
        return rate, t, [lnl], kind

    parts = [None] * threads
    pidx = [None]

    def setup(t):
        lo, hi = slices[t]
        parts[t], pidx[0] = S.build_partition(ref, w_full, PLL_ATTRIB_ARCH_AVX2 | PLL_ATTRIB_PATTERN_TIP,
                                              lo=lo, hi=hi)

    ths = [threading.Thread(target=setup, args=(t,)) for t in range(threads)]
    [t.start() for t in ths]
    [t.join() for t in ths]

    times = np.zeros((threads, warm + reps))
    lnl = np.zeros(threads)
    barrier = threading.Barrier(threads)

    def work(t):
        for r in range(warm + reps):
            barrier.wait()
            t0 = time.perf_counter()
            lnl[t] = S.full_evaluation(parts[t], w_full, pidx[0])  # ctypes releases the GIL
            times[t, r] = time.perf_counter() - t0

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for p in parts:
        p.destroy()
    step = times[:, warm:].max(axis=0)  # slowest thread per repetition
    t = float(step.mean()) if use_mean else float(step.min())
    return len(w_full.ops) * total_sites / t, t, [float(x) for x in lnl], kind


def even_slices(sites: int, parts: int):
    """`parts` contiguous pattern slices, 64-pattern aligned (the device slicing rule)."""
    per = ((sites + parts - 1) // parts + 63) // 64 * 64
    out = []
    lo = 0
    while lo < sites:
        out.append((lo, min(lo + per, sites)))
        lo += per
    return out


CPU_SLOTS = 64  # the CPU arm keeps its CLVs in a pool of recycled slots so that host RAM holds 1M patterns


def full_config_cpu(w, cores: int, reps: int, warm: int, use_mean: bool):
    """The reference on the WHOLE workload: `cores` threads x (sites / cores) patterns.  Same tree,
    tips, model and operations; the inner CLVs live in 64 recycled slots per thread (legal pll.h
    use; the plain list would need 128 B x 998 x 1M = 128 GB of host memory)."""
    from libpll_b200 import synthetic as S

    wr = S.recycle_slots(w, CPU_SLOTS)
    slices = even_slices(w.sites, cores)
    rate, t, lnls, kind = cpu_traversal(wr, slices, reps=reps, warm=warm, use_mean=use_mean)
    return rate, t, lnls, kind, slices


# ----------------------------------------------------------------------------------------
def workload_string(tips: int, sites: int, per_gpu: int) -> str:
    return (f"synthetic {tips}-taxon x {sites}-pattern GTR+G4 DNA ({per_gpu} patterns per GPU), "
            f"full post-order traversal + edge logL")


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's own CPU implementation of the path on host cores, on the
    full configuration of the GPU arm (all patterns, all operations)."""
    if rank != 0:
        return
    from libpll_b200 import synthetic as S

    memoize_tips()
    cores = host_cores()
    total = args.sites_per_gpu * max(world, 1)
    w = S.make_workload(args.tips, total, states=4)
    t0 = time.time()
    rate, step_s, lnls, kind, slices = full_config_cpu(w, cores, reps=args.steps, warm=args.warmup, use_mean=True)
    log(f"[bench] reference arm: {cores} threads, {step_s:.2f} s per step, total {time.time() - t0:.0f} s")
    # one-thread row: a 10,000-pattern slice on a single core
    one_rate, one_t, _, _ = cpu_traversal(S.recycle_slots(w, CPU_SLOTS), [(0, min(10_000, total))], reps=1, warm=0)
    sample = (f"ALL {total} patterns: {cores} threads x {slices[0][1] - slices[0][0]} patterns each, "
              f"{len(w.ops)} operations per traversal, inner CLVs in {CPU_SLOTS} recycled slots per thread, "
              f"mean of {args.steps} steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_string(args.tips, total, args.sites_per_gpu),
                   "attributes": "PLL_ATTRIB_ARCH_AVX2|PLL_ATTRIB_PATTERN_TIP, per-site scalers",
                   "operations": len(w.ops), "rate_cats": 4,
                   "timed": "the whole configuration (no extrapolation), see cpu_baseline.sample"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "cpu_model": cpu_model(),
                         "one_thread": {"value": one_rate, "unit": UNIT, "cores": 1,
                                        "sample": f"one thread x {min(10_000, total)} patterns, {one_t:.2f} s"}},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lnl": float(np.sum(lnls)),
    }
    print_json(json.dumps(line))


# ----------------------------------------------------------------------------------------
# GPU legs
# ----------------------------------------------------------------------------------------
class Dist:
    """torch.distributed plumbing (NCCL): barriers and the scalar all-reduce of a step."""

    def __init__(self, world: int, local_rank: int):
        import torch

        self.torch = torch
        self.world = world
        torch.cuda.set_device(local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x: float, op) -> float:
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.MAX if self.dist else None)

    def sum(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.SUM if self.dist else None)

    def library_comm(self, lib, rank: int):
        """Hands the library its own communicator: the 128-byte NCCL id of rank 0 travels over
        torch.distributed, pll_gpu_comm_init does the rest.  From here on the pll.h calls return
        the sums over all ranks (the scalar all-reduce happens inside the library, on the
        partition's stream) - no Python all-reduce on the path."""
        import ctypes as C

        if self.dist is None:
            return False
        buf = C.create_string_buffer(128)
        if rank == 0:
            assert lib.pll_gpu_comm_unique_id(buf) == 1, lib.errmsg()
        t = self.torch.tensor(list(buf.raw), dtype=self.torch.uint8, device="cuda")
        self.dist.broadcast(t, 0)
        buf = C.create_string_buffer(bytes(t.cpu().tolist()), 128)
        assert lib.pll_gpu_comm_init(buf, self.world, rank) == 1, lib.errmsg()
        return True

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def leg_lists(lib, part, pidx, tips: int, sites: int, steps: int):
    """e2e with a DIFFERENT operations list every step: the virtual root moves to a random inner
    node, the partial traversal that re-orients the CLVs (pll_utree_traverse with a pruning
    callback, reference test/src/partial-traversal.c) is issued through the public API together
    with its P-matrices, and the edge lnL is read back.  No list repeats, so nothing replays a
    cached CUDA graph: build_plan + staging + the uncaptured launch are inside the timed calls.
    Wall clock of the three pll.h calls vs. the device time between CUDA events on the stream."""
    from libpll_b200 import trees as T

    T.bind(lib)
    tree = T.Tree(lib, newick=T.random_newick(tips, 4242))
    walker = T.RootWalker(lib, tree)
    rng = np.random.default_rng(7)

    def evaluate(root, timed):
        ops, mats, bls = walker.move(root)
        e = walker.edge(root)
        t0 = time.perf_counter()
        if timed:
            part.timer_start()
        part.update_prob_matrices(pidx, mats, bls)
        part.update_partials(ops)
        lnl = part.edge_loglikelihood(e[0], e[1], e[2], e[3], e[4], pidx)
        dev_ms = part.timer_stop() if timed else 0.0
        return lnl, len(ops), (time.perf_counter() - t0) * 1e3, dev_ms

    ops, mats, bls = walker.full(tree.root)
    part.update_prob_matrices(pidx, mats, bls)
    part.update_partials(ops)
    e = walker.edge(tree.root)
    lnl0 = part.edge_loglikelihood(e[0], e[1], e[2], e[3], e[4], pidx)

    def random_root():
        node = tree.node(tips + int(rng.integers(0, tree.inner)))
        for _ in range(int(rng.integers(0, 3))):
            node = node.contents.next
        return node

    for _ in range(3):
        evaluate(random_root(), False)
    part.reset_stats()
    wall, dev, nops, worst = [], [], [], 0.0
    for _ in range(steps):
        lnl, n, w_ms, d_ms = evaluate(random_root(), True)
        if n == 0:
            continue
        wall.append(w_ms)
        dev.append(d_ms)
        nops.append(n)
        worst = max(worst, abs(lnl - lnl0) / abs(lnl0))
    st = part.stats()
    out = {
        "what": "moving virtual root: a different partial traversal + its P-matrices + edge lnL per step, "
                "through pll_update_prob_matrices / pll_update_partials / pll_compute_edge_loglikelihood",
        "steps": len(wall), "operations_per_step_mean": float(np.mean(nops)), "operations_per_step_max": int(max(nops)),
        "wall_ms_per_step": float(np.mean(wall)), "device_ms_per_step": float(np.mean(dev)),
        "host_overhead_ms_per_step": float(np.mean(wall) - np.mean(dev)),
        "host_overhead_frac_of_device": float((np.mean(wall) - np.mean(dev)) / np.mean(dev)),
        "value": float(np.sum(nops)) * sites / (np.sum(wall) * 1e-3), "unit": UNIT,
        "graph_launches": st["graph_launches"], "gpu_launches": st["kernel_launches"],
        "h2d_bytes_per_step": st["h2d_bytes"] // max(len(wall), 1),
        "lnl_drift_rel": worst,
    }
    assert worst < 1e-9, f"moving the virtual root changed the likelihood (rel {worst})"
    return out, tree, walker


def c_caller_latencies():
    """The same pll.h calls issued by a C program (tools/newton_c.c, built by build()): per-call
    latency without the ctypes argument marshalling that the Python loop of this leg pays."""
    exe = os.path.join(ROOT, "tools", "newton_c")
    if not os.path.exists(exe):
        return None
    try:
        run = subprocess.run([exe, "64", "1000000"], capture_output=True, text=True, timeout=120)
        out = {"program": "tools/newton_c.c: 64-taxon ladder x 1M patterns, GTR+G4, one inner-inner edge"}
        for line in run.stdout.splitlines():
            f = line.split()
            if line.startswith("pll_") and "us" in f:
                out[f[0] + "_us"] = float(f[f.index("us") - 1])
            elif line.startswith("Newton"):
                out["newton_32_iterations_ms"] = float(f[f.index("ms") - 1])
        return out
    except Exception as e:  # pragma: no cover
        return {"error": repr(e)}


def leg_c4(lib, part, pidx, tree, walker, sites: int, branches: int = 10, iters: int = 32):
    """BASELINE configs[3]: Newton branch-length optimisation (reference examples/newton/newton.c
    :31-100) on the evaluation edge and on `branches - 1` further edges reached by re-rooting
    (each move = a short partial traversal): pll_update_sumtable once per edge, then `iters`
    pll_compute_likelihood_derivatives calls with the Newton update between them."""
    import ctypes as C

    def ring(rec):
        out, n = [rec], rec.contents.next
        while n and C.addressof(n.contents) != C.addressof(rec.contents):
            out.append(n)
            n = n.contents.next
        return out

    root = tree.root
    seen_edges = set()
    key = np.zeros(8)  # under the GPU flag the sumtable argument is only a key (include/pll.h)
    t_move = t_sum = t_der = 0.0
    n_der = 0
    lengths = []
    per_branch = []
    # warm-up, untimed: the first sumtable of a partition allocates its 128 MB slot and loads the
    # level-by-level kernels (the traversal of this benchmark runs the fused one): ~90 ms once
    e0 = walker.edge(root)
    for _ in range(2):
        part.update_sumtable(e0[0], e0[2], e0[1], e0[3], pidx, key)
        part.likelihood_derivatives(e0[1], e0[3], root.contents.length, pidx, key)
    part.synchronize()
    t_all = time.perf_counter()
    for b in range(branches):
        if b:
            # next edge: another record of this inner node, or one of the node across the edge
            here = ring(root)
            across = ring(root.contents.back) if root.contents.back.contents.next else []
            root = next(r for r in here[1:] + across[1:] + here if r.contents.pmatrix_index not in seen_edges)
        seen_edges.add(root.contents.pmatrix_index)
        t0 = time.perf_counter()
        ops, mats, bls = walker.move(root)
        part.update_prob_matrices(pidx, mats, bls)
        part.update_partials(ops)
        part.synchronize()
        t_move += time.perf_counter() - t0
        e = walker.edge(root)
        t0 = time.perf_counter()
        part.update_sumtable(e[0], e[2], e[1], e[3], pidx, key)
        part.synchronize()
        t_sum += time.perf_counter() - t0
        length = root.contents.length
        t0 = time.perf_counter()
        for _ in range(iters):
            d1, d2 = part.likelihood_derivatives(e[1], e[3], length, pidx, key)
            n_der += 1
            # Newton where -lnL is convex, otherwise a bounded move downhill
            length = length - d1 / d2 if d2 > 0 else (length * 0.5 if d1 > 0 else length * 2.0)
            length = min(max(length, 1e-6), 10.0)
        t_der += time.perf_counter() - t0
        lengths.append(length)
        per_branch.append((len(ops), t_move, t_sum, t_der))
    total = time.perf_counter() - t_all
    prev = (0.0, 0.0, 0.0)
    for b, (n_ops, a, c, d) in enumerate(per_branch):   # where a slow branch spent its time
        log(f"c4 branch {b}: {n_ops} operations, move {1e6 * (a - prev[0]):.0f} us, "
            f"sumtable {1e6 * (c - prev[1]):.0f} us, {iters} derivative calls {1e6 * (d - prev[2]):.0f} us")
        prev = (a, c, d)
    return {
        "c_caller": c_caller_latencies(),
        "what": f"Newton on {branches} branches (the evaluation edge + {branches - 1} re-rooted ones): "
                f"partial traversal, pll_update_sumtable, {iters} x pll_compute_likelihood_derivatives each",
        "branches": branches, "iterations_per_branch": iters,
        "value": n_der / total, "unit": "Newton iterations/s",
        "derivative_call_us": t_der / n_der * 1e6, "sumtable_call_us": t_sum / branches * 1e6,
        "reroot_call_us": t_move / branches * 1e6, "newton_32_iterations_ms": t_der / branches * 1e3,
        "ms_per_branch": total / branches * 1e3,
        "derivative_pass_GBps": (4 * 4 * 8 + 4) * sites / (t_der / n_der) / 1e9,
        "optimised_lengths": [round(x, 6) for x in lengths[:3]],
    }


def c3_recycled(lib, w, steps: int, slots: int = 64):
    """The C3 list with its inner CLVs / scale buffers in a pool of recycled slots (the memory-saving
    mode of large analyses): the library's default for such lists - the whole-list walk k_walk_aa,
    which stores only the last value of a buffer - next to the level-by-level kernels forced with
    PLL_GPU_FUSED_AA=0.  Same lnL from both (and from the one-slot-per-node list)."""
    from libpll_b200 import synthetic as S
    from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP

    wr = S.recycle_slots(w, slots)
    out = {"slots": slots, "operations": len(wr.ops)}
    saved = os.environ.get("PLL_GPU_FUSED_AA")
    try:
        for name, env in (("default", None), ("level_by_level", "0")):
            if env is None:
                os.environ.pop("PLL_GPU_FUSED_AA", None)
            else:
                os.environ["PLL_GPU_FUSED_AA"] = env
            part, pidx = S.build_partition(lib, wr, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
            root = (wr.root_a, wr.scaler_of(wr.root_a), wr.root_b, wr.scaler_of(wr.root_b), wr.root_matrix, pidx)
            part.update_prob_matrices(pidx, wr.matrix_indices, wr.branch_lengths)
            for _ in range(3):
                part.update_partials(wr.ops)
                lnl = part.edge_loglikelihood(*root)
            part.reset_stats()
            part.timer_start()
            for _ in range(steps):
                part.update_partials(wr.ops)
                lnl = part.edge_loglikelihood(*root)
            ms = part.timer_stop() / steps
            st = part.stats()
            part.destroy()
            out[name] = {"ms_per_step": ms, "value": len(wr.ops) * w.sites / (ms * 1e-3), "unit": UNIT,
                         "kernel": "k_walk_aa" if st["kernel_launches"] <= 8 * steps else "k_partial_dmma_aa / k_partial_tt_aa",
                         "gpu_launches_per_step": st["kernel_launches"] / steps,
                         "compulsory_GB_per_step": st["compulsory_bytes"] / steps / 1e9, "lnl": lnl}
    finally:
        if saved is None:
            os.environ.pop("PLL_GPU_FUSED_AA", None)
        else:
            os.environ["PLL_GPU_FUSED_AA"] = saved
    return out


def leg_c3(lib, steps: int, warmup: int, local_rank: int):
    """BASELINE configs[2]: 500 taxa x 200k patterns, LG+G4 protein, traversal + edge lnL."""
    from libpll_b200 import synthetic as S
    from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP

    tips, sites = 500, 200_000
    w = S.make_workload(tips, sites, states=20)
    part, pidx = S.build_partition(lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
    root = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
    part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    sampler = ClockSampler(local_rank).start()
    for _ in range(max(warmup, 3)):
        part.update_partials(w.ops)
        lnl = part.edge_loglikelihood(*root)
    part.reset_stats()
    part.timer_start()
    for _ in range(steps):
        part.update_partials(w.ops)
        lnl = part.edge_loglikelihood(*root)
    ms = part.timer_stop() / steps
    stats = part.stats()
    part.reset_stats()
    part.timer_start()
    for _ in range(steps):
        part.update_partials(w.ops)
    trav_ms = part.timer_stop() / steps
    trav = part.stats()
    t0 = time.perf_counter()
    for i in range(steps):
        bl = w.branch_lengths * (1.0 + 1e-3 * ((i % 7) - 3))
        part.update_prob_matrices(pidx, w.matrix_indices, bl)
        part.update_partials(w.ops.copy())
        part.edge_loglikelihood(*root)
    e2e_s = (time.perf_counter() - t0) / steps
    clocks = sampler.stop()
    part.destroy()
    recycled = c3_recycled(lib, w, steps)
    tt, ti, ii = w.op_kinds()
    peak, peak_src = hbm_peak()
    comp = trav["compulsory_bytes"] / steps
    alg = trav["algorithmic_bytes"] / steps
    issued = sites / 8 * 4 * 15 * 512.0 * (2 * ii + ti)       # DMMA flops issued (N padded 20 -> 24)
    useful = sites * 4 * 20 * 20 * 2.0 * (2 * ii + ti)
    fused = trav["kernel_launches"] <= 3 * steps
    return {
        "workload": f"synthetic {tips}-taxon x {sites}-pattern LG+G4 protein, full traversal + edge logL",
        "operations": len(w.ops), "ops_tt_ti_ii": [tt, ti, ii],
        "value": len(w.ops) * sites / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
        "e2e": {"value": len(w.ops) * sites / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3},
        "gpu_launches": stats["kernel_launches"], "clocks": clocks, "lnl": lnl,
        "recycled_slots": recycled,
        "roofline": {
            "kernel": "k_walk_aa (whole list, FP64 tensor cores)" if fused else "k_partial_dmma_aa (level by level)",
            "avg_traversal_ms": trav_ms,
            "hbm": {"compulsory_bytes_per_launch": comp, "achieved": comp / (trav_ms * 1e-3) / 1e9, "peak": peak,
                    "unit": "GB/s", "frac": comp / (trav_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src},
            "tensor": {"dmma_flops_issued": issued, "useful_flops": useful,
                       "achieved": issued / (trav_ms * 1e-3) / 1e12, "peak": DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                       "frac": issued / (trav_ms * 1e-3) / 1e12 / DMMA_PEAK_TFLOPS,
                       "peak_source": "measured DMMA m8n8k4 peak (tools/ubench/dmma_chains.cu)"},
            "algorithmic": {"bytes_per_launch": alg, "GBps": alg / (trav_ms * 1e-3) / 1e9},
        },
    }


def leg_c5(lib, D: Dist, rank: int, world: int, local_rank: int, steps: int, slots: int):
    """BASELINE configs[4]: 5,000 taxa x 10M patterns sharded over the GPUs (strong scaling: the
    alignment is fixed), CLVs / scalers in a pool of recycled slots, tips generated on the device
     from a counter-based generator the host can restate for any slice."""
    from libpll_b200 import synthetic as S
    from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP

    tips, total = 5000, 10_000_000
    per = (total // world + 63) // 64 * 64
    lo, hi = rank * per, min((rank + 1) * per, total)
    w = S.recycle_slots(S.make_workload(tips, 64, states=4), slots)   # tree / operations only
    t0 = time.time()
    part = lib.partition(tips=tips, clv_buffers=w.inner, states=4, sites=hi - lo, rate_matrices=1,
                         prob_matrices=w.prob_matrices, rate_cats=4, scale_buffers=w.inner,
                         attributes=PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
    part.set_frequencies(0, S.GTR_FREQS)
    part.set_subst_params(0, S.GTR_RATES)
    part.set_category_rates(lib.gamma_rates(w.alpha, 4))
    part.set_category_weights(np.full(4, 0.25))
    for t in range(tips):
        assert lib.pll_gpu_generate_tip_states(part.ptr, t, 43, lo) == 1, lib.errmsg()
    pidx = np.zeros(4, np.uint32)
    part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    root = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
    setup_s = time.time() - t0

    def step():
        part.update_partials(w.ops)
        return part.edge_loglikelihood(*root)          # summed over the ranks inside the library

    sampler = ClockSampler(local_rank).start()
    for _ in range(3):
        lnl = step()
    D.barrier()
    part.reset_stats()
    part.timer_start()
    for _ in range(steps):
        lnl = step()
    ms = D.max(part.timer_stop()) / steps
    D.barrier()
    clocks = sampler.stop()
    st = part.stats()
    part.destroy()
    peak, peak_src = hbm_peak()
    comp = st["compulsory_bytes"] / steps
    return {
        "workload": f"synthetic {tips}-taxon x {total}-pattern GTR+G4 DNA sharded over {world} GPU(s) "
                    f"({hi - lo} patterns on rank 0), {slots} recycled CLV slots, device-generated tips",
        "scaling": "strong", "n_gpus": world, "operations": len(w.ops),
        "value": len(w.ops) * total / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
        "setup_s": setup_s, "gpu_launches": st["kernel_launches"], "clocks": clocks, "lnl": lnl,
        "roofline": {"kernel": "k_traverse_dna", "compulsory_bytes_per_launch_per_gpu": comp,
                     "achieved": comp / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": comp / (ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                     "note": "per GPU; only the last value of each recycled buffer and tile-cache misses are "
                             "stored (dead-store elimination), tip characters are read",
                     "algorithmic_GBps_per_gpu": st["algorithmic_bytes"] / steps / (ms * 1e-3) / 1e9},
    }


def run_gpu(args, rank: int, world: int, local_rank: int):
    import libpll_b200
    from libpll_b200 import synthetic as S
    from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP

    lib = libpll_b200.load()
    if lib.plg_device_count() == 0:
        raise SystemExit("bench.py: no B200 visible - the CUDA path has no CPU fallback")
    D = Dist(world, local_rank)  # torch: plumbing only (NCCL scalar all-reduce, barriers)
    lib.pll_gpu_set_device(local_rank)
    in_library = D.library_comm(lib, rank)   # N > 1: lnL / derivatives are all-reduced inside the library
    memoize_tips()
    also = args.also
    if also == "auto":
        also = "lists,c4,c3,c5" if world == 1 else "c5"
    also = [x for x in also.split(",") if x and x != "none"]

    S_gpu = args.sites_per_gpu
    w = S.make_workload(args.tips, S_gpu * world, states=4)
    lo, hi = rank * S_gpu, (rank + 1) * S_gpu
    t0 = time.time()
    part, pidx = S.build_partition(lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP, lo=lo, hi=hi)
    if rank == 0:
        log(f"[bench] setup {time.time() - t0:.1f}s: {args.tips} tips x {S_gpu} patterns per GPU, "
            f"ops tt/ti/ii = {w.op_kinds()}, {w.algorithmic_bytes_per_site()} algorithmic B per pattern")
    root = (w.root_a, w.scaler_of(w.root_a), w.root_b, w.scaler_of(w.root_b), w.root_matrix, pidx)
    n_ops = len(w.ops)

    # ---- leg A: resident (device-timed) ------------------------------------------------
    part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)

    def resident_step():
        part.update_partials(w.ops)
        return part.edge_loglikelihood(*root)          # N > 1: already the sum over the ranks

    # clocks are sampled from the warm-up on: nvidia-smi needs ~0.2 s to deliver its first
    # sample and the timed region of a short run is not much longer (same load throughout)
    sampler = ClockSampler(local_rank).start()
    for _ in range(max(args.warmup, 3)):
        lnl = resident_step()
    D.barrier()
    part.reset_stats()
    part.timer_start()
    for _ in range(args.steps):
        lnl = resident_step()
    ms = part.timer_stop()
    D.barrier()
    clocks = sampler.stop()
    stats = part.stats()
    ms_per_step = D.max(ms) / args.steps
    value = n_ops * S_gpu * world / (ms_per_step * 1e-3)

    # ---- leg B: end to end through the public API, host buffers, wall clock -------------
    def e2e_step(i):
        # fresh host inputs every step: branch lengths change, so every P-matrix is recomputed
        bl = w.branch_lengths * (1.0 + 1e-3 * ((i % 7) - 3))
        ops = w.ops.copy()
        part.update_prob_matrices(pidx, w.matrix_indices, bl)
        part.update_partials(ops)
        return part.edge_loglikelihood(*root)

    for i in range(max(args.warmup, 3)):
        e2e_step(i)
    D.barrier()
    part.reset_stats()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    D.barrier()
    e2e_s = D.max(time.perf_counter() - t0) / args.steps
    e2e_stats = part.stats()
    e2e_value = n_ops * S_gpu * world / e2e_s

    # ---- leg C: roofline of the dominant kernel, CUDA events per launch ------------------
    # (1) the traversal as the library runs it: ONE kernel walks the whole list
    #     (k_traverse_dna, libpll_b200/csrc/gpu/plg_traverse.cu); timed alone, on its stream
    n_trav = max(args.steps, 3)
    part.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
    for _ in range(2):
        part.update_partials(w.ops)
    part.reset_stats()
    trav_ms = 0.0
    for _ in range(n_trav):
        part.timer_start()              # CUDA events on the library's stream around ONE call:
        part.update_partials(w.ops)     # k_fused_pack (6 us) + k_traverse_dna
        trav_ms += part.timer_stop()
    trav_ms = D.max(trav_ms) / n_trav
    trav_stats = part.stats()
    fused = trav_stats["kernel_launches"] <= 3 * n_trav  # pack + traverse (+ nothing else) per call
    alg_bytes = trav_stats["algorithmic_bytes"] / n_trav
    comp_bytes = trav_stats["compulsory_bytes"] / n_trav
    # the edge-lnL call of a step, alone: what a step spends OUTSIDE the traversal kernel
    part.timer_start()
    for _ in range(n_trav):
        part.edge_loglikelihood(*root)
    lnl_call_ms = D.max(part.timer_stop()) / n_trav
    # the traversal kernel's launch duration INSIDE the timed region = step - lnL call (the separate
    # loop above runs later, hotter and - sustained - under the board's power cap: it is reported
    # next to it, not used for the roofline)
    trav_alone_ms = trav_ms
    trav_ms = max(ms_per_step - lnl_call_ms, 1e-6)
    persite = np.zeros(S_gpu)
    lnl = part.edge_loglikelihood(*root, persite=persite)   # same state as the resident leg
    # (2) the level-by-level kernels (one launch per dependency level and kind), per-kind times
    part.set_profiling(True)
    part.reset_stats()
    for _ in range(3):
        part.update_partials(w.ops)
    prof = part.stats()
    part.set_profiling(False)
    kinds = ["tip-tip", "tip-inner", "inner-inner"]
    KERNEL_NAMES = ["k_partial_tt_dna", "k_partial_stream_dna<TI>", "k_partial_stream_dna<II>"]
    shares = {}
    tot_ns = sum(prof["kind_ns"]) or 1
    for i, name in enumerate(kinds):
        if prof["kind_launches"][i]:
            shares[name] = {
                "share_of_traversal": prof["kind_ns"][i] / tot_ns,
                "GBps": prof["kind_bytes"][i] / max(prof["kind_ns"][i], 1),
                "launches_per_traversal": prof["kind_launches"][i] // 3,
            }
    dom = int(np.argmax(prof["kind_ns"]))
    peak, peak_src = hbm_peak()
    level_achieved = prof["kind_bytes"][dom] / max(prof["kind_ns"][dom], 1)
    level_path = {
        "kernel": f"{KERNEL_NAMES[dom]} ({kinds[dom]})", "achieved": level_achieved, "frac": level_achieved / peak,
        "algorithmic_bytes_per_launch": prof["kind_bytes"][dom] / max(prof["kind_launches"][dom], 1),
        "avg_launch_ms": prof["kind_ns"][dom] / max(prof["kind_launches"][dom], 1) * 1e-6,
        "traversal_ms": tot_ns / 3 * 1e-6, "by_kind": shares,
    }
    if fused:
        achieved = comp_bytes / (trav_ms * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic("c2_k_traverse_dna", tips=args.tips, sites=S_gpu)
        roofline = {
            "bound": "hbm", "kernel": "k_traverse_dna (the whole operations list in one launch)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
            "traffic": traffic, "traffic_source": traffic_src,
            "compulsory_bytes_per_launch": comp_bytes, "avg_launch_ms": trav_ms,
            "avg_launch_ms_source": "CUDA events over the timed region: ms_per_step minus the edge-lnL call "
                                    f"({lnl_call_ms:.3f} ms, timed alone); the kernel timed alone afterwards: "
                                    f"{trav_alone_ms:.3f} ms",
            "bytes_model": "compulsory DRAM bytes of the executed plan: every observable parent CLV + scaler written "
                           "once, tip characters and tile-cache misses read (plg_stats.compulsory_bytes)",
            "algorithmic": {"bytes_per_launch": alg_bytes, "GBps": alg_bytes / (trav_ms * 1e-3) / 1e9,
                            "note": "SURVEY 8d counts every child CLV read; the kernel keeps fresh tiles on chip, "
                                    "so this figure exceeds the DRAM-level one"},
            "level_by_level_path": level_path,
        }
    else:
        roofline = {
            "bound": "hbm", "kernel": level_path["kernel"], "achieved": level_achieved,
            "peak": peak, "unit": "GB/s", "frac": level_achieved / peak, "peak_source": peak_src,
            "traffic": None, "algorithmic_bytes_per_launch": level_path["algorithmic_bytes_per_launch"],
            "avg_launch_ms": level_path["avg_launch_ms"], "by_kind": shares,
        }
    roofline["frac_of_8TBps_nominal"] = roofline["achieved"] * 1e9 / 8e12

    # ---- the other configurations ---------------------------------------------------------
    extra = {}
    tree = walker = None
    if "lists" in also or "c4" in also:
        try:
            extra["lists"], tree, walker = leg_lists(lib, part, pidx, args.tips, S_gpu, steps=24)
            if "lists" not in also:
                extra.pop("lists")
        except Exception as e:  # pragma: no cover
            extra["lists"] = {"error": repr(e)}
    if "c4" in also and tree is not None:
        try:
            extra["c4"] = leg_c4(lib, part, pidx, tree, walker, S_gpu)
        except Exception as e:  # pragma: no cover
            extra["c4"] = {"error": repr(e)}
    if tree is not None:
        tree.destroy()
    part.destroy()
    if "c3" in also and rank == 0:
        try:
            extra["c3"] = leg_c3(lib, steps=5, warmup=3, local_rank=local_rank)
        except Exception as e:  # pragma: no cover
            extra["c3"] = {"error": repr(e)}
    if "c5" in also:
        try:
            extra["c5"] = leg_c5(lib, D, rank, world, local_rank, steps=3, slots=args.slots)
        except Exception as e:  # pragma: no cover
            extra["c5"] = {"error": repr(e)}

    # ---- CPU baseline (rank 0, N = 1 only): the whole workload on all host cores ----------
    cpu = None
    lnl_check = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        try:
            if args.cpu_sites_per_thread:
                spt = args.cpu_sites_per_thread
                slices = [(t * spt, (t + 1) * spt) for t in range(cores)]
                rate, best, lnls, kind = cpu_traversal(S.recycle_slots(w, CPU_SLOTS), slices, reps=2, warm=1)
                sample = f"{cores} threads x {spt} patterns each ({cores * spt} of {w.sites})"
            else:
                rate, best, lnls, kind, slices = full_config_cpu(w, cores, reps=2, warm=1, use_mean=False)
                sample = f"ALL {w.sites} patterns: {cores} threads x {slices[0][1] - slices[0][0]} patterns each"
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "cpu_model": cpu_model(),
                   "sample": sample + f", {n_ops} operations per traversal, inner CLVs in {CPU_SLOTS} recycled slots "
                                      f"per thread, best of 2, {best:.2f} s per step"}
            cpu_lnl = float(np.sum(lnls))
            gpu_lnl = float(sum(persite[a:b].sum() for a, b in slices))
            rel = abs(gpu_lnl - cpu_lnl) / abs(cpu_lnl)
            lnl_check = {"gpu": gpu_lnl, "cpu": cpu_lnl, "rel": rel, "ok": bool(rel <= 1e-10),
                         "over": f"{sum(b - a for a, b in slices)} patterns ({len(slices)} CPU slices)"}
        except Exception as e:  # pragma: no cover
            cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "unavailable", "sample": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_string(args.tips, S_gpu * world, S_gpu),
                "attributes": "PLL_ATTRIB_ARCH_GPU|PLL_ATTRIB_PATTERN_TIP, per-site scalers",
                "operations": n_ops, "rate_cats": 4,
                "l2": "no flush needed: each step writes %.0f GB per GPU through a 126 MB L2" % (comp_bytes / 1e9),
                "sharding": ("site patterns; scalar ncclAllReduce of lnL inside the library (pll_gpu_comm_init), "
                             f"{stats['collectives']} all-reduces in the timed region") if world > 1 else "single GPU",
            },
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": e2e_stats["h2d_bytes"] // args.steps,
                    "d2h_bytes_per_step": e2e_stats["d2h_bytes"] // args.steps,
                    "ms_per_step": e2e_s * 1e3},
            "lnl_evals_per_s": 1.0 / e2e_s,
            "gpu_launches": stats["kernel_launches"],
            "graph_launches": stats["graph_launches"],
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "lnl": lnl,
            "lnl_check": lnl_check,
            "also": extra,
        }
        print_json(json.dumps(line))
        if lnl_check is not None and not lnl_check["ok"]:
            D.close()
            raise SystemExit(f"bench.py: GPU lnL differs from the reference's: {lnl_check}")
    if in_library:
        lib.pll_gpu_comm_finalize()
    D.close()


def main():
    # Only the JSON line may reach stdout: libraries (NCCL's version banner, for one) write
    # there too, so fd 1 is pointed at stderr for the whole run and the line goes to the
    # saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: str):
        os.write(real_stdout, (line + "\n").encode())

    global print_json
    print_json = emit
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--tips", type=int, default=1000)
    ap.add_argument("--sites-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--cpu-sites-per-thread", type=int, default=0,
                    help="GPU arm's cpu_baseline: 0 = the whole workload (default), else a bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--also", default="auto",
                    help="comma list of sub-records: lists,c4,c3,c5 | none | auto (1 GPU: all; N > 1: c5)")
    ap.add_argument("--slots", type=int, default=64, help="c5: recycled CLV / scaler slots")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
