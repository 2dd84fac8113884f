import sys, os, itertools
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
from libpll_b200 import synthetic as S
import libpll_b200
from libpll_b200.binding import *
gpu = libpll_b200.load()
ref = PllLibrary("oracle/_ref/libpll_ref.so", is_gpu=False)
RTOL = 1e-10
bad = n = 0
for states, sites, rs, slices, cats in itertools.product((4, 20), (1, 2, 3, 7, 31, 33, 63, 65, 127, 129, 255, 257, 1000, 4097), (0, PLL_ATTRIB_RATE_SCALERS), (1, 2), (4, 1)):
    if cats == 1 and rs: continue
    w = S.make_workload(40, sites, states=states, rate_cats=cats, seed=sites)
    rates = ref.gamma_rates(w.alpha, w.rate_cats)
    extra = PLL_ATTRIB_PATTERN_TIP | rs
    tag = f"K={states} sites={sites} rs={int(bool(rs))} slices={slices} cats={cats}"
    try:
        gpu.pll_gpu_set_devices(slices)
        try:
            pg, pidx = S.build_partition(gpu, w, PLL_ATTRIB_ARCH_GPU | extra, rates=rates)
        finally:
            gpu.pll_gpu_set_devices(0)
        pr, _ = S.build_partition(ref, w, PLL_ATTRIB_ARCH_AVX2 | extra, rates=rates)
        for p in (pg, pr):
            p.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
            p.update_partials(w.ops)
        a, b = w.root_a, w.root_b
        sg, sr = np.zeros(sites), np.zeros(sites)
        args = (a, w.scaler_of(a), b, w.scaler_of(b), w.root_matrix, pidx)
        eg, er = pg.edge_loglikelihood(*args, persite=sg), pr.edge_loglikelihood(*args, persite=sr)
        res = [("edge", eg, er, abs(er))]
        if not np.allclose(sg, sr, rtol=RTOL, atol=0): res.append(("edge persite", float(np.abs(sg - sr).max()), 0.0, 1.0))
        for k in range(w.inner):
            cg, cr = pg.get_clv(w.tips + k), pr.get_clv(w.tips + k)
            if states == 4:
                if not np.array_equal(cg, cr): res.append((f"clv {k}", float(np.abs(cg - cr).max()), 0.0, 1.0)); break
            elif not np.allclose(cg, cr, rtol=1e-12, atol=0): res.append((f"clv {k}", float(np.abs(cg - cr).max()), 0.0, 1.0)); break
        tg, tr = pg.new_sumtable(), pr.new_sumtable()
        pg.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, tg)
        pr.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, tr)
        dg = pg.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), 0.3, pidx, tg)
        dr = pr.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), 0.3, pidx, tr)
        sc = max(abs(dr[0]), float(w.weights.sum()) * 1e-3)
        res += [("d1", dg[0], dr[0], sc), ("d2", dg[1], dr[1], max(abs(dr[1]), sc))]
        for r in res:
            n += 1
            if not (np.isfinite(r[2]) and abs(r[1] - r[2]) <= RTOL * max(r[3], 1e-300)):
                bad += 1; print("MISMATCH", tag, r, flush=True)
        pg.destroy(); pr.destroy()
    except Exception as e:
        bad += 1; print("ERROR", tag, type(e).__name__, str(e)[:300], flush=True)
print(f"checked {n} values, {bad} problems")
