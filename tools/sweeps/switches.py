import sys, os, itertools
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
from libpll_b200 import synthetic as S
import libpll_b200
from libpll_b200.binding import *
from test_parity_gpu import _caterpillar
gpu = libpll_b200.load()
ref = PllLibrary("oracle/_ref/libpll_ref.so", is_gpu=False)
RTOL = 1e-10
bad = 0; n = 0
def close(a, b, scale=None):
    s = abs(b) if scale is None else scale
    return np.isfinite(b) and abs(a - b) <= RTOL * max(s, 1e-300)
for states, ptip, rs, slices, pinv, ab in itertools.product(tuple(int(x) for x in os.environ.get('SWEEP_STATES','4,20').split(',')), (True, False), (0, PLL_ATTRIB_RATE_SCALERS), (1, 3), (0.0, 0.2),
                                                          (None, 0, PLL_ATTRIB_AB_LEWIS, PLL_ATTRIB_AB_FELSENSTEIN, PLL_ATTRIB_AB_STAMATAKIS)):
    if ab and pinv: continue          # incompatible in the reference too
    if states != 4 and ptip and ab is not None: continue
    if states not in (4, 20) and ptip and rs: continue   # plain-C tip-inner ignores per-rate scalers (src/core_partials.c:461-510)   # the reference reads past its tables there (src/pll.c:885-903)
    if states == 4 and os.environ.get('SWEEP_SKIP_DNA'): continue
    tips, sites = (300, 150) if states <= 7 else (140, 150)
    w = _caterpillar(tips, sites, states, seed=7)
    w.rate_cats = int(os.environ.get('SWEEP_CATS', '4'))
    rates = ref.gamma_rates(w.alpha, w.rate_cats)
    extra = (PLL_ATTRIB_PATTERN_TIP if ptip else 0) | rs | (PLL_ATTRIB_AB_FLAG if ab is not None else 0)
    gpu.pll_gpu_set_devices(slices)
    try:
        pg, pidx = S.build_partition(gpu, w, PLL_ATTRIB_ARCH_GPU | extra, rates=rates)
    finally:
        gpu.pll_gpu_set_devices(0)
    pr, _ = S.build_partition(ref, w, (PLL_ATTRIB_ARCH_AVX2 if states in (4, 20) else PLL_ATTRIB_ARCH_CPU) | extra, rates=rates)
    tag = f"K={states} ptip={int(ptip)} rs={int(bool(rs))} slices={slices} pinv={pinv} ab={ab}"
    print(tag, flush=True)
    try:
        for p in (pg, pr):
            if pinv:
                p.update_invariant_sites()
                for i in set(int(x) for x in pidx): p.update_invariant_sites_proportion(i, pinv)
            if ab:
                p.set_asc_bias_type(ab); p.set_asc_state_weights(list(range(3, 3 + states)))
            p.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
            p.update_partials(w.ops)
        res = []
        top = w.tips + w.inner - 1
        a, b = w.root_a, w.root_b          # inner, tip
        inner_edge = None
        for op in w.ops[::-1]:
            c1, c2 = int(op["child1_clv_index"]), int(op["child2_clv_index"])
            if c1 >= w.tips and c2 >= w.tips: inner_edge = (c1, c2, int(op["child2_matrix_index"])); break
        if inner_edge is None:     # caterpillar: parent / child along the spine
            op = w.ops[-1]; inner_edge = (int(op["child1_clv_index"]), int(w.ops[-2]["child1_clv_index"]), int(op["child1_matrix_index"]))
        sg, sr = np.zeros(sites), np.zeros(sites)
        res.append(("root", pg.root_loglikelihood(top, w.scaler_of(top), pidx, persite=sg), pr.root_loglikelihood(top, w.scaler_of(top), pidx, persite=sr)))
        if not np.allclose(sg, sr, rtol=RTOL, atol=0): res.append(("root persite", float(np.abs(sg - sr).max()), 0.0))
        sg, sr = np.zeros(sites), np.zeros(sites)
        res.append(("edge", pg.edge_loglikelihood(a, w.scaler_of(a), b, w.scaler_of(b), w.root_matrix, pidx, persite=sg),
                    pr.edge_loglikelihood(a, w.scaler_of(a), b, w.scaler_of(b), w.root_matrix, pidx, persite=sr)))
        if not np.allclose(sg, sr, rtol=RTOL, atol=0): res.append(("edge persite", float(np.abs(sg - sr).max()), 0.0))
        for (x, y) in ((a, b),):
            tg, tr = pg.new_sumtable(), pr.new_sumtable()
            pg.update_sumtable(x, y, w.scaler_of(x), w.scaler_of(y), pidx, tg)
            pr.update_sumtable(x, y, w.scaler_of(x), w.scaler_of(y), pidx, tr)
            for t in (0.02, 0.7):
                dg = pg.likelihood_derivatives(w.scaler_of(x), w.scaler_of(y), t, pidx, tg)
                dr = pr.likelihood_derivatives(w.scaler_of(x), w.scaler_of(y), t, pidx, tr)
                sc = max(abs(dr[0]), float(w.weights.sum()) * 1e-3)
                res.append((f"d1 t={t}", dg[0], dr[0], sc)); res.append((f"d2 t={t}", dg[1], dr[1], max(abs(dr[1]), sc)))
        for r in res:
            n += 1
            ok = close(r[1], r[2], r[3] if len(r) > 3 else None)
            if not ok:
                bad += 1
                print("MISMATCH", tag, r, flush=True)
    except Exception as e:
        print("ERROR", tag, type(e).__name__, str(e)[:200], flush=True)
        bad += 1
    pg.destroy(); pr.destroy()
print(f"checked {n} values, {bad} problems")
