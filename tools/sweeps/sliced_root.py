import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
from libpll_b200 import synthetic as S
import libpll_b200
from libpll_b200.binding import *
from test_parity_gpu import _caterpillar
gpu = libpll_b200.load()
ref = PllLibrary("oracle/_ref/libpll_ref.so", is_gpu=False)
for abflag in (0, PLL_ATTRIB_AB_FLAG):
  for rs in (0, PLL_ATTRIB_RATE_SCALERS):
    for slices in (1, 3):
      for sites in (150, 192):
        os.environ["PLL_GPU_DEVICES"] = str(slices)
        w = _caterpillar(300, sites, 4, seed=5)
        rates = ref.gamma_rates(w.alpha, w.rate_cats)
        extra = PLL_ATTRIB_PATTERN_TIP | rs | abflag
        pg, pidx = S.build_partition(gpu, w, PLL_ATTRIB_ARCH_GPU | extra, rates=rates)
        pr, _ = S.build_partition(ref, w, PLL_ATTRIB_ARCH_AVX2 | extra, rates=rates)
        for p in (pg, pr):
            p.update_prob_matrices(pidx, w.matrix_indices, w.branch_lengths)
            p.update_partials(w.ops)
        top = w.tips + w.inner - 1
        rg = pg.root_loglikelihood(top, w.scaler_of(top), pidx); rr = pr.root_loglikelihood(top, w.scaler_of(top), pidx)
        a, b = w.root_a, w.root_b
        eg = pg.edge_loglikelihood(a, w.scaler_of(a), b, w.scaler_of(b), w.root_matrix, pidx)
        er = pr.edge_loglikelihood(a, w.scaler_of(a), b, w.scaler_of(b), w.root_matrix, pidx)
        sg = np.asarray(pg.get_scaler(w.scaler_of(top))); sr = np.asarray(pr.get_scaler(w.scaler_of(top)))
        print(f"ab={abflag} rs={rs} slices={slices}/{gpu.pll_gpu_partition_devices(pg.ptr)} sites={sites}: root diff {rg-rr:.6g} edge diff {eg-er:.6g} scalers equal {np.array_equal(sg, sr)}", flush=True)
        pg.destroy(); pr.destroy()
