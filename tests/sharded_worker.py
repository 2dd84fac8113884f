"""Worker of tests/test_sharded_nccl_gpu.py: one rank of a site-sharded evaluation whose scalar
results are combined INSIDE the library (pll_gpu_comm_init: ncclAllReduce of 1-2 doubles on the
partition's stream).  No torch: the 128-byte NCCL id travels through a file, as an MPI program
would broadcast it.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

rank, world, sites, id_path = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
states = int(sys.argv[5]) if len(sys.argv) > 5 else 4

import numpy as np  # noqa: E402

import libpll_b200  # noqa: E402
from libpll_b200 import sharding, synthetic as S  # noqa: E402
from libpll_b200.binding import PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP  # noqa: E402

lib = libpll_b200.load()
lib.pll_gpu_set_device(rank)
buf = C.create_string_buffer(128)
if rank == 0:
    assert lib.pll_gpu_comm_unique_id(buf) == 1, lib.errmsg()
    with open(id_path + ".tmp", "wb") as f:
        f.write(buf.raw)
    os.replace(id_path + ".tmp", id_path)
else:
    t0 = time.time()
    while not os.path.exists(id_path):
        assert time.time() - t0 < 120, "rank 0 never published the communicator id"
        time.sleep(0.01)
    buf = C.create_string_buffer(open(id_path, "rb").read(), 128)
assert lib.pll_gpu_comm_init(buf, world, rank) == 1, lib.errmsg()

w = S.make_workload(24, sites, states=states, seed=23)
lo, hi = sharding.slice_bounds(sites, world, rank)
part, pidx = S.build_partition(lib, w, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP, lo=lo, hi=hi)
lnl = S.full_evaluation(part, w, pidx)                  # already the sum over all ranks
a, b = w.root_a, w.root_b
persite = np.zeros(hi - lo)
lnl2 = part.edge_loglikelihood(a, w.scaler_of(a), b, w.scaler_of(b), w.root_matrix, pidx, persite=persite)
top = a if a >= w.tips else b
root = part.root_loglikelihood(top, w.scaler_of(top), pidx)
tab = part.new_sumtable()
part.update_sumtable(a, b, w.scaler_of(a), w.scaler_of(b), pidx, tab)
derivs = [part.likelihood_derivatives(w.scaler_of(a), w.scaler_of(b), t, pidx, tab) for t in (0.05, 0.3)]
st = part.stats()
part.destroy()
lib.pll_gpu_comm_finalize()
print(json.dumps({"rank": rank, "lo": lo, "hi": hi, "lnl": lnl, "lnl2": lnl2, "local_lnl": float(persite.sum()),
                  "root": root, "derivs": derivs, "collectives": st["collectives"]}))
