"""ctypes binding of the pll.h C API.

The SAME binding class drives two different shared objects:
  * libpll_b200/libpll_b200.so  - this repository's GPU implementation (the product), and
  * oracle/_ref/libpll_ref.so   - the unmodified reference compiled from /root/reference
                                   (test infrastructure only; loaded by tests/ and by
                                   bench.py's cpu_baseline / --impl reference legs).
Both export the reference's pll.h symbols (reference src/pll.h:530-653), so parity tests
read like the reference's own test programs: same calls, same arguments.

No computation happens here: every method is a 1:1 forward to the C entry point.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

# ---- constants (include/pll.h) -------------------------------------------------------
PLL_SUCCESS = 1
PLL_FAILURE = 0
PLL_ATTRIB_ARCH_CPU = 0
PLL_ATTRIB_AB_LEWIS = 1 << 5
PLL_ATTRIB_AB_FELSENSTEIN = 2 << 5
PLL_ATTRIB_AB_STAMATAKIS = 3 << 5
PLL_ATTRIB_AB_FLAG = 1 << 8
PLL_ATTRIB_ARCH_SSE = 1 << 0
PLL_ATTRIB_ARCH_AVX = 1 << 1
PLL_ATTRIB_ARCH_AVX2 = 1 << 2
PLL_ATTRIB_PATTERN_TIP = 1 << 4
PLL_ATTRIB_RATE_SCALERS = 1 << 9
PLL_ATTRIB_ARCH_GPU = 1 << 10
PLL_SCALE_BUFFER_NONE = -1
PLL_GAMMA_RATES_MEAN = 0
PLL_GAMMA_RATES_MEDIAN = 1

c_uint_p = C.POINTER(C.c_uint)
c_double_p = C.POINTER(C.c_double)


class PllPartition(C.Structure):
    """pll_partition_t (include/pll.h; layout of reference src/pll.h:202-244, 216 bytes)."""

    _fields_ = [
        ("tips", C.c_uint),
        ("clv_buffers", C.c_uint),
        ("states", C.c_uint),
        ("sites", C.c_uint),
        ("pattern_weight_sum", C.c_uint),
        ("rate_matrices", C.c_uint),
        ("prob_matrices", C.c_uint),
        ("rate_cats", C.c_uint),
        ("scale_buffers", C.c_uint),
        ("attributes", C.c_uint),
        ("alignment", C.c_size_t),
        ("states_padded", C.c_uint),
        ("clv", C.POINTER(c_double_p)),
        ("pmatrix", C.POINTER(c_double_p)),
        ("rates", c_double_p),
        ("rate_weights", c_double_p),
        ("subst_params", C.POINTER(c_double_p)),
        ("scale_buffer", C.POINTER(c_uint_p)),
        ("frequencies", C.POINTER(c_double_p)),
        ("prop_invar", c_double_p),
        ("invariant", C.POINTER(C.c_int)),
        ("pattern_weights", c_uint_p),
        ("eigen_decomp_valid", C.POINTER(C.c_int)),
        ("eigenvecs", C.POINTER(c_double_p)),
        ("inv_eigenvecs", C.POINTER(c_double_p)),
        ("eigenvals", C.POINTER(c_double_p)),
        ("maxstates", C.c_uint),
        ("tipchars", C.POINTER(C.POINTER(C.c_ubyte))),
        ("charmap", C.POINTER(C.c_ubyte)),
        ("ttlookup", c_double_p),
        ("tipmap", c_uint_p),
        ("asc_bias_alloc", C.c_int),
    ]


assert C.sizeof(PllPartition) == 216


class PllOperation(C.Structure):
    """pll_operation_t (reference src/pll.h:249-259, 32 bytes)."""

    _fields_ = [
        ("parent_clv_index", C.c_uint),
        ("parent_scaler_index", C.c_int),
        ("child1_clv_index", C.c_uint),
        ("child1_matrix_index", C.c_uint),
        ("child1_scaler_index", C.c_int),
        ("child2_clv_index", C.c_uint),
        ("child2_matrix_index", C.c_uint),
        ("child2_scaler_index", C.c_int),
    ]


assert C.sizeof(PllOperation) == 32

OP_DTYPE = np.dtype(
    [
        ("parent_clv_index", "<u4"),
        ("parent_scaler_index", "<i4"),
        ("child1_clv_index", "<u4"),
        ("child1_matrix_index", "<u4"),
        ("child1_scaler_index", "<i4"),
        ("child2_clv_index", "<u4"),
        ("child2_matrix_index", "<u4"),
        ("child2_scaler_index", "<i4"),
    ]
)
assert OP_DTYPE.itemsize == 32


class PlgStats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_ulonglong),
        ("graph_launches", C.c_ulonglong),
        ("h2d_bytes", C.c_ulonglong),
        ("d2h_bytes", C.c_ulonglong),
        ("partial_ops", C.c_ulonglong),
        ("partial_levels", C.c_ulonglong),
        ("algorithmic_bytes", C.c_ulonglong),
        ("kind_ns", C.c_ulonglong * 3),
        ("kind_bytes", C.c_ulonglong * 3),
        ("kind_launches", C.c_ulonglong * 3),
        ("compulsory_bytes", C.c_ulonglong),
        ("graph_evictions", C.c_ulonglong),
        ("collectives", C.c_ulonglong),
    ]


PART_P = C.POINTER(PllPartition)

# name -> (restype, argtypes); the reference's prototypes, reference src/pll.h:530-653
_PLL_API = {
    "pll_partition_create": (PART_P, [C.c_uint] * 9),
    "pll_partition_destroy": (None, [PART_P]),
    "pll_set_tip_states": (C.c_int, [PART_P, C.c_uint, c_uint_p, C.c_char_p]),
    "pll_set_tip_clv": (C.c_int, [PART_P, C.c_uint, c_double_p, C.c_int]),
    "pll_set_pattern_weights": (None, [PART_P, c_uint_p]),
    "pll_set_subst_params": (None, [PART_P, C.c_uint, c_double_p]),
    "pll_set_frequencies": (None, [PART_P, C.c_uint, c_double_p]),
    "pll_set_category_rates": (None, [PART_P, c_double_p]),
    "pll_set_category_weights": (None, [PART_P, c_double_p]),
    "pll_update_eigen": (C.c_int, [PART_P, C.c_uint]),
    "pll_update_prob_matrices": (C.c_int, [PART_P, c_uint_p, c_uint_p, c_double_p, C.c_uint]),
    "pll_count_invariant_sites": (C.c_uint, [PART_P, c_uint_p]),
    "pll_update_invariant_sites": (C.c_int, [PART_P]),
    "pll_update_invariant_sites_proportion": (C.c_int, [PART_P, C.c_uint, C.c_double]),
    "pll_update_partials": (None, [PART_P, C.c_void_p, C.c_uint]),
    "pll_compute_root_loglikelihood": (C.c_double, [PART_P, C.c_uint, C.c_int, c_uint_p, c_double_p]),
    "pll_compute_edge_loglikelihood": (
        C.c_double,
        [PART_P, C.c_uint, C.c_int, C.c_uint, C.c_int, C.c_uint, c_uint_p, c_double_p],
    ),
    "pll_update_sumtable": (C.c_int, [PART_P, C.c_uint, C.c_uint, C.c_int, C.c_int, c_uint_p, c_double_p]),
    "pll_compute_likelihood_derivatives": (
        C.c_int,
        [PART_P, C.c_int, C.c_int, C.c_double, c_uint_p, c_double_p, c_double_p, c_double_p],
    ),
    "pll_set_asc_bias_type": (C.c_int, [PART_P, C.c_int]),
    "pll_set_asc_state_weights": (None, [PART_P, c_uint_p]),
    "pll_compute_gamma_cats": (C.c_int, [C.c_double, C.c_uint, c_double_p, C.c_int]),
    "pll_compress_site_patterns": (c_uint_p, [C.POINTER(C.c_char_p), c_uint_p, C.c_int, C.POINTER(C.c_int)]),
    "pll_aligned_alloc": (C.c_void_p, [C.c_size_t, C.c_size_t]),
    "pll_aligned_free": (None, [C.c_void_p]),
}

# extensions only this repository's library has (include/pll_gpu.h)
_GPU_API = {
    "pll_gpu_compress_site_patterns": (c_uint_p, [C.POINTER(C.c_char_p), c_uint_p, C.c_int, C.POINTER(C.c_int)]),
    "pll_gpu_set_device": (C.c_int, [C.c_int]),
    "pll_gpu_device_count": (C.c_int, []),
    "pll_gpu_set_devices": (C.c_int, [C.c_int]),
    "pll_gpu_partition_devices": (C.c_int, [PART_P]),
    "pll_gpu_slice_bounds": (C.c_uint, [C.c_uint, C.c_uint, c_uint_p]),
    "pll_gpu_context_of": (C.c_void_p, [PART_P, C.c_uint, c_uint_p, c_uint_p]),
    "pll_gpu_context": (C.c_void_p, [PART_P]),
    "pll_gpu_sync_clv": (C.c_int, [PART_P, C.c_uint]),
    "pll_gpu_sync_scaler": (C.c_int, [PART_P, C.c_uint]),
    "pll_gpu_sync_tipchars": (C.c_int, [PART_P, C.c_uint]),
    "pll_gpu_sync_pmatrix": (C.c_int, [PART_P, C.c_uint]),
    "pll_gpu_push_pmatrix": (C.c_int, [PART_P, C.c_uint]),
    "pll_gpu_push_clv": (C.c_int, [PART_P, C.c_uint]),
    "pll_gpu_synchronize": (C.c_int, [PART_P]),
    "pll_gpu_free_sumtable": (C.c_int, [PART_P, C.c_void_p]),
    "pll_gpu_comm_unique_id": (C.c_int, [C.c_char_p]),
    "pll_gpu_comm_init": (C.c_int, [C.c_char_p, C.c_int, C.c_int]),
    "pll_gpu_comm_finalize": (C.c_int, []),
    "pll_gpu_generate_tip_states": (C.c_int, [PART_P, C.c_uint, C.c_ulonglong, C.c_ulonglong]),
    "plg_last_error": (C.c_char_p, []),
    "plg_device_count": (C.c_int, []),
    "plg_timer_start": (C.c_int, [C.c_void_p]),
    "plg_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "plg_get_stats": (C.c_int, [C.c_void_p, C.POINTER(PlgStats)]),
    "plg_reset_stats": (C.c_int, [C.c_void_p]),
    "plg_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "plg_flush_l2": (C.c_int, [C.c_void_p]),
    "plg_mem_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "plg_synchronize": (C.c_int, [C.c_void_p]),
    "plg_set_scaler": (C.c_int, [C.c_void_p, C.c_uint, c_uint_p]),
}


class PllError(RuntimeError):
    pass


def _as_uint(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class PllLibrary:
    """One loaded shared object exporting the pll.h API."""

    def __init__(self, path: str, is_gpu: bool):
        if not os.path.exists(path):
            raise ImportError(f"shared library not found: {path}")
        self.path = path
        self.is_gpu = is_gpu
        self.dll = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        self.missing = []
        for name, (res, args) in _PLL_API.items():
            self._bind(name, res, args)
        if is_gpu:
            for name, (res, args) in _GPU_API.items():
                self._bind(name, res, args)
        self.map_nt = (C.c_uint * 256).in_dll(self.dll, "pll_map_nt")
        self.map_aa = (C.c_uint * 256).in_dll(self.dll, "pll_map_aa")

    def _bind(self, name, res, args):
        try:
            fn = getattr(self.dll, name)
        except AttributeError:
            self.missing.append(name)
            return
        fn.restype = res
        fn.argtypes = args
        setattr(self, name, fn)

    # -- error channel -----------------------------------------------------------------
    def errmsg(self) -> str:
        # pll_errno / pll_errmsg are TLS objects; ctypes resolves TLS symbols through
        # in_dll on the calling thread
        try:
            buf = (C.c_char * 200).in_dll(self.dll, "pll_errmsg")
            return buf.value.decode(errors="replace")
        except Exception:  # pragma: no cover
            return "<pll_errmsg unavailable>"

    def errno(self) -> int:
        try:
            return C.c_int.in_dll(self.dll, "pll_errno").value
        except Exception:  # pragma: no cover
            return -1

    def aa_table(self, name: str, shape) -> np.ndarray:
        n = int(np.prod(shape))
        arr = (C.c_double * n).in_dll(self.dll, name)
        return np.ctypeslib.as_array(arr).reshape(shape).copy()

    def gamma_rates(self, alpha: float, cats: int, mode: int = PLL_GAMMA_RATES_MEAN) -> np.ndarray:
        out = np.zeros(cats, dtype=np.float64)
        ok = self.pll_compute_gamma_cats(alpha, cats, out.ctypes.data_as(c_double_p), mode)
        if not ok:
            raise PllError(self.errmsg())
        return out

    @staticmethod
    def make_map(table: dict):
        """char -> state-set mask table (the `map` argument of pll_set_tip_states) for alphabets
        other than pll_map_nt / pll_map_aa; `table` maps one-character strings to masks."""
        arr = (C.c_uint * 256)()
        for ch, mask in table.items():
            arr[ord(ch)] = int(mask)
        return arr

    def partition(self, **kw) -> "Partition":
        return Partition(self, **kw)


class Partition:
    """A pll_partition_t owned through the C API (pll_partition_create .. _destroy)."""

    def __init__(self, lib: PllLibrary, tips, clv_buffers, states, sites, rate_matrices,
                 prob_matrices, rate_cats, scale_buffers, attributes):
        self.lib = lib
        self.ptr = lib.pll_partition_create(tips, clv_buffers, states, sites, rate_matrices,
                                            prob_matrices, rate_cats, scale_buffers, attributes)
        if not self.ptr:
            raise PllError(f"pll_partition_create failed: {lib.errmsg()} (errno {lib.errno()})")
        self.p = self.ptr.contents
        self._keep = []  # host buffers whose address the library may use as a key

    # -- lifetime ----------------------------------------------------------------------
    def destroy(self):
        if self.ptr:
            self.lib.pll_partition_destroy(self.ptr)
            self.ptr = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.destroy()

    def __del__(self):  # pragma: no cover
        try:
            self.destroy()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != PLL_SUCCESS:
            raise PllError(f"{what} failed: {self.lib.errmsg()} (errno {self.lib.errno()})")

    # -- shape helpers -----------------------------------------------------------------
    @property
    def span(self) -> int:
        return self.p.rate_cats * self.p.states_padded

    @property
    def scaler_len(self) -> int:
        return self.p.sites * (self.p.rate_cats if self.p.attributes & PLL_ATTRIB_RATE_SCALERS else 1)

    # -- setters -----------------------------------------------------------------------
    def set_tip_states(self, tip: int, seq: bytes, amap=None):
        amap = amap if amap is not None else (self.lib.map_nt if self.p.states == 4 else self.lib.map_aa)
        self._check(self.lib.pll_set_tip_states(self.ptr, tip, amap, seq), "pll_set_tip_states")

    def set_tip_clv(self, tip: int, clv: np.ndarray, padding: bool = False):
        clv = _as_f64(clv)
        self._check(self.lib.pll_set_tip_clv(self.ptr, tip, clv.ctypes.data_as(c_double_p), int(padding)),
                    "pll_set_tip_clv")

    def set_pattern_weights(self, w):
        w = _as_uint(w)
        assert w.size == self.p.sites
        self.lib.pll_set_pattern_weights(self.ptr, w.ctypes.data_as(c_uint_p))

    def set_asc_bias_type(self, asc_bias_type: int):
        self._check(self.lib.pll_set_asc_bias_type(self.ptr, asc_bias_type), "pll_set_asc_bias_type")

    def set_asc_state_weights(self, w):
        w = _as_uint(w)
        assert w.size == self.p.states
        self.lib.pll_set_asc_state_weights(self.ptr, w.ctypes.data_as(c_uint_p))

    def set_subst_params(self, idx: int, params):
        a = _as_f64(params)
        self.lib.pll_set_subst_params(self.ptr, idx, a.ctypes.data_as(c_double_p))

    def set_frequencies(self, idx: int, freqs):
        a = _as_f64(freqs)
        self.lib.pll_set_frequencies(self.ptr, idx, a.ctypes.data_as(c_double_p))

    def set_category_rates(self, rates):
        a = _as_f64(rates)
        self.lib.pll_set_category_rates(self.ptr, a.ctypes.data_as(c_double_p))

    def set_category_weights(self, w):
        a = _as_f64(w)
        self.lib.pll_set_category_weights(self.ptr, a.ctypes.data_as(c_double_p))

    def update_invariant_sites(self):
        self._check(self.lib.pll_update_invariant_sites(self.ptr), "pll_update_invariant_sites")

    def update_invariant_sites_proportion(self, idx: int, pinv: float):
        self._check(self.lib.pll_update_invariant_sites_proportion(self.ptr, idx, pinv),
                    "pll_update_invariant_sites_proportion")

    # -- hot path ----------------------------------------------------------------------
    def update_prob_matrices(self, params_indices, matrix_indices, branch_lengths):
        pi, mi, bl = _as_uint(params_indices), _as_uint(matrix_indices), _as_f64(branch_lengths)
        assert mi.size == bl.size and pi.size == self.p.rate_cats
        self._check(
            self.lib.pll_update_prob_matrices(self.ptr, pi.ctypes.data_as(c_uint_p),
                                              mi.ctypes.data_as(c_uint_p),
                                              bl.ctypes.data_as(c_double_p), mi.size),
            "pll_update_prob_matrices")

    def update_partials(self, ops: np.ndarray):
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        self.lib.pll_update_partials(self.ptr, ops.ctypes.data, ops.size)

    def edge_loglikelihood(self, parent_clv, parent_scaler, child_clv, child_scaler, matrix,
                           freqs_indices, persite: Optional[np.ndarray] = None) -> float:
        fi = _as_uint(freqs_indices)
        ps = persite.ctypes.data_as(c_double_p) if persite is not None else None
        return self.lib.pll_compute_edge_loglikelihood(self.ptr, parent_clv, parent_scaler, child_clv,
                                                       child_scaler, matrix,
                                                       fi.ctypes.data_as(c_uint_p), ps)

    def root_loglikelihood(self, clv, scaler, freqs_indices, persite: Optional[np.ndarray] = None) -> float:
        fi = _as_uint(freqs_indices)
        ps = persite.ctypes.data_as(c_double_p) if persite is not None else None
        return self.lib.pll_compute_root_loglikelihood(self.ptr, clv, scaler,
                                                       fi.ctypes.data_as(c_uint_p), ps)

    def new_sumtable(self) -> np.ndarray:
        """Caller-allocated sumtable buffer as in reference examples/newton/newton.c:47-51.
        (Under the GPU backend only its address matters; see include/pll.h.)"""
        # ascertainment-bias storage adds `states` per-state sites (reference src/derivatives.c:56)
        n = (self.p.sites + (self.p.states if self.p.asc_bias_alloc else 0)) * self.span
        raw = np.zeros(n + 8, dtype=np.float64)
        off = (-raw.ctypes.data // 8) % 4  # 32-byte alignment like pll_aligned_alloc
        buf = raw[off:off + n]
        self._keep.append(raw)
        return buf

    def update_sumtable(self, parent_clv, child_clv, parent_scaler, child_scaler, params_indices,
                        sumtable: np.ndarray):
        pi = _as_uint(params_indices)
        self._check(
            self.lib.pll_update_sumtable(self.ptr, parent_clv, child_clv, parent_scaler, child_scaler,
                                         pi.ctypes.data_as(c_uint_p),
                                         sumtable.ctypes.data_as(c_double_p)),
            "pll_update_sumtable")

    def likelihood_derivatives(self, parent_scaler, child_scaler, branch_length, params_indices,
                               sumtable: np.ndarray):
        # called up to 32 times per branch: the ctypes arguments are built once per (indices, table)
        cache = self.__dict__.setdefault("_der_args", {})
        key = (tuple(int(x) for x in params_indices), sumtable.ctypes.data)
        args = cache.get(key)
        if args is None:
            pi = _as_uint(params_indices)
            args = (pi, pi.ctypes.data_as(c_uint_p), sumtable.ctypes.data_as(c_double_p), C.c_double(0), C.c_double(0))
            if len(cache) > 64:
                cache.clear()
            cache[key] = args
        _, pi_p, tab_p, d1, d2 = args
        if not self.lib.pll_compute_likelihood_derivatives(self.ptr, parent_scaler, child_scaler, branch_length,
                                                           pi_p, tab_p, C.byref(d1), C.byref(d2)):
            self._check(0, "pll_compute_likelihood_derivatives")
        return d1.value, d2.value

    # -- reading state back (host arrays in the reference, mirrors under the GPU backend) --
    def get_clv(self, idx: int) -> np.ndarray:
        if self.lib.is_gpu:
            self._check(self.lib.pll_gpu_sync_clv(self.ptr, idx), "pll_gpu_sync_clv")
        n = self.p.sites * self.span
        return np.ctypeslib.as_array(self.p.clv[idx], shape=(n,)).copy().reshape(
            self.p.sites, self.p.rate_cats, self.p.states_padded)

    def get_scaler(self, idx: int) -> np.ndarray:
        if self.lib.is_gpu:
            self._check(self.lib.pll_gpu_sync_scaler(self.ptr, idx), "pll_gpu_sync_scaler")
        return np.ctypeslib.as_array(self.p.scale_buffer[idx], shape=(self.scaler_len,)).copy()

    def get_pmatrix(self, idx: int) -> np.ndarray:
        if self.lib.is_gpu:
            self._check(self.lib.pll_gpu_sync_pmatrix(self.ptr, idx), "pll_gpu_sync_pmatrix")
        n = self.p.rate_cats * self.p.states * self.p.states_padded
        return np.ctypeslib.as_array(self.p.pmatrix[idx], shape=(n,)).copy().reshape(
            self.p.rate_cats, self.p.states, self.p.states_padded)

    def set_pmatrix(self, idx: int, values: np.ndarray):
        """Overwrite a P-matrix set (tests that inject the reference's matrices)."""
        n = self.p.rate_cats * self.p.states * self.p.states_padded
        dst = np.ctypeslib.as_array(self.p.pmatrix[idx], shape=(n,))
        dst[:] = _as_f64(values).reshape(-1)
        if self.lib.is_gpu:
            self._check(self.lib.pll_gpu_push_pmatrix(self.ptr, idx), "pll_gpu_push_pmatrix")

    def get_invariant(self) -> Optional[np.ndarray]:
        if not self.p.invariant:
            return None
        return np.ctypeslib.as_array(self.p.invariant, shape=(self.p.sites,)).copy()

    def get_eigen(self, idx: int):
        K, Kp = self.p.states, self.p.states_padded
        ev = np.ctypeslib.as_array(self.p.eigenvecs[idx], shape=(K * Kp,)).copy().reshape(K, Kp)
        iev = np.ctypeslib.as_array(self.p.inv_eigenvecs[idx], shape=(K * Kp,)).copy().reshape(K, Kp)
        val = np.ctypeslib.as_array(self.p.eigenvals[idx], shape=(Kp,)).copy()
        return ev, iev, val

    def set_eigen(self, idx: int, eigenvecs, inv_eigenvecs, eigenvals):
        """Inject an eigendecomposition (parity tests feed both libraries the same one)."""
        K, Kp = self.p.states, self.p.states_padded
        np.ctypeslib.as_array(self.p.eigenvecs[idx], shape=(K * Kp,))[:] = _as_f64(eigenvecs).reshape(-1)
        np.ctypeslib.as_array(self.p.inv_eigenvecs[idx], shape=(K * Kp,))[:] = _as_f64(inv_eigenvecs).reshape(-1)
        np.ctypeslib.as_array(self.p.eigenvals[idx], shape=(Kp,))[:] = _as_f64(eigenvals).reshape(-1)
        self.p.eigen_decomp_valid[idx] = 1

    # -- GPU-only measurement helpers ------------------------------------------------------
    def ctx(self):
        assert self.lib.is_gpu
        return self.lib.pll_gpu_context(self.ptr)

    def synchronize(self):
        if self.lib.is_gpu:
            self._check(self.lib.pll_gpu_synchronize(self.ptr), "pll_gpu_synchronize")

    def timer_start(self):
        assert self.lib.plg_timer_start(self.ctx()) == 0, self.lib.plg_last_error()

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        assert self.lib.plg_timer_stop(self.ctx(), C.byref(ms)) == 0, self.lib.plg_last_error()
        return float(ms.value)

    def stats(self) -> dict:
        st = PlgStats()
        assert self.lib.plg_get_stats(self.ctx(), C.byref(st)) == 0
        out = {}
        for name, typ in PlgStats._fields_:
            v = getattr(st, name)
            out[name] = int(v) if typ is C.c_ulonglong else [int(x) for x in v]
        return out

    def set_profiling(self, on: bool):
        assert self.lib.plg_set_profiling(self.ctx(), int(on)) == 0

    def reset_stats(self):
        assert self.lib.plg_reset_stats(self.ctx()) == 0

    def flush_l2(self):
        assert self.lib.plg_flush_l2(self.ctx()) == 0, self.lib.plg_last_error()

    def mem_info(self):
        f, t = C.c_size_t(0), C.c_size_t(0)
        assert self.lib.plg_mem_info(self.ctx(), C.byref(f), C.byref(t)) == 0
        return int(f.value), int(t.value)
