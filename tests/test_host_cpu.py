"""CPU tests of the host C layer and of the C-ABI surface (no compute calls: no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import libpll_b200
from libpll_b200.binding import (PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_ARCH_GPU, PLL_ATTRIB_PATTERN_TIP,
                                 PllError, PllPartition, c_double_p)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for hdr in ("pll.h", "pll_gpu.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"#define PLL_EXPORT.*", "", text)
        for m in re.finditer(r"PLL_EXPORT\s+[^;(]*?\b(\w+)\s*\(", text):
            names.add(m.group(1))
        for m in re.finditer(r"PLL_EXPORT\s+extern\s+[^;]*?\b(\w+)\s*(\[[^;]*)?;", text):
            names.add(m.group(1))
    return names


def test_library_loads_and_exports_every_declared_symbol(gpu_lib):
    declared = _declared_symbols()
    assert len(declared) > 60
    out = subprocess.run(["nm", "-D", "--defined-only", libpll_b200.LIB_PATH], capture_output=True, text=True,
                         check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = sorted(declared - exported)
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    assert not gpu_lib.missing


def test_only_sm100a_code_in_the_library():
    out = subprocess.run(["cuobjdump", "-lelf", libpll_b200.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_struct_layout_matches_reference_abi():
    assert C.sizeof(PllPartition) == 216
    assert PllPartition.states_padded.offset == 48 and PllPartition.clv.offset == 56
    assert PllPartition.maxstates.offset == 168 and PllPartition.asc_bias_alloc.offset == 208


def test_create_without_gpu_flag_is_refused(gpu_lib):
    with pytest.raises(PllError) as e:
        gpu_lib.partition(tips=4, clv_buffers=2, states=4, sites=8, rate_matrices=1, prob_matrices=5,
                          rate_cats=4, scale_buffers=2, attributes=PLL_ATTRIB_ARCH_AVX2)
    assert "PLL_ATTRIB_ARCH_GPU" in str(e.value)
    with pytest.raises(PllError):
        gpu_lib.partition(tips=4, clv_buffers=2, states=4, sites=8, rate_matrices=1, prob_matrices=5,
                          rate_cats=4, scale_buffers=2, attributes=PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_ARCH_AVX2)


def test_no_cpu_fallback_without_a_device(gpu_lib, has_gpu):
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(PllError) as e:
        gpu_lib.partition(tips=4, clv_buffers=2, states=4, sites=8, rate_matrices=1, prob_matrices=5,
                          rate_cats=4, scale_buffers=2, attributes=PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP)
    assert "no CUDA device" in str(e.value) and gpu_lib.errno() == 131


def test_force_switch_replaces_the_architecture_bits(gpu_lib, has_gpu, monkeypatch):
    """PLL_GPU_FORCE=1 (relink-only drop-in): a request for the reference's AVX2 kernels becomes a
    GPU request - here, without a device, it now fails for the lack of one (131), not as an
    unsupported architecture (113).  There is still no CPU path."""
    if has_gpu:
        pytest.skip("a GPU is present (covered by tests/test_reference_programs_gpu.py)")
    kw = dict(tips=4, clv_buffers=2, states=4, sites=8, rate_matrices=1, prob_matrices=5, rate_cats=4,
              scale_buffers=2, attributes=PLL_ATTRIB_ARCH_AVX2 | PLL_ATTRIB_PATTERN_TIP)
    with pytest.raises(PllError):
        gpu_lib.partition(**kw)
    assert gpu_lib.errno() == 113
    monkeypatch.setenv("PLL_GPU_FORCE", "1")
    with pytest.raises(PllError) as e:
        gpu_lib.partition(**kw)
    assert "no CUDA device" in str(e.value) and gpu_lib.errno() == 131
    monkeypatch.setenv("PLL_GPU_FORCE", "0")
    with pytest.raises(PllError):
        gpu_lib.partition(**kw)
    assert gpu_lib.errno() == 113


@pytest.mark.parametrize("sites", [1, 63, 64, 65, 400, 777, 1000, 5000, 1_000_000, 10_000_000, 4_294_967_295])
@pytest.mark.parametrize("slices", [1, 2, 3, 5, 8, 16])
def test_slice_bounds(gpu_lib, sites, slices):
    """The pattern-slicing rule of multi-device partitions (pll_gpu_slice_bounds): contiguous,
    64-aligned, covering, no empty slice, balanced to within one 64-pattern granule."""
    lo = np.zeros(slices + 1, dtype=np.uint32)
    n = gpu_lib.pll_gpu_slice_bounds(sites, slices, lo.ctypes.data_as(C.POINTER(C.c_uint)))
    assert 1 <= n <= slices
    b = lo[:n + 1].astype(np.int64)
    assert b[0] == 0 and b[-1] == sites
    assert np.all(np.diff(b) > 0)
    assert np.all(b[:-1] % 64 == 0)
    sizes = np.diff(b)
    assert sizes.max() - (-(-sites // slices)) < 64
    assert np.all(sizes[:-1] == sizes[0])          # only the last slice is short
    if sites >= 64 * slices * slices:
        assert n == slices


def test_maps_match_reference(gpu_lib, ref_lib):
    for name in ("pll_map_nt", "pll_map_aa", "pll_map_bin"):
        a = np.array((C.c_uint * 256).in_dll(gpu_lib.dll, name))
        b = np.array((C.c_uint * 256).in_dll(ref_lib.dll, name))
        assert np.array_equal(a, b), name


def test_aa_model_tables_match_reference(gpu_lib, ref_lib):
    single = ("dayhoff", "lg", "dcmut", "jtt", "mtrev", "wag", "rtrev", "cprev", "vt", "blosum62", "mtmam",
              "mtart", "mtzoa", "pmb", "hivb", "hivw", "jttdcmut", "flu", "stmtrev")   # src/pll.h:480-522
    tables = [(f"pll_aa_rates_{m}", (190,)) for m in single] + [(f"pll_aa_freqs_{m}", (20,)) for m in single]
    tables += [("pll_aa_rates_lg4m", (4, 190)), ("pll_aa_freqs_lg4m", (4, 20)),
               ("pll_aa_rates_lg4x", (4, 190)), ("pll_aa_freqs_lg4x", (4, 20))]
    for name, shape in tables:
        assert gpu_lib.aa_table(name, shape).tobytes() == ref_lib.aa_table(name, shape).tobytes(), name


def test_gamma_rates_bit_identical_to_reference(gpu_lib, ref_lib):
    for alpha in (0.02, 0.05, 0.1, 0.2, 0.5, 0.75, 1.0, 1.5, 2.0, 10.0, 99.0):
        for cats in (1, 2, 3, 4, 5, 8, 16):
            for mode in (0, 1):
                a, b = gpu_lib.gamma_rates(alpha, cats, mode), ref_lib.gamma_rates(alpha, cats, mode)
                assert a.tobytes() == b.tobytes(), (alpha, cats, mode)
    out = np.zeros(4)
    assert gpu_lib.pll_compute_gamma_cats(0.0, 4, out.ctypes.data_as(c_double_p), 0) == 0
    assert gpu_lib.errno() == 113  # PLL_ERROR_PARAM_INVALID, as reference test 00010 expects


def test_gamma_rates_close_to_exact_discretisation(gpu_lib):
    from oracle import port

    for alpha in (0.3, 0.5, 1.0, 4.0):
        np.testing.assert_allclose(gpu_lib.gamma_rates(alpha, 4), port.gamma_mean_rates(alpha, 4), rtol=2e-5)


def _fake_partition(lib_like, K, subst, freqs):
    """A hand-built pll_partition_t carrying only what pll_update_eigen touches, so that the
    host eigensolver can be exercised without a device (reference src/models.c:251-331)."""
    p = PllPartition()
    p.states = K
    p.states_padded = K
    p.rate_matrices = 1
    keep = {}

    def arr(n, fill=None):
        raw = np.zeros(n + 8)
        off = (-raw.ctypes.data // 8) % 4
        a = raw[off:off + n]
        if fill is not None:
            a[:] = fill
        keep[id(a)] = raw
        return a

    ev, iev, val = arr(K * K), arr(K * K), arr(K)
    fr, sp = arr(K, freqs), arr(K * (K - 1) // 2, subst)
    valid = np.zeros(1, dtype=np.int32)
    ptrs = {}
    for name, a in (("eigenvecs", ev), ("inv_eigenvecs", iev), ("eigenvals", val), ("frequencies", fr),
                    ("subst_params", sp)):
        ptrs[name] = (c_double_p * 1)(a.ctypes.data_as(c_double_p))
        setattr(p, name, ptrs[name])
    p.eigen_decomp_valid = valid.ctypes.data_as(C.POINTER(C.c_int))
    return p, (ev, iev, val), (keep, ptrs, valid, fr, sp)


@pytest.mark.parametrize("K", [4, 20])
def test_eigendecomposition_bit_identical_to_reference(gpu_lib, ref_lib, K):
    rng = np.random.default_rng(K)
    for trial in range(5):
        subst = rng.uniform(0.1, 5.0, K * (K - 1) // 2)
        freqs = rng.uniform(0.5, 1.5, K)
        freqs /= freqs.sum()
        if K == 20 and trial == 0:
            subst, freqs = ref_lib.aa_table("pll_aa_rates_lg", (190,)), ref_lib.aa_table("pll_aa_freqs_lg", (20,))
        pa, outa, keepa = _fake_partition(gpu_lib, K, subst, freqs)
        pb, outb, keepb = _fake_partition(ref_lib, K, subst, freqs)
        assert gpu_lib.pll_update_eigen(C.byref(pa), 0) == 1
        assert ref_lib.pll_update_eigen(C.byref(pb), 0) == 1
        for x, y, name in zip(outa, outb, ("eigenvecs", "inv_eigenvecs", "eigenvals")):
            assert x.tobytes() == y.tobytes(), (name, float(np.max(np.abs(x - y))))
        # and it really is a decomposition: V^-1 diag(l) V rows sum to zero (rate matrix)
        V, iV, lam = outa[0].reshape(K, K), outa[1].reshape(K, K), outa[2]
        Q = iV @ np.diag(lam) @ V
        np.testing.assert_allclose(Q.sum(axis=1), 0, atol=1e-12)
        np.testing.assert_allclose(iV @ V, np.eye(K), atol=1e-12)


def test_product_does_not_depend_on_the_oracle():
    """No file of the product tree may mention the oracle (parity would be void)."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "libpll_b200")):
        if "build" in base or "__pycache__" in base:
            continue
        for f in files:
            if f.endswith((".c", ".h", ".cu", ".cuh", ".py", "Makefile")):
                text = open(os.path.join(base, f), errors="replace").read()
                if re.search(r"(from|import)\s+oracle|oracle/|libpll_oracle|libpll_ref", text) and f != "binding.py":
                    bad.append(os.path.join(base, f))
    assert not bad, bad
    out = subprocess.run(["ldd", libpll_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libpll_ref" not in out


def _compress(lib, seqs, amap):
    bufs = [C.create_string_buffer(s) for s in seqs]
    arr = (C.c_char_p * len(bufs))(*[C.cast(b, C.c_char_p) for b in bufs])
    n = C.c_int(len(seqs[0]))
    w = lib.pll_compress_site_patterns(arr, amap, len(seqs), C.byref(n))
    return [b.value for b in bufs], np.ctypeslib.as_array(w, shape=(n.value,)).copy()


@pytest.mark.parametrize("alphabet", ["nt", "aa"])
def test_pattern_compression_bit_exact(gpu_lib, ref_lib, alphabet):
    """pll_compress_site_patterns: sorted unique columns and weights identical to the reference
    (reference src/compress.c:138-286), including the byte-range remap for amino-acid masks and
    the last-character-wins decode."""
    rng = np.random.default_rng(5)
    chars = b"ACGTacgtNRYKM-?" if alphabet == "nt" else b"ARNDCQEGHILKMFPSTWYVBZX-*arnd"
    for trial in range(25):
        T, S = int(rng.integers(1, 14)), int(rng.integers(1, 600))
        base = bytes(rng.choice(list(chars), S).tolist())
        seqs = []
        for _ in range(T):
            mut = rng.random(S) < 0.15
            alt = rng.choice(list(chars), S)
            seqs.append(bytes(np.where(mut, alt, np.frombuffer(base, np.uint8)).astype(np.uint8).tolist()))
        ga = _compress(gpu_lib, seqs, gpu_lib.map_nt if alphabet == "nt" else gpu_lib.map_aa)
        rb = _compress(ref_lib, seqs, ref_lib.map_nt if alphabet == "nt" else ref_lib.map_aa)
        assert ga[0] == rb[0], trial
        assert np.array_equal(ga[1], rb[1]), trial
        assert int(ga[1].sum()) == S
