"""GPU: site-pattern compression on the device (libpll_b200/csrc/gpu/plg_compress.cu)
against the reference's pll_compress_site_patterns (oracle/_ref, src/compress.c:138-286): the
compressed sequences (unique columns in the reference's sorted order, decoded with its
last-character-wins inverse map) and the weights must be identical, byte for byte."""
import ctypes as C
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compress(fn, seqs, amap):
    bufs = [C.create_string_buffer(s) for s in seqs]
    arr = (C.c_char_p * len(bufs))(*[C.cast(b, C.c_char_p) for b in bufs])
    n = C.c_int(len(seqs[0]))
    t0 = time.perf_counter()
    w = fn(arr, amap, len(seqs), C.byref(n))
    dt = time.perf_counter() - t0
    assert w, "compression failed"
    return [b.value for b in bufs], np.ctypeslib.as_array(w, shape=(n.value,)).copy(), dt


def _alignment(rng, chars, T, S, mutation=0.15):
    base = rng.choice(np.frombuffer(chars, np.uint8), S)
    seqs = []
    for _ in range(T):
        mut = rng.random(S) < mutation
        alt = rng.choice(np.frombuffer(chars, np.uint8), S)
        seqs.append(np.where(mut, alt, base).astype(np.uint8).tobytes())
    return seqs


@pytest.mark.parametrize("alphabet", ["nt", "aa"])
def test_device_compression_matches_reference(gpu_lib, ref_lib, alphabet):
    rng = np.random.default_rng(11)
    chars = b"ACGTacgtNRYKM-?" if alphabet == "nt" else b"ARNDCQEGHILKMFPSTWYVBZX-*arnd"
    gmap = gpu_lib.map_nt if alphabet == "nt" else gpu_lib.map_aa
    rmap = ref_lib.map_nt if alphabet == "nt" else ref_lib.map_aa
    for T, S in [(1, 1), (1, 700), (2, 5), (7, 333), (8, 1000), (9, 1001), (16, 4096), (33, 2500), (100, 20000)]:
        seqs = _alignment(rng, chars, T, S)
        g = _compress(gpu_lib.pll_gpu_compress_site_patterns, seqs, gmap)
        r = _compress(ref_lib.pll_compress_site_patterns, seqs, rmap)
        assert g[0] == r[0], (T, S)
        assert np.array_equal(g[1], r[1]), (T, S)
        assert int(g[1].sum()) == S


def test_device_compression_low_diversity_and_illegal_characters(gpu_lib, ref_lib):
    """few distinct columns (long runs of equal keys) and characters the map does not know:
    those encode as 0, which ENDS the column for the reference's string comparison"""
    rng = np.random.default_rng(3)
    seqs = _alignment(rng, b"AC", 12, 30000, mutation=0.01)
    g = _compress(gpu_lib.pll_gpu_compress_site_patterns, seqs, gpu_lib.map_nt)
    r = _compress(ref_lib.pll_compress_site_patterns, seqs, ref_lib.map_nt)
    assert g[0] == r[0] and np.array_equal(g[1], r[1])
    seqs = _alignment(rng, b"ACGT!#", 11, 5000, mutation=0.3)
    g = _compress(gpu_lib.pll_gpu_compress_site_patterns, seqs, gpu_lib.map_nt)
    r = _compress(ref_lib.pll_compress_site_patterns, seqs, ref_lib.map_nt)
    assert np.array_equal(g[1], r[1])
    # an illegal character decodes through inv_charmap[0], which the reference leaves
    # uninitialised (src/compress.c:173-175): only the weights - the partition of the columns and
    # its order - are comparable


def test_device_compression_large(gpu_lib, ref_lib, capsys):
    """200 taxa x 1 M columns: identical output; prints both timings"""
    rng = np.random.default_rng(1)
    seqs = _alignment(rng, b"ACGT-", 200, 1_000_000, mutation=0.02)
    g = _compress(gpu_lib.pll_gpu_compress_site_patterns, seqs, gpu_lib.map_nt)
    r = _compress(ref_lib.pll_compress_site_patterns, seqs, ref_lib.map_nt)
    assert g[0] == r[0] and np.array_equal(g[1], r[1])
    with capsys.disabled():
        print(f"\n[compress 200 x 1M -> {len(g[1])} patterns] device {g[2]*1e3:.0f} ms, reference (1 core) {r[2]*1e3:.0f} ms")
