"""GPU: the reference's OWN known-answer programs run on the device, relink-only.

`make -C oracle dropin` (part of `__graft_entry__.build()` whenever /root/reference is present)
compiles the unmodified reference test sources - test/src/<name>.c with the reference's
test/src/common.c and the reference's own pll.h - and links them against libpll_b200.so instead
of -lpll.  The binaries and the expected text travel to the GPU box in the git-ignored
oracle/_ref/dropin/ (nothing of the reference is copied into the repository):

  oracle/_ref/dropin/<name>                 the program under test
  oracle/_ref/dropin/expected/<name>.out    what the same program prints on the reference
                                            library (AVX2 path); byte-identical to the
                                            reference's fixture test/out/<name>.out where that
                                            exists for the same input (13 programs)
  oracle/_ref/dropin/testdata/              synthetic inputs for the data-driven programs

The programs ask for PLL_ATTRIB_ARCH_CPU/AVX2 (they predate the GPU flag); PLL_GPU_FORCE=1 makes
pll_partition_create replace the architecture bits by PLL_ATTRIB_ARCH_GPU.  Like the reference's
test/runtest.py:265-347 every program runs under several attribute sets and must print the one
fixture; unlike its byte-wise diff, numbers may differ by one unit in the last printed digit
(device expm1/log/exp are 1-2 ulp from glibc's) and values at the cancellation floor of the
derivatives (|x| <= 1e-10, e.g. -6.6613e-15) are compared absolutely.
"""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin")

PROGRAMS = ["00010_NMDU_lkcalc", "00011_NMAU_lkcalc", "00012_NMOU_lkcalc", "00020_NMDR_lkcalc",
            "00021_NMAR_lkcalc", "00022_NMOR_lkcalc", "00030_NMDU_gamma", "00032_NMOU_gamma",
            "alpha-cats", "hky", "pmatrix", "derivatives", "derivatives-oddstates", "protein-models",
            # data-driven programs on synthetic stand-ins for the reference's downloadable test data
            # (oracle/make_testdata.py); expected text = what the reference library prints for them
            "scaling", "asc-bias", "partial-traversal"]
# Rooted-tree callers (pll_rtree_*, libpll_b200/csrc/host/pll_utree.c): green on the device since
# the round-1 driver run, ordinary members of the list now.
PROGRAMS += ["rooted", "rooted-tipinner"]
# scaling.c reads partition->scale_buffer[i] directly (test/src/scaling.c:84-101): it needs the
# host mirrors kept current
EXTRA_ENV = {"scaling": {"PLL_GPU_MIRROR": "1"}}
ATTRIBUTE_SETS = [(), ("tv",), ("tv", "avx2")]          # tokens of reference test/src/common.c:22-56

NUMBER = re.compile(r"^([-+]?)(\d+)\.(\d+)(?:[eE]([-+]?\d+))?$")


def tokens_agree(got: str, want: str) -> bool:
    if got == want:
        return True
    if got.lstrip("+-").rstrip(",") == "nan" and want.lstrip("+-").rstrip(",") == "nan":
        return True      # printf shows the sign bit of a NaN ("-nan"): not a value
    g, w = NUMBER.match(got), NUMBER.match(want)
    if not g or not w:
        # "-0.000000" vs "0.000000" style sign of a rounded zero is the only non-numeric slack
        return False
    a, b = float(got), float(want)
    exponent = int(w.group(4)) if w.group(4) else 0
    last_place = 10.0 ** (exponent - len(w.group(3)))
    # one unit in the last printed digit; where a program prints more digits than double
    # arithmetic in a different summation order can reproduce (examples/newick-phylip-unrooted
    # prints 15 decimals of a log-likelihood of -7.8e5), 1e-12 relative - a hundredth of the
    # parity tolerance; values at the cancellation floor absolutely
    return abs(a - b) <= 1.01 * last_place or abs(a - b) <= 1e-12 * abs(b) or abs(a - b) <= 1e-10


def line_agrees(x: str, y: str) -> bool:
    xt, yt = x.split(), y.split()
    return len(xt) == len(yt) and all(tokens_agree(a, b) for a, b in zip(xt, yt))


def compare_text(got: str, want: str, what: str, alternative: str = None):
    """`alternative`: a second recording of the reference (its plain-C kernels instead of AVX2)
    for inputs on which the reference's own back-ends disagree; a line may follow either."""
    gl, wl = got.splitlines(), want.splitlines()
    al = alternative.splitlines() if alternative else wl
    assert len(gl) == len(wl) == len(al), f"{what}: {len(gl)} lines printed, {len(wl)} expected"
    inexact = 0
    for n, (x, y, z) in enumerate(zip(gl, wl, al), 1):
        if x == y:
            continue
        assert line_agrees(x, y) or line_agrees(x, z), \
            f"{what}:{n}:\n  got : {x}\n  want: {y}" + (f"\n  or  : {z}" if z != y else "")
        inexact += 1
    return inexact, len(wl)


@pytest.mark.parametrize("attrs", ATTRIBUTE_SETS, ids=lambda a: "+".join(a) or "default")
@pytest.mark.parametrize("name", PROGRAMS)
def test_reference_program_prints_its_fixture(name, attrs, capsys):
    exe = os.path.join(DROPIN, name)
    expected = os.path.join(DROPIN, "expected", name + ".out")
    if not (os.path.exists(exe) and os.path.exists(expected)):
        pytest.skip("oracle/_ref/dropin not built (needs /root/reference at build time)")
    env = dict(os.environ, PLL_GPU_FORCE="1", **EXTRA_ENV.get(name, {}))
    limit = 60 if name.startswith("rooted") else 300     # first-run programs must not stall the suite
    out = subprocess.run([exe, *attrs], capture_output=True, text=True, timeout=limit, env=env, cwd=DROPIN)
    if out.stdout == "Skip\n":           # the program itself declines this attribute set
        pytest.skip(f"{name} skips {attrs}")   # (reference test/out/skip.out, runtest.py:339-345)
    assert out.returncode == 0, (out.returncode, out.stderr[-2000:], out.stdout[-2000:])
    plain_c = os.path.join(DROPIN, "expected", name + ".plain-c.out")
    inexact, total = compare_text(out.stdout, open(expected).read(), f"{name} {' '.join(attrs)}",
                                  open(plain_c).read() if os.path.exists(plain_c) else None)
    with capsys.disabled():
        print(f"\n[{name} {' '.join(attrs) or '-'}] {total - inexact}/{total} lines byte-identical, "
              f"{inexact} within one unit in the last printed digit")


EXAMPLES = ["unrooted", "newton", "heterotachy", "lg4", "newick-fasta-unrooted", "newick-phylip-unrooted",
            "protein-list", "rooted", "newick-fasta-rooted"]


@pytest.mark.parametrize("name", EXAMPLES)
def test_reference_example_prints_what_it_prints_on_the_reference(name, capsys):
    """The reference's example programs (examples/<name>/*.c, unmodified, compiled against the
    reference's pll.h; they ask for PLL_ATTRIB_ARCH_AVX / _CPU) relinked against libpll_b200.so:
    examples/unrooted is BASELINE.json configs[0], examples/newton the Newton recipe of configs[3],
    examples/lg4 the mixture path of configs[2].  Inputs: the reference's own lg4 data and the
    synthetic 246 x 4465 alignment as FASTA and as interleaved PHYLIP (with site-pattern
    compression).  Expected text = the same program on the reference library."""
    exe = os.path.join(DROPIN, "example-" + name)
    expected = os.path.join(DROPIN, "expected", f"example-{name}.out")
    if not (os.path.exists(exe) and os.path.exists(expected)):
        pytest.skip("oracle/_ref/dropin not built (needs /root/reference at build time)")
    args = open(os.path.join(DROPIN, "expected", f"example-{name}.args")).read().split()
    env = dict(os.environ, PLL_GPU_FORCE="1")
    limit = 60 if "rooted" in name and "unrooted" not in name else 300
    out = subprocess.run([exe, *args], capture_output=True, text=True, timeout=limit, env=env, cwd=DROPIN)
    assert out.returncode == 0, (out.returncode, out.stderr[-2000:], out.stdout[-2000:])
    inexact, total = compare_text(out.stdout, open(expected).read(), f"example {name}")
    with capsys.disabled():
        print(f"\n[example {name}] {total - inexact}/{total} lines byte-identical, "
              f"{inexact} within tolerance")


def test_without_the_switch_a_cpu_request_is_refused():
    """No CPU path: the same binary without PLL_GPU_FORCE fails in pll_partition_create."""
    exe = os.path.join(DROPIN, "00010_NMDU_lkcalc")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin not built")
    env = {k: v for k, v in os.environ.items() if k != "PLL_GPU_FORCE"}
    out = subprocess.run([exe, "tv", "avx2"], capture_output=True, text=True, timeout=60, env=env, cwd=DROPIN)
    assert out.returncode != 0
    assert "PLL_ATTRIB_ARCH_GPU only" in out.stdout + out.stderr
