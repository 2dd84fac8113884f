"""CPU: host C code under AddressSanitizer + UndefinedBehaviorSanitizer, no GPU needed.

(1) The callers'-side host code (Newick reader and tree utilities, FASTA / PHYLIP readers,
site-pattern compression) under AddressSanitizer + UndefinedBehaviorSanitizer with mutated
inputs (tests/c/fuzz_frontend.c).  These functions take files written by people; the reference
parses them with flex/bison-generated code, ours are hand-written, so they get the hostile-input
treatment: no crash, no out-of-bounds access, no leak on any rejected input, and every accepted
tree passes its integrity check, yields a full operations list and survives export -> re-parse.

(2) The whole pll.h host layer (libpll_b200/csrc/host/*.c) linked against a "null device"
(tests/c/null_device.c) that computes nothing but reads / writes every array over exactly the
extent the plg_* contract of include/pll_gpu.h states: 576 combinations of alphabet size, category
count, tip representation, scaler mode, ascertainment-bias type and pattern slicing
(tests/c/host_scenarios.c) - any disagreement about buffer sizes between the wrappers and the
device ABI, any overflow or leak in the wrappers or their error paths, is a sanitizer report."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "libpll_b200", "csrc", "host")
SOURCES = ["pll_utree.c", "pll_fasta.c", "pll_phylip.c", "pll_compress.c", "pll_maps.c"]


@pytest.fixture(scope="module")
def fuzzer(tmp_path_factory):
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    out = tmp_path_factory.mktemp("fuzz") / "fuzz_frontend"
    cmd = [gcc, "-std=gnu99", "-g", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-fno-omit-frame-pointer", "-I", os.path.join(ROOT, "include"), "-I", HOST, "-o", str(out),
           os.path.join(ROOT, "tests", "c", "fuzz_frontend.c")] + [os.path.join(HOST, s) for s in SOURCES] + ["-lm"]
    built = subprocess.run(cmd, capture_output=True, text=True)
    if built.returncode != 0 and "sanitize" in built.stderr:
        pytest.skip("this gcc has no sanitizer runtime")
    assert built.returncode == 0, built.stderr
    return str(out)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_front_end_survives_mutated_inputs(fuzzer, tmp_path, seed):
    run = subprocess.run([fuzzer, "20000", str(seed), str(tmp_path)], capture_output=True, text=True,
                         timeout=300, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0"))
    assert run.returncode == 0, (run.stdout[-500:], run.stderr[-3000:])
    m = re.search(r"trees=(\d+) fasta_records=(\d+) alignments=(\d+) compressed=(\d+) rooted=(\d+)", run.stdout)
    assert m, run.stdout
    trees, records, alignments, compressed, rooted = map(int, m.groups())
    # the mutations must leave enough valid inputs to exercise the accepting paths too
    assert trees > 1000 and records > 2000 and alignments > 300 and compressed > 1000 and rooted > 500


def _build(tmp_path_factory, name, sources):
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    out = tmp_path_factory.mktemp(name) / name
    cmd = [gcc, "-std=gnu99", "-g", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-fno-omit-frame-pointer", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", HOST,
           "-o", str(out)] + sources + ["-lm", "-pthread"]
    built = subprocess.run(cmd, capture_output=True, text=True)
    if built.returncode != 0 and "sanitize" in built.stderr:
        pytest.skip("this gcc has no sanitizer runtime")
    assert built.returncode == 0, built.stderr
    return str(out)


def test_host_layer_against_the_null_device(tmp_path_factory):
    import glob
    exe = _build(tmp_path_factory, "host_scenarios",
                 [os.path.join(ROOT, "tests", "c", "host_scenarios.c"), os.path.join(ROOT, "tests", "c", "null_device.c")]
                 + sorted(glob.glob(os.path.join(HOST, "*.c"))))
    for mirror in ("0", "1"):      # PLL_GPU_MIRROR=1: host mirrors refreshed by every update / tip upload
        run = subprocess.run([exe], capture_output=True, text=True, timeout=300,
                             env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0",
                                      PLL_GPU_MIRROR=mirror))
        assert run.returncode == 0, run.stderr[-4000:]
        m = re.search(r"scenarios=(\d+) refused=(\d+)", run.stderr)
        assert m and int(m.group(1)) == 576 and int(m.group(2)) == 0, run.stderr[-500:]
