"""GPU: a plain C program written against include/pll.h, linked with -lpll_b200, reproduces the
reference's examples/unrooted and examples/newton outputs (drop-in check of the C boundary)."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_program_links_and_matches_reference_values(tmp_path):
    exe = str(tmp_path / "unrooted_gpu")
    libdir = os.path.join(ROOT, "libpll_b200")
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "unrooted_gpu.c"), "-o", exe, "-L", libdir,
                    "-lpll_b200", "-lm", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    vals = [float(x) for x in re.findall(r"Log-L[^:]*: (-?\d+\.\d+)", out.stdout)]
    assert len(vals) == 3
    for got, want in zip(vals, (-33.387713, -34.550204, -36.830297)):
        assert abs(got - want) < 5e-7, out.stdout
    m = re.search(r"Newton: (\d+\.\d+) after (\d+) iterations", out.stdout)
    assert m and abs(float(m.group(1)) - 2.607098) < 5e-7 and int(m.group(2)) == 7, out.stdout
    assert "CLV 4, site 0" in out.stdout
