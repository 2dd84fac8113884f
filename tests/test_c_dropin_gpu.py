"""GPU: a plain C program written against include/pll.h, linked with -lpll_b200, reproduces the
reference's examples/unrooted and examples/newton outputs (drop-in check of the C boundary)."""
import json
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


def _build(tmp_path, name):
    exe = str(tmp_path / name)
    libdir = os.path.join(ROOT, "libpll_b200")
    subprocess.run(["gcc", "-std=gnu99", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", name + ".c"), "-o", exe, "-L", libdir,
                    "-lpll_b200", "-lm", f"-Wl,-rpath,{libdir}"], check=True)
    return exe


@pytest.mark.parametrize("fmt", ["fas", "phy"])
def test_c_program_runs_the_lg4_example_from_files(tmp_path, fmt):
    """tests/c/lg4_gpu.c: Newick + FASTA / PHYLIP files -> device pattern compression -> plain and
    recycled operation lists -> LG4M / LG4X log-likelihoods, against the values recorded from the
    reference (tests/golden/lg4_example.json; compression does not change a log-likelihood)."""
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "lg4_example.json")))
    (tmp_path / "example.tree").write_text(gold["newick"])
    if fmt == "fas":
        (tmp_path / "example.fas").write_text(gold["fasta_text"])
    else:
        recs = gold["fasta"]
        (tmp_path / "example.phy").write_text(
            f"{len(recs)} {len(recs[0][1])}\n" + "".join(f"{h} {s}\n" for h, s, _ in recs))
    exe = _build(tmp_path, "lg4_gpu")
    out = subprocess.run([exe, str(tmp_path / "example.tree"), str(tmp_path / f"example.{fmt}")],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr + out.stdout
    assert "21 taxa, 113 sites" in out.stdout
    m = re.findall(r"\[(plain|recycled), (\d+) CLV buffers\] Log-L \((LG4M|LG4X)\): (-?\d+\.\d+)", out.stdout)
    assert len(m) == 4, out.stdout
    want = {"LG4M": gold["expect"]["tv"]["lg4m"], "LG4X": gold["expect"]["tv"]["lg4x"]}
    for kind, buffers, model, value in m:
        assert abs(float(value) - want[model]) < 2e-6, out.stdout
        assert int(buffers) == (19 if kind == "plain" else 3)
