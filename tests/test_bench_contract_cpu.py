"""CPU: the reference arm of bench.py (`--impl reference`: the reference's own AVX2 path from
oracle/_ref on the host cores) prints ONE JSON line with the keys the benchmark contract names,
and the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line(ref_lib):
    run = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert run.returncode == 0, run.stderr[-2000:]
    lines = [l for l in run.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, run.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    assert line["metric"].startswith("CLV site-updates/sec") and line["unit"] == "site-updates/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["vs_baseline"] is None
    assert "workload" in line["config"] and "1000-taxon x 1000000-pattern" in line["config"]["workload"]
    assert line["value"] > 1e6 and line["ms_per_step"] > 0
    base = line["cpu_baseline"]
    assert base["kind"] in ("reference", "port") and base["cores"] >= 1 and base["value"] == line["value"]
    assert "sample" in base
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}


def test_gpu_arm_refuses_to_run_without_a_device(has_gpu):
    if has_gpu:
        pytest.skip("a GPU is present")
    run = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert run.returncode != 0
    assert "no CPU fallback" in run.stderr + run.stdout
    assert not [l for l in run.stdout.splitlines() if l.startswith("{")]
