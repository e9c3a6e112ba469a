"""bench.py's contract lines that can be produced without a GPU: the reference arm prints ONE JSON line with the keys the driver
reads, and the product arm refuses to run without a CUDA device (there is no CPU fallback to time by accident)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=600)


@pytest.mark.parametrize("extra", [(), ("--aruco3", "0.02")])
def test_reference_arm_prints_one_contract_line(extra):
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--batch", "2", *extra)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"].startswith("frames/sec detect+pose") and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0
    assert "C2" in d["config"]["workload"] and ("ArUco3" in d["config"]["workload"]) == bool(extra)
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "1", "--batch", "2", "--no-cpu-baseline")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
