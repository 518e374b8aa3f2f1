"""bench.py contract checks that need no GPU: the reference arm's JSON line, and that the product arm
fails loudly (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=300,
                          env=dict(os.environ, **(env or {})))


def test_reference_arm_json_line():
    r = run("--impl", "reference", "--rows", "300000", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "rows/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["n_gpus"] == 1 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and "C2" in line["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    r = run("--impl", "reference", "--gpus", "2", "--rows", "100000", "--steps", "1", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_product_arm_has_no_cpu_fallback():
    r = run("--rows", "100000", "--steps", "1")
    assert r.returncode != 0
    assert "no CUDA device available" in r.stderr and "no CPU fallback" in r.stderr
    assert "{" not in r.stdout  # no bench line is printed
