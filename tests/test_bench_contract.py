"""CPU tests of bench.py's contract: the reference arm (the CPU restatement timed with every host thread) prints ONE JSON line with
the keys the driver reads, under a single process and as rank 0 / rank 1 of a 2-rank launch; the GPU arm refuses to run without a
device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, cwd=ROOT)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "iterations/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("ICP iterations/sec") and d["config"]["workload"].startswith("config2")
    assert d["value"] > 0 and abs(d["value"] - 30.0 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "30 iterations" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = _run(["--steps", "1", "--warmup", "3"])
    assert r.returncode != 0
    assert r.stdout.strip() == ""
