"""bench.py's reference arm (the CPU oracle dataflow) runs without a GPU: its JSON line is checked
against the driver's contract here, and the GPU arm is checked to refuse to run without CUDA."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    return subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, env=env
    )


def test_reference_arm_prints_one_contract_line():
    p = run_bench("--impl", "reference", "--sf", "1", "--steps", "3", "--warmup", "3")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1  # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "update_rows_per_sec" and d["unit"] == "rows/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = run_bench("--impl", "reference", "--gpus", "2", "--sf", "1", "--steps", "3", "--warmup", "3", env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_cuda():
    """The product path has no CPU fallback: without a GPU the bench must fail, not time the oracle."""
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present")
    p = run_bench("--steps", "3", "--warmup", "3", "--sf", "1")
    assert p.returncode != 0
    assert '"metric"' not in p.stdout
