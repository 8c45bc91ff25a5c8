"""bench.py without a GPU: the reference arm (the oracle port on the host cores) prints the contract's
JSON line, and the product arm refuses to run - there is no CPU fallback to measure."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _run(*flags, env=None):
    import os
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *flags], cwd=ROOT, capture_output=True, text=True,
                          timeout=300, env={**os.environ, **(env or {})})


def test_reference_arm_prints_one_contract_line():
    done = _run("--impl", "reference", "--pairs", "3000", "--steps", "2", "--warmup", "1", "--cpu-seconds", "0.5")
    assert done.returncode == 0, done.stderr[-2000:]
    lines = [x for x in done.stdout.splitlines() if x.strip()]
    assert len(lines) == 1                      # one JSON line on stdout, everything else on stderr
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["unit"] == "GCUPS" and line["higher_is_better"] is True and line["value"] > 0
    assert line["metric"].startswith("GCUPS") and line["config"]["workload"].startswith("C2")
    base = line["cpu_baseline"]
    assert base["kind"] == "port" and base["cores"] >= 1 and base["value"] == line["value"] and "3000" in base["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None


def test_reference_arm_runs_on_rank_0_only():
    done = _run("--impl", "reference", "--gpus", "2", "--pairs", "3000", "--steps", "1", "--warmup", "0",
                env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert done.returncode == 0 and done.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the product arm runs")
    done = _run("--pairs", "1000", "--steps", "1", "--warmup", "0", "--configs", "none", "--no-cpu-baseline")
    assert done.returncode != 0 and done.stdout.strip() == ""
    assert "no CPU fallback" in done.stderr
