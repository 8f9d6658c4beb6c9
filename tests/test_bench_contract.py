"""The bench.py contract that can be checked without a GPU: the reference arm (`--impl reference`, the reference's own
shaders compiled for the CPU, or the oracle port) prints exactly ONE JSON line on stdout with the keys the driver reads,
and the product arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _run(args, env=None):
    return subprocess.run([sys.executable, str(ROOT / "bench.py")] + args, capture_output=True, text=True, timeout=600,
                          env={**os.environ, **(env or {})})


def test_reference_arm_prints_one_json_line():
    # C1 (640x360, one quad light) keeps the CPU work to a second or two
    p = _run(["--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "Gsamples/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["steps"] == 1 and j["n_gpus"] == 1 and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["value"] == j["value"] == j["e2e"]["value"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in j["config"]


def test_reference_arm_other_ranks_exit_quietly():
    p = _run(["--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = _run(["--steps", "1", "--warmup", "1"])
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout)
    assert p.stdout.strip() == ""      # nothing that could be mistaken for a bench line
