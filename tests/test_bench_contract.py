"""bench.py's reference arm runs without a GPU: its JSON line must keep the contract's keys (the driver parses it),
and under a 2-process launch only rank 0 prints."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libref4mc.so")


def _run(extra, env=None):
    e = dict(os.environ, **(env or {}))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-gib", "0.0625"] + extra, capture_output=True, text=True, env=e, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return [ln for ln in p.stdout.splitlines() if ln.strip()]


@pytest.mark.parametrize("codec", ["4mc", "4mz"])
def test_reference_arm_line(codec):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not built")
    lines = _run(["--codec", codec])
    assert len(lines) == 1                                       # ONE JSON line on stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "GB/s" and j["higher_is_better"] is True
    assert j["n_gpus"] == 1 and j["steps"] == 1 and j["warmup"] == 1 and j["value"] > 0 and j["ms_per_step"] > 0
    assert j["dtype"] == "u8" and j["data"] == "synthetic" and j["vs_baseline"] is None and j["scaling"] == "weak"
    assert "workload" in j["config"] and "model" not in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["gpu_launches"] == 0
    assert ("4mz" in j["metric"]) == (codec == "4mz")


def test_reference_arm_only_rank_zero_prints():
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not built")
    assert _run(["--gpus", "2"], {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []
    lines = _run(["--gpus", "2"], {"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "2"})
    assert len(lines) == 1 and json.loads(lines[0])["n_gpus"] == 2


def test_our_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm exits non-zero and prints no result line."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu", "--no-e2e"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode != 0
    assert p.stdout.strip() == ""
