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
    lines = _run(["--codec", codec], {"LD_DEBUG": "files", "LD_DEBUG_OUTPUT": "/dev/null"})
    assert len(lines) == 1                                       # ONE JSON line on stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "GB/s" and j["higher_is_better"] is True
    assert j["n_gpus"] == 1 and j["steps"] == 1 and j["warmup"] == 1 and j["value"] > 0 and j["ms_per_step"] > 0
    assert j["dtype"] == "u8" and j["data"] == "synthetic" and j["vs_baseline"] is None and j["scaling"] == "strong"
    assert "workload" in j["config"] and "model" not in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["gpu_launches"] == 0
    assert ("4mz" in j["metric"]) == (codec == "4mz")


@pytest.mark.parametrize("config", [3, 4])
def test_reference_arm_other_configs(config):
    """configs[3] times the reference's LZ4_compress_HC level 4, configs[4] the reader's half (XXH32 + LZ4_decompress_safe);
    both arms print the same `config` object for a workload (the driver compares them)."""
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not built")
    lines = _run(["--config", str(config)])
    j = json.loads(lines[0])
    assert f"configs[{config}]" in j["config"]["workload"] and j["value"] > 0
    assert ("LZ4_compress_HC level 4" in j["cpu_baseline"]["sample"]) == (config == 3)
    sys.path.insert(0, ROOT)
    import bench
    assert j["config"] == bench.config_dict(bench.CONFIGS[config])


def test_reference_arm_never_maps_the_product_library():
    """The reference arm measures the reference: the CUDA library of this repo is not even loaded (its input comes from
    the host-only generator 4mc_b200/host/libfourmcgen.so)."""
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not built")
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu-gib', '0.0625'];"
            "runpy.run_path(%r, run_name='__main__'); print('MAPS', 'lib4mcgpu' in open('/proc/self/maps').read())" % os.path.join(ROOT, "bench.py"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "MAPS False" in p.stdout


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
