"""bench.py contract checks that need no GPU: the reference arm (CPU legs) prints ONE JSON line with the
keys the driver reads, and ranks other than 0 exit quietly under torchrun-style environments."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def _check_reference_line(r, batch, kind):
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "512x512 pairs/sec (cost volume+warp)" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] / 1000.0 - batch) < 1e-6      # one batch per step
    cb = d["cpu_baseline"]
    assert cb["kind"] == kind and cb["cores"] == os.cpu_count() and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0
    # the reference arm describes the SAME workload with the SAME config keys as our arm
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == json.loads(json.dumps(bench.workload_config(1, batch)))
    return d


def test_reference_arm_line_runs_the_reference_itself():
    """With the reference copy present (baseline/_ref, made by build() where /root/reference exists) the CPU arm
    times the reference's own functions: kind == "reference", torch's parallel info in the line."""
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "core")):
        import pytest
        pytest.skip("baseline/_ref absent (build() has not run where the reference tree exists)")
    r = _run(None, "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0", "--batch", "2")
    d = _check_reference_line(r, 2, "reference")
    assert d["cpu_baseline"]["torch_parallel_info"] and d["cpu_baseline"]["oracle_port"]["kind"] == "port"


def test_reference_arm_line_falls_back_to_the_port():
    r = _run({"STITCH_REF_COPY": "/nonexistent"}, "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0",
             "--batch", "2")
    _check_reference_line(r, 2, "port")


def test_reference_arm_other_ranks_do_no_work():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2", "--steps", "1",
             "--warmup", "0")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: on a machine without a B200 the product arm must not print a number."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run(None, "--gpus", "1", "--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
