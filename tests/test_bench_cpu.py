"""bench.py's contract where no GPU is needed: the reference arm (the oracle port of the reference's PyTorch step on the
host cores) prints ONE JSON line with the agreed keys, and the B200 arm refuses to run without a CUDA device instead of
falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--workload", "cfg1", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train images/sec" and d["unit"] == "images/sec"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["config"]["workload"].startswith("resnet18_attention_unet")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "batch 8" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--workload", "cfg1", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
