"""bench.py contract checks that need no GPU: the reference arm (the CPU oracle port, `--impl reference`) prints one JSON line
with the agreed keys, and the product arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, env=e,
                          timeout=600)


def test_reference_arm_prints_one_json_line():
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm must still use every core it may run on
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--scans-per-map", "2", env={"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "scene_pairs_per_sec" and d["unit"] == "pairs/s" and d["higher_is_better"]
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and "sample" in d["cpu_baseline"]
    assert d["config"]["pairs_per_step"] == 2 and set(d["recall"]) == {"1m_5deg", "0.3m_15deg", "0.6m_1.5deg", "2m_5deg"}
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("configs[1]") and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run("--steps", "1", "--warmup", "3")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
