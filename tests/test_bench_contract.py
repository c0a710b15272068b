"""bench.py's JSON-line contract, as far as it can be exercised without a GPU: the `--impl reference` arm (the reference's CPU
implementation of the path on the host cores) prints exactly one JSON line with the keys the driver reads. Runs the arm on the
flan-t5-small shape through the test-only B200RANK_BENCH_MODEL override (the real metric is quoted on flan-t5-large)."""
import json
import os
import subprocess
import sys

from helpers import ROOT


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, B200RANK_BENCH_MODEL="flan-t5-small")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "docs/s" and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert d["metric"].startswith("docs scored/sec (flan-t5-small q32/p128")
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0 and d["dtype"] == "f32"
    assert isinstance(d["config"]["workload"], str) and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "documents" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "docs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", B200RANK_BENCH_MODEL="flan-t5-small")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_engine_arm_fails_loudly_without_a_gpu():
    if __import__("conftest").HAS_GPU:
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-text-api"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout)   # no CPU fallback behind the bench either
