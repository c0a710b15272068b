"""bench.py's JSON-line contract, as far as it can be exercised without a GPU: the `--impl reference` arm (the reference's CPU
implementation of the path on the host cores) prints exactly one JSON line with the keys the driver reads. Runs the arm on the
flan-t5-small shape through the test-only B200RANK_BENCH_MODEL override (the real metric is quoted on flan-t5-large)."""
import json
import os
import subprocess
import sys

import pytest

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


@pytest.mark.parametrize("qps", [1, 2])
def test_engine_arm_json_assembly_with_a_stand_in_engine(monkeypatch, capsys, qps):
    """run_engine end to end on CPU: b200rank.Engine is replaced by a stand-in that answers with the transformers fp32 forward (+ noise of
    bf16 size) and reports a canned per-kernel profile, so every leg that shapes the JSON line runs — value / e2e / launches / roofline
    (dominant kernel, traffic lookup, HBM entry) / cpu_baseline / parity (incl. the bf16 yardstick and Kendall tau) / text API. This is a
    test of bench.py's bookkeeping, not of performance: the numbers it prints here mean nothing."""
    import json as _json
    import time
    import types
    import numpy as np
    sys.path.insert(0, ROOT)
    monkeypatch.setenv("B200RANK_BENCH_MODEL", "flan-t5-small")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    import importlib
    import bench
    bench = importlib.reload(bench)
    import b200rank as br
    from b200rank.synthetic import model_cfg
    from oracle import hf_cpu

    class StandIn:
        def __init__(self, cfg, device):
            self.model = None
            self.staged = None
            self.n_launch = 0
            self.t = [0.0, 0.0]
            self.prof = False
            self.tick = 0
            self.pending = {}
            self.cache = {}

        def load_state_dict(self, items):
            self.model = hf_cpu.build_model(model_cfg(bench.MODEL), dict(items), threads=4)

        def _score(self, ids, lengths, yes_id, no_id):
            # per-document cache: like the engine, a document's result does not depend on what shares its pass (the bench scores the same
            # queries over and over, and checks a merged pass against the headline query scored alone)
            self.n_launch += 438
            ids, lengths = np.asarray(ids), np.asarray(lengths)
            keys = [(ids[i, :lengths[i]].tobytes(), yes_id, no_id) for i in range(len(lengths))]
            todo = [i for i, k in enumerate(keys) if k not in self.cache]
            if todo:
                sub, ln = ids[todo], lengths[todo]
                mask = (np.arange(sub.shape[1])[None] < ln[:, None]).astype(np.int64)
                lg, _ = hf_cpu.score_yes_no(self.model, np.asarray(sub, np.int64) * mask, mask, yes_id, no_id, 32)
                lg = (lg + np.random.default_rng(0).normal(0, 0.01, lg.shape)).astype(np.float32)
                for i, row in zip(todo, lg):
                    self.cache[keys[i]] = row
            lg = np.stack([self.cache[k] for k in keys])
            return lg, (np.exp(lg[:, 0]) / np.exp(lg).sum(1)).astype(np.float32)

        def score_yes_no(self, ids, lengths, yes_id, no_id):
            return self._score(ids, lengths, yes_id, no_id)

        def stage(self, ids, lengths):
            self.staged = (np.asarray(ids), np.asarray(lengths))

        def submit_yes_no_staged(self, yes_id, no_id):
            return self.submit_yes_no(self.staged[0], self.staged[1], yes_id, no_id)

        def submit_yes_no(self, ids, lengths, yes_id, no_id):
            self.tick += 1
            self.pending[self.tick] = self._score(ids, lengths, yes_id, no_id)
            return self.tick

        def wait_yes_no(self, ticket):
            return self.pending.pop(ticket)

        def run_yes_no_staged(self, yes_id, no_id):
            self.last = self._score(self.staged[0], self.staged[1], yes_id, no_id)

        def fetch_yes_no(self):
            return self.last

        def event_record(self, which):
            self.t[which] = time.perf_counter()

        def event_elapsed_ms(self):
            return (self.t[1] - self.t[0]) * 1e3

        def launch_count(self):
            return self.n_launch

        def profile(self, on):
            self.prof = on

        def profile_report(self):
            n_tok = int(self.staged[1].sum())
            return {f"gemm_tcgen05<bn256,epi2> M{n_tok} N2048 K512": {"ms": 2.0, "n": 16}, f"gemm_tcgen05<bn256,epi0> M{n_tok} N1152 K512": {"ms": 1.0, "n": 16},
                    "gemm_tcgen05<bn32,epi1> M100 N512 K384": {"ms": 0.2, "n": 16}, "rmsnorm": {"ms": 0.5, "n": 34}, "rmsnorm_small": {"ms": 0.1, "n": 50},
                    "enc_attention_tc2": {"ms": 0.8, "n": 16}}

        def sync(self):
            pass

        def close(self):
            pass

    monkeypatch.setattr(br, "Engine", StandIn)
    monkeypatch.setattr(bench.ClockSampler, "run", lambda self: None)       # no nvidia-smi here
    # the informational text-API leg fails here (no tokenizer pipeline behind the stand-in): the headline line must survive it

    def broken_text_api(*a, **k):
        raise RuntimeError("text leg down")
    monkeypatch.setattr(bench, "text_api_docs_per_s", broken_text_api)
    args = types.SimpleNamespace(gpus=1, steps=3, warmup=3, no_cpu_baseline=False, no_text_api=False, no_pipeline=False, hf_cuda=False, no_hf_cuda=True, no_sustained=False, queries_per_step=qps)
    assert bench.run_engine(args) == 0
    line = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(line) == 1
    d = _json.loads(line[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "parity"):
        assert key in d, key
    assert d["api_text"] == {"unavailable": "RuntimeError: text leg down"}
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["scaling"] == "weak" and d["dtype"] == "bf16" and d["vs_baseline"] is None
    assert d["gpu_launches"] == 3 * 438 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] == 1200 * qps
    assert d["config"]["queries_per_step"] == qps and d["config"]["docs_per_step_per_gpu"] == 100 * qps
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["kernel"].startswith(f"gemm_tcgen05<bn256,epi2> M{18400 * qps}") and r["unit"] == "TFLOP/s" and r["frac"] > 0
    assert abs(r["flop_per_launch"] - 2.0 * 18400 * qps * 2048 * 512) < 1 and r["launches_per_step"] == 16 / 3
    assert r["hbm_kernels"][0]["kernel"] == "rmsnorm_kernel" and r["hbm_kernels"][0]["launches_per_step"] == 34 / 3
    assert "rmsnorm_small" in r["by_kernel_ms_per_step"]
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["cores"] >= 1
    p = d["parity"]
    assert p["docs"] >= 2 and p["within_logit_tolerance"] is True and p["max_abs_logit_diff"] < 0.06 and 0.9 < p["kendall_tau"] <= 1.0
    assert p["inversions_beyond_tolerance"] == 0 and "max_abs_logit_diff" in p["reference_bf16_yardstick"]
    sus = d["sustained"]
    assert sus["steps"] >= 3 and sus["value"] > 0 and 0 < sus["step_frac"] and "peak_source" in sus and "clocks" in sus
    assert r["peak_source"].startswith(("MEASURED_PEAKS.json bf16_tflops (", "fallback burst")) and "frac_of_sustained" in r and "step_peak_source" in r
