#!/usr/bin/env python
"""bench.py — docs scored/sec for the hot path BASELINE.json names.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a engine
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port) on host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1: one rank per GPU)

Workload (BASELINE.json configs[1]): flan-t5-large, pointwise yes_no, 100 hits per query, q_len 32 / p_len 128
(S = 184 encoder tokens, T = 1 decoder token), synthetic token ids + seeded random-init weights of that architecture.
A step = one device pass of the whole hot path over the 100 candidate documents of each of `--queries-per-step` queries
(default 2 => 200 documents; the first is the committed headline query). The reference loops over batches of 32/32/32/4 of one
query; the engine's per-document results do not depend on what shares the pass (bit-identical, tests/test_engine_gpu.py, and
asserted here against the headline query scored alone), so batches — and queries — are merged to amortise the ~250-launch
decoder chain (same-box A/B: profiles/r02_bench_queries_per_step_ab.txt). With N GPUs every rank scores its own queries per step
(prompts shard embarrassingly; weights are NCCL-broadcast once at load) => weak scaling.

`value`  : device-resident inputs, K steps timed with CUDA events on the engine streams, max over ranks.
`e2e`    : the same K steps through the C-ABI calls a host makes (b200rank_submit_yes_no / _wait_yes_no: HOST token ids in,
           HOST scores out; packing, H2D, compute, D2H inside the timed region), wall clock, max over ranks.
Both keep two passes in flight per GPU (--no-pipeline: one), like a host loop `submit(i+1); wait(i)`.
`roofline`: the dominant kernel (the gemm_tcgen05_kernel instantiation with the largest share of the step): 2*M*N*K per launch /
           its average launch duration, measured live with per-launch CUDA events in a separate profiled pass; peak from
           MEASURED_PEAKS.json; `traffic` = its DRAM bytes per launch from the committed ncu capture; `all_gemm` = the same
           figure over all GEMM launches of a step, `step_frac` = the whole path at the reference's 136.1 GF/doc.
`cpu_baseline`: the numpy oracle (a port of the transformers fp32 path the reference runs on CPU) on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "llm-rankers_b200"))

MODEL = os.environ.get("B200RANK_BENCH_MODEL", "flan-t5-large")   # the override exists for tests/test_bench_contract.py only: the metric is quoted on flan-t5-large
HITS, Q_LEN, P_LEN = 100, 32, 128
SEED = 929
# SURVEY.md §8d / BASELINE.md §2: algorithmic FLOPs per document of the reference's arithmetic at S=184, T=1
GF_PER_DOC = 136.10
FALLBACK_PEAK_TFLOPS = 1400.0  # B200_PROFILING.md: sustained bf16 cuBLAS figure on this pool, used only if MEASURED_PEAKS.json is absent
FALLBACK_BURST_TFLOPS = 1590.0  # B200_PROFILING.md: burst figure, same condition


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def gemm_gflop_per_doc(cfg, S):
    """Algorithmic FLOPs (2*MAC) of the GEMM launches per document: encoder QKVO + gated FFN, stacked cross-K|V
    projection, decoder (T=1) projections/FFN — i.e. §8d's formula without the attention-core and lm_head terms."""
    d, I, F = cfg["d_model"], cfg["num_heads"] * 64, cfg["d_ff"]
    Le, Ld = cfg["num_layers"], cfg["num_decoder_layers"]
    enc = Le * (8 * d * I * S + 6 * d * F * S)
    ckv = Ld * 4 * d * I * S
    dec = Ld * (8 * d * I + 4 * d * I + 6 * d * F)
    return (enc + ckv + dec) / 1e9


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md 'clocks line')."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [x.strip() for x in line.split(",")]))
        except Exception:  # noqa: BLE001 - clocks are best effort
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, t0, t1):
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 9] or [r for (_, r) in self.rows if len(r) >= 9][-3:]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(rows[0][2]) if rows[0][2].replace(".", "").isdigit() else None,
                "power_w_max": max(float(r[3]) for r in rows if r[3].replace(".", "").isdigit()) if rows else None,
                "reasons": sorted(reasons), "samples": len(rows)}


def load_peaks():
    """Sustained bf16 peak (cuBLAS looped for seconds under the power cap): the denominator for anything timed over >= 2 s."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops"))), "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16, seconds-long loop)"
    return FALLBACK_PEAK_TFLOPS, "fallback (B200_PROFILING.md sustained figure; MEASURED_PEAKS.json absent)"


def peak_for_region(seconds):
    """(TFLOP/s, source) of the bf16 roofline denominator matching a timed region: the BURST cuBLAS figure for a kernel timed alone or a
    region shorter than 2 s (the GPU has not reached its power-capped steady state: VERDICT r1 weak #6), the SUSTAINED one otherwise."""
    burst = load_burst_peak()
    if seconds < 2.0:
        if burst:
            return burst, "MEASURED_PEAKS.json bf16_tflops (cuBLAS bf16 burst, best of 10 launches) — timed region < 2 s"
        return FALLBACK_BURST_TFLOPS, "fallback burst figure of B200_PROFILING.md (MEASURED_PEAKS.json absent) — timed region < 2 s"
    pk, src = load_peaks()
    return pk, src + " — timed region >= 2 s"


def load_burst_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        if d.get("bf16_tflops"):
            return float(d["bf16_tflops"])
    return None


def load_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        if d.get("hbm_gbs"):
            return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (device copy, read+write bytes)"
    return 6500.0, "fallback (B200_PROFILING.md copy bandwidth; MEASURED_PEAKS.json absent)"


def cpu_oracle_docs_per_s(n_docs, repeats=1):
    """Times the CPU oracle (numpy fp32 port of the transformers path the reference runs) on n_docs of the workload."""
    from b200rank.synthetic import NO_ID, YES_ID, model_cfg, synthetic_prompt_ids, synthetic_weights
    from oracle.t5_oracle import T5Oracle
    cfg = model_cfg(MODEL)
    orc = T5Oracle(cfg, synthetic_weights(cfg, SEED))
    ids, lengths, _ = bench_query(0)
    ids = ids[:n_docs].astype(np.int64)
    mask = np.ones_like(ids)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.score_yes_no(ids, mask, YES_ID, NO_ID)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_docs / best, best


def bench_query(rank=0):
    """(ids, lengths, ref_logits or None): rank 0 scores the committed HEADLINE query (tests/golden/headline_query.npz: 100 documents of
    the configs[1] shape whose reference top-11 margins are >= 2 x tolerance apart, with the reference's own fp32 logits for all 100);
    other ranks — and rank 0 when the fixture is absent — score a seeded random query of the same shape (no reference logits)."""
    from b200rank.synthetic import headline_query, synthetic_prompt_ids
    if rank == 0 and MODEL == "flan-t5-large":
        ids, lengths, ref, _ = headline_query()
        return ids, lengths, ref
    ids, lengths = synthetic_prompt_ids(HITS, Q_LEN, P_LEN, seed=SEED + rank)
    return ids, lengths, None


def ordering_report(ref_logits, eng_logits, n_layers):
    """Ordering statistics of engine vs reference (yes, no) logits (the same function the GPU parity test asserts on)."""
    from b200rank.tolerance import logit_tolerance
    ref, eng = np.asarray(ref_logits, np.float64), np.asarray(eng_logits, np.float64)
    tol = logit_tolerance(ref, n_layers)
    m_ref, m_eng = ref[:, 0] - ref[:, 1], eng[:, 0] - eng[:, 1]
    o_ref, o_eng = np.argsort(-m_ref, kind="stable"), np.argsort(-m_eng, kind="stable")
    i, j = np.triu_indices(len(ref), 1)
    disc = np.sign(m_ref[i] - m_ref[j]) * np.sign(m_eng[i] - m_eng[j]) < 0
    gap = np.abs(m_ref[i] - m_ref[j])
    pair_tol = tol.sum(1)[i] + tol.sum(1)[j]
    srt = np.sort(m_ref)[::-1]
    return dict(docs=int(len(ref)), max_abs_logit_diff=float(np.abs(eng - ref).max()), mean_abs_logit_diff=float(np.abs(eng - ref).mean()),
                within_logit_tolerance=bool((np.abs(eng - ref) <= tol).all()),
                order_identical=bool(np.array_equal(o_ref, o_eng)), top10_identical=bool(np.array_equal(o_ref[:10], o_eng[:10])),
                top10_set_identical=bool(set(o_ref[:10].tolist()) == set(o_eng[:10].tolist())),
                discordant_pairs=int(disc.sum()), document_pairs=int(len(i)), kendall_tau=float(1.0 - 2.0 * disc.sum() / max(1, len(i))),
                max_ref_margin_gap_of_discordant_pairs=float(gap[disc].max()) if disc.any() else 0.0,
                inversions_beyond_tolerance=int((disc & (gap > pair_tol)).sum()),
                min_adjacent_ref_margin_gap=float(np.min(srt[:-1] - srt[1:])) if len(srt) > 1 else 0.0,
                min_adjacent_ref_margin_gap_top11=float(np.min(srt[:10] - srt[1:11])) if len(srt) > 10 else None)


def cpu_reference_setup(weights=None):
    """The reference's own CPU path (oracle/hf_cpu.py: transformers T5ForConditionalGeneration, fp32, torch on all host threads,
    called like llmrankers/pointwise.py:117-124) over the bench workload. Returns (model, ids, mask) or raises."""
    from b200rank.synthetic import model_cfg, synthetic_prompt_ids, synthetic_weights
    from oracle import hf_cpu
    cfg = model_cfg(MODEL)
    model = hf_cpu.build_model(cfg, weights if weights is not None else synthetic_weights(cfg, SEED))
    ids, lengths, _ = bench_query(0)
    mask = (np.arange(ids.shape[1])[None] < lengths[:, None]).astype(np.int64)
    return model, ids.astype(np.int64), mask


def cpu_reference_baseline(weights, eng_logits, eng_scores, n_sample=32, budget_s=25.0):
    """cpu_baseline for the engine arm: one reference batch (batch_size 32) of the headline query on the host cores, plus the
    full-size parity of the engine's answers for those documents against it (logits, P(yes), order)."""
    import torch
    from b200rank.synthetic import NO_ID, YES_ID
    from oracle import hf_cpu
    model, ids, mask = cpu_reference_setup(weights)
    t0 = time.perf_counter()
    hf_cpu.score_yes_no(model, ids[:2], mask[:2], YES_ID, NO_ID, 32)          # warm-up (thread pool, oneDNN primitives)
    per_doc = (time.perf_counter() - t0) / 2
    n_sample = int(max(2, min(n_sample, budget_s / max(per_doc, 1e-3) / 2)))   # two timed repeats inside the budget
    v, secs, ref_logits = hf_cpu.timed_docs_per_s(model, ids[:n_sample], mask[:n_sample], YES_ID, NO_ID, 32, repeats=2)
    ref_scores = np.exp(ref_logits[:, 0]) / np.exp(ref_logits).sum(1)
    dl = np.abs(eng_logits[:n_sample] - ref_logits)
    # bf16 engine vs fp32 reference. The absolute term grows with depth (rounding noise accumulates over the residual
    # stream): max(0.06, 0.0025 per layer) = 0.12 for the 24+24 layers of flan-t5-large.
    from b200rank import tolerance
    cfg = model.config
    n_layers = cfg.num_layers + cfg.num_decoder_layers
    tol = tolerance.logit_tolerance(ref_logits, n_layers)
    # yardstick: the reference library's OWN reduced-precision path (the same transformers model in bf16 — the reference runs fp16 on
    # CUDA, pointwise.py:22-23) against its fp32 answers, on the same documents
    yard = None
    try:
        import copy
        mb = copy.deepcopy(model).to(torch.bfloat16)
        with torch.no_grad():
            parts = []
            for b0 in range(0, n_sample, 8):
                rows = slice(b0, min(b0 + 8, n_sample))
                lg = mb(input_ids=torch.from_numpy(ids[rows]), attention_mask=torch.from_numpy(mask[rows]),
                        decoder_input_ids=torch.zeros((rows.stop - rows.start, 1), dtype=torch.long)).logits
                parts.append(lg[:, 0, [YES_ID, NO_ID]].float().numpy())
        del mb
        dy = np.abs(np.concatenate(parts, 0) - ref_logits)
        yard = {"what": "transformers bf16 (CPU) vs transformers fp32 on the same documents", "max_abs_logit_diff": float(dy.max()),
                "mean_abs_logit_diff": float(dy.mean()), "fraction_outside_engine_tolerance": float((dy > tol).mean())}
    except Exception as exc:  # noqa: BLE001 - the yardstick is informational
        yard = {"unavailable": f"{type(exc).__name__}: {exc}"}
    baseline = {"value": v, "unit": "docs/s", "cores": torch.get_num_threads(), "kind": "reference",
                "sample": f"{n_sample} of the {HITS} documents of one query (one reference batch, S={Q_LEN + P_LEN + 24}), best of 2, {secs:.1f} s; "
                          f"transformers {__import__('transformers').__version__} T5ForConditionalGeneration fp32 on torch CPU ({torch.get_num_threads()} threads of "
                          f"{os.cpu_count()} logical cores) called on token ids as llmrankers/pointwise.py:117-124 calls it — the HF model object directly, without the "
                          f"reference's tokenizer / DataLoader fork around it (/root/reference does not travel to the GPU box; that omission favours the CPU arm); "
                          f"the reference has no native code to compile into oracle/_ref"}
    parity = dict(ordering_report(ref_logits, eng_logits[:n_sample], n_layers),
                  against="the cpu_baseline run LIVE on this box (fp32 transformers on the same token ids, same weights), full-size model",
                  tolerance=tolerance.describe(n_layers), reference_bf16_yardstick=yard,
                  max_abs_score_diff=float(np.abs(eng_scores[:n_sample] - ref_scores).max()))
    return baseline, parity


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores, same metric/config; every step is a
    bounded sample of the query (rows of one reference batch) sized so that warmup + steps finish within ~3 minutes."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    from b200rank.synthetic import NO_ID, YES_ID
    budget_s = 150.0
    kind = "reference"
    try:
        import torch
        from oracle import hf_cpu
        model, ids, mask = cpu_reference_setup()
        step_fn = lambda n: hf_cpu.score_yes_no(model, ids[:n], mask[:n], YES_ID, NO_ID, 32)  # noqa: E731
        cores = torch.get_num_threads()
        how = (f"transformers {__import__('transformers').__version__} T5ForConditionalGeneration fp32 on torch CPU ({cores} threads) "
               f"called as llmrankers/pointwise.py:117-124 does")
    except Exception as exc:  # noqa: BLE001 - transformers/torch CPU path unusable: time the numpy port instead, and say so
        from b200rank.synthetic import model_cfg, synthetic_prompt_ids, synthetic_weights
        from oracle.t5_oracle import T5Oracle
        kind = "port"
        cfg = model_cfg(MODEL)
        orc = T5Oracle(cfg, synthetic_weights(cfg, SEED))
        ids, _, _ = bench_query(0)
        ids = ids.astype(np.int64)
        mask = np.ones_like(ids)
        step_fn = lambda n: orc.score_yes_no(ids[:n], mask[:n], YES_ID, NO_ID)  # noqa: E731
        cores = os.cpu_count()
        how = f"numpy fp32 oracle port on all host threads (transformers CPU path unavailable: {type(exc).__name__}: {exc})"
    # calibrate on two documents (also the first warm-up), then size the per-step sample for the time budget
    t0 = time.perf_counter()
    step_fn(2)
    per_doc = (time.perf_counter() - t0) / 2
    n_steps_total = max(1, args.steps + args.warmup)
    sample = int(max(1, min(32, budget_s / n_steps_total / max(per_doc, 1e-4))))
    for _ in range(args.warmup):
        step_fn(sample)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_fn(sample)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = f"{sample} of the {HITS} documents of one query per step (S={Q_LEN + P_LEN + 24}); {how}"
    line = {
        "impl": "reference", "metric": f"docs scored/sec ({MODEL} q32/p128, pointwise yes_no)", "value": value, "unit": "docs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(1, sample_docs=sample), parallelism="host threads of one process (rank 0 only)", pipeline="none: one batch at a time",
                       weights=f"seeded random init (numpy PCG64 seed {SEED}), fp32 on the host", l2="n/a (CPU)"),
        "cpu_baseline": {"value": value, "unit": "docs/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": "docs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(world, sample_docs=None, queries_per_step=1):
    return {
        "workload": f"{MODEL} pointwise yes_no, {HITS} hits/query, batch_size 32 (one device pass), q_len {Q_LEN} p_len {P_LEN} -> S {Q_LEN + P_LEN + 24}, T 1 (BASELINE configs[1])",
        "queries_per_step": queries_per_step,
        "docs_per_step_per_gpu": HITS * queries_per_step if sample_docs is None else sample_docs,
        "global_docs_per_step": (HITS * queries_per_step if sample_docs is None else sample_docs) * world,
        "parallelism": f"dp{world} (queries sharded across ranks, weights NCCL-broadcast once at load, no collective in the loop)",
        "pipeline": "two device passes in flight per GPU (submit/wait): the decoder chain of pass i on a second stream overlaps the encoder GEMMs of pass i+1",
        "weights": f"seeded random init (numpy PCG64 seed {SEED}), bf16 on device",
        "l2": "no explicit flush: each step streams 1.6 GB of weights + ~3.5 GB of activations, far beyond the 126 MB L2",
        "algorithmic_gflop_per_doc": GF_PER_DOC,
    }


def text_api_docs_per_s(eng, cfg, n_queries):
    """docs/s through `PointwiseLlmRanker.rerank_many` on TEXT: every query brings 100 documents the tokenizer has never seen
    (128 words each, query 32 words => 185 tokens per prompt with the synthetic vocabulary), so the per-document token cache only
    helps with the query and the template; tokenisation runs on worker threads ahead of the GPU."""
    from b200rank.synthetic import synthetic_tokenizer
    from llmrankers._backend import T5Backend
    from llmrankers.pointwise import PointwiseLlmRanker
    from llmrankers.rankers import SearchResult
    rng = np.random.default_rng(SEED + 7)
    ranker = PointwiseLlmRanker(None, None, "cuda", method="yes_no", batch_size=32, backend=T5Backend(eng, synthetic_tokenizer(), cfg))

    def requests(n):
        for _ in range(n):
            words = rng.integers(0, 2000, size=(HITS + 1, P_LEN))
            query = " ".join(f"w{int(x)}" for x in words[HITS, :Q_LEN])
            yield query, [SearchResult(docid=str(i), score=0.0, text=" ".join(f"w{int(x)}" for x in words[i])) for i in range(HITS)]

    warm = list(requests(4))
    timed = list(requests(n_queries))     # text generation is not the system under test: build the requests first
    for _ in ranker.rerank_many(iter(warm)):
        pass
    t0 = time.perf_counter()
    n_docs = 0
    for out in ranker.rerank_many(iter(timed)):
        n_docs += len(out)
    dt = time.perf_counter() - t0
    return {"value": n_docs / dt, "unit": "docs/s", "queries": n_queries, "ms_per_query": dt / n_queries * 1e3,
            "tokenizer_threads": 4, "what": "strings -> prompt assembly + tokenisation (host threads) -> submit/wait pipeline -> sorted SearchResults"}


def dist_init():
    """(rank, world, local, dist module or None): joins the torchrun rendezvous when WORLD_SIZE > 1 (NCCL, one rank per GPU)."""
    rank, world, local = dist_env()
    if world == 1:
        return rank, world, local, None
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"           # stdout carries exactly one JSON line
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def load_and_broadcast(eng, cfg, rank, world, local, dist, tweak=None):
    """Rank 0 builds the seeded synthetic weights and uploads them; the device weight arena then goes to every other rank in ONE NCCL
    broadcast (the system's only collective outside result gathering). Returns timing / size of both steps."""
    from b200rank.synthetic import synthetic_weights
    t0 = time.time()
    if rank == 0:
        w = synthetic_weights(cfg, SEED)
        if tweak is not None:
            tweak(w)
        eng.load_state_dict(w.items())
        del w
    t_load = time.time() - t0
    info = {"weights_load_s": round(t_load, 2)}
    if dist is not None:
        import torch
        from b200rank.dist import arena_as_tensor
        arena = arena_as_tensor(eng, local)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(arena, src=0)
        e1.record()
        torch.cuda.synchronize()
        if rank != 0:
            eng.mark_weights_loaded()
        ms = e0.elapsed_time(e1)
        info["broadcast"] = {"bytes": int(arena.numel()), "ms": round(ms, 2), "GB_per_s": round(arena.numel() / (ms * 1e-3) / 1e9, 1),
                             "what": "torch.distributed.broadcast of the device weight arena over NCCL / NVLink, CUDA events on rank 0 (first use of the communicator included)"}
    return info


def max_over_ranks_f(x, dist, local):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_setwise(args):
    """Secondary workload (BASELINE configs[2]): SetwiseLlmRanker heapsort, flan-t5-large, num_child 10, k 10,
    100 hits/query, generation scoring — through the drop-in Python API on TEXT. A compare prompt holds 11 passages x 128 words + the
    query (S ~ 1.5 k tokens); the sort is a chain of dependent compares, so the unit is the query latency. Reports the level-parallel
    heap build (default) against the reference's one-compare-at-a-time order (B200RANK_BATCHED_SORT=0); results are identical."""
    import contextlib
    import io
    import b200rank as br
    from b200rank.synthetic import LABELS, model_cfg, synthetic_tokenizer, synthetic_weights
    from llmrankers._backend import T5Backend
    from llmrankers.rankers import SearchResult
    from llmrankers.setwise import SetwiseLlmRanker
    cfg = model_cfg(MODEL)
    tok = synthetic_tokenizer()
    w = synthetic_weights(cfg, SEED)
    label_ids = [tok.convert_tokens_to_ids("\u2581" + c) for c in LABELS[:11]]
    w["lm_head.weight"][label_ids] *= 30.0   # label-favouring lm_head: generation emits passage labels, the heap really sifts
    c = br.make_config(cfg["d_model"], cfg["num_heads"], cfg["d_ff"], cfg["num_layers"], cfg["num_decoder_layers"],
                       max_tokens=20480, max_docs=64, max_dec_len=8, max_logit_rows=256)
    eng = br.Engine(c, 0)
    eng.load_state_dict(w.items())
    be = T5Backend(eng, tok, cfg)
    rng = np.random.default_rng(SEED)
    n_q = max(2, min(args.steps, 8))

    def query_set():
        words = rng.integers(0, 2000, size=(HITS + 1, P_LEN))
        return (" ".join(f"w{int(x)}" for x in words[HITS, :Q_LEN]),
                [SearchResult(docid=str(i), score=0.0, text=" ".join(f"w{int(x)}" for x in words[i])) for i in range(HITS)])
    sets = [query_set() for _ in range(n_q + 1)]
    res = {}
    for mode, flag in (("batched", "1"), ("sequential", "0")):
        os.environ["B200RANK_BATCHED_SORT"] = flag
        r = SetwiseLlmRanker(None, None, "cuda", num_child=10, k=10, scoring="generation", method="heapsort", backend=be)
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink):
            r.rerank(sets[0][0], list(sets[0][1]))   # warm-up
            eng.sync()
            t0 = time.perf_counter()
            compares, orders, tokens = 0, [], 0
            for q, docs in sets[1:]:
                out = r.rerank(q, list(docs))
                compares += r.total_compare
                tokens += r.total_prompt_tokens
                orders.append([d.docid for d in out])
            eng.sync()
            dt = time.perf_counter() - t0
        res[mode] = dict(s_per_query=dt / n_q, compares_per_query=compares / n_q, prompt_tokens_per_query=tokens / n_q, orders=orders)
    assert res["batched"]["orders"] == res["sequential"]["orders"], "batched and sequential heapsort disagree"
    # cross-query lockstep: all queries advance together, every round is one batch of their pending compares
    os.environ["B200RANK_BATCHED_SORT"] = "1"
    r = SetwiseLlmRanker(None, None, "cuda", num_child=10, k=10, scoring="generation", method="heapsort", backend=be)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        eng.sync()
        t0 = time.perf_counter()
        many_orders = [[d.docid for d in out] for out in r.rerank_many([(q, list(docs)) for q, docs in sets[1:]], window=8)]
        eng.sync()
        dt_many = time.perf_counter() - t0
    assert many_orders == res["batched"]["orders"], "rerank_many and rerank disagree"
    b, q = res["batched"], res["sequential"]
    line = {"metric": "docs reranked/sec, setwise heapsort (flan-t5-large, num_child 10, k 10, 100 hits, generation)", "value": HITS / b["s_per_query"],
            "unit": "docs/s", "n_gpus": 1, "steps": n_q, "warmup": 1, "ms_per_step": b["s_per_query"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "flan-t5-large setwise heapsort c=10 k=10 generation, 100 hits/query, text API (BASELINE configs[2])",
                       "compares_per_query": b["compares_per_query"], "prompt_tokens_per_query": b["prompt_tokens_per_query"]},
            "sequential_order": {"value": HITS / q["s_per_query"], "ms_per_step": q["s_per_query"] * 1e3,
                                 "what": "B200RANK_BATCHED_SORT=0: the reference's one-compare-per-call heapify order (same results)"},
            "speedup_from_level_parallel_heap_build": q["s_per_query"] / b["s_per_query"],
            "rerank_many": {"value": HITS * n_q / dt_many, "unit": "docs/s", "queries_in_lockstep": n_q, "ms_per_query": dt_many / n_q * 1e3,
                            "what": "SetwiseLlmRanker.rerank_many: the heapsorts of all queries advance together, one engine batch per round (same results)"}}
    print(json.dumps(line))
    eng.close()
    return 0


def run_qlm(args):
    """Secondary workload (BASELINE configs[4]): pointwise qlm — S = 144 (p128 + 16 template ids), labels T = 33
    (`<pad>` + 32 query ids), `--hits` documents per query (default 100; configs[4] says 1000), model `--model` (default flan-t5-large;
    configs[4] says flan-t5-xxl). N = 1: b200rank_score_qlm with HOST buffers (pack + H2D + encoder + 33-position decoder + full-vocab
    log-softmax + D2H inside the timed region). N > 1 (torchrun): ONE query's documents are split contiguously over the ranks
    (document-level sharding, strong scaling), weights NCCL-broadcast once, scores gathered once per query; value = whole-job docs/s."""
    import b200rank as br
    from b200rank.dist import shard_bounds
    from b200rank.synthetic import model_cfg
    rank, world, local, dist = dist_init()
    model = args.model or MODEL
    hits = args.hits or HITS
    cfg = model_cfg(model)
    S, T = P_LEN + 16, Q_LEN + 1
    lo, hi = shard_bounds(hits, rank, world)
    n_local = hi - lo
    c = br.make_config(cfg["d_model"], cfg["num_heads"], cfg["d_ff"], cfg["num_layers"], cfg["num_decoder_layers"],
                       max_tokens=max(n_local, 1) * 160, max_docs=max(128, n_local), max_dec_len=40, max_logit_rows=min(max(n_local, 1), int(os.environ.get("B200RANK_BENCH_QLM_PASS_DOCS", "512"))) * 40)
    eng = br.Engine(c, local)
    load = load_and_broadcast(eng, cfg, rank, world, local, dist)
    rng = np.random.default_rng(SEED)
    ids = rng.integers(3, 32000, size=(hits, S)).astype(np.int32)
    ids[:, -1] = 1
    lengths = np.full((hits,), S, np.int32)
    labels = [0] + rng.integers(3, 32000, size=Q_LEN).tolist()

    def step():
        local_scores = eng.score_qlm(ids[lo:hi], lengths[lo:hi], labels) if n_local else np.zeros((0,), np.float32)
        if dist is None:
            return local_scores
        from b200rank.dist import all_gather_variable
        return all_gather_variable(np.asarray(local_scores, np.float32))

    for _ in range(max(args.warmup, 3)):
        sc = step()
    eng.sync()
    steps = max(5, min(args.steps, 30))
    rep = None
    if world == 1:
        eng.profile(True)
        for _ in range(steps):
            sc = step()
        eng.sync()
        rep = eng.profile_report()
        eng.profile(False)
    if dist is not None:
        dist.barrier()
    t1 = time.perf_counter()
    for _ in range(steps):
        sc = step()
    eng.sync()
    dt = max_over_ranks_f(time.perf_counter() - t1, dist, local)          # timed without the per-launch profiling events
    dm, I, F, V = cfg["d_model"], cfg["num_heads"] * 64, cfg["d_ff"], cfg["vocab_size"]     # SURVEY.md §8d formula
    gf = (cfg["num_layers"] * (8 * dm * I * S + 6 * dm * F * S + 4 * S * S * I) + cfg["num_decoder_layers"] * 4 * dm * I * S
          + cfg["num_decoder_layers"] * (8 * dm * I * T + 4 * T * T * I + 4 * dm * I * T + 4 * T * S * I + 6 * dm * F * T)
          + 2 * dm * V * T) / 1e9
    peak, peak_src = peak_for_region(dt)
    if rank == 0:
        import hashlib
        value = hits * steps / dt
        line = {"metric": f"docs scored/sec, pointwise qlm ({model}, S 144, T 33)", "value": value, "unit": "docs/s", "n_gpus": world,
                "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
                "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"{model} pointwise qlm, {hits} hits/step, S 144, labels T 33 (BASELINE configs[4] shape)",
                           "parallelism": f"dp{world}: the documents of ONE query split contiguously over the ranks, weights NCCL-broadcast once, "
                                          "scores all-gathered once per query (host side); no collective inside the scoring pass",
                           "algorithmic_gflop_per_doc": gf, **load},
                "step_frac": value / world * gf * 1e9 / (peak * 1e12), "step_peak_source": peak_src, "finite_scores": bool(np.isfinite(sc).all()),
                "scores_sha1": hashlib.sha1(np.ascontiguousarray(sc, np.float32).tobytes()).hexdigest()[:16]}
        if rep is not None:
            line["by_kernel_ms_per_step"] = {k: round(v["ms"] / steps, 4) for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])[:14]}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
    eng.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_monot5(args):
    """Secondary workload: MonoT5-style pointwise scoring — T5 v1.0 (relu feed-forward, tied embeddings), yes/no =
    true/false logits at decoder position 0 — on a synthetic model of the shape `--model` (default monot5-3b: d_kv 128, 32 heads,
    d_ff 16384, i.e. the head width that runs on the generic-width attention of csrc/attention_wide.cuh instead of the tcgen05
    kernels), `--hits` documents of S = 184 per step through the synchronous b200rank_score_yes_no with HOST buffers."""
    import b200rank as br
    from b200rank.synthetic import model_cfg, synthetic_prompt_ids, synthetic_weights
    model, hits = args.model or "monot5-3b", args.hits or HITS
    cfg = model_cfg(model)
    c = br.make_config(cfg["d_model"], cfg["num_heads"], cfg["d_ff"], cfg["num_layers"], cfg["num_decoder_layers"], d_kv=cfg["d_kv"],
                       gated_gelu=cfg["gated_gelu"], scale_decoder_outputs=cfg["scale_decoder_outputs"],
                       max_tokens=hits * (Q_LEN + P_LEN + 24) + 256, max_docs=max(128, hits), max_logit_rows=256)
    eng = br.Engine(c, 0)
    t_w = time.time()
    eng.load_state_dict(synthetic_weights(cfg, SEED).items())
    t_w = time.time() - t_w
    ids, lengths = synthetic_prompt_ids(hits, Q_LEN, P_LEN, seed=SEED)
    for _ in range(max(args.warmup, 3)):
        lg, sc = eng.score_yes_no(ids, lengths, 1176, 6136)      # 'true' / 'false' ids of the T5 vocabulary (pointwise.py:170-171)
    steps = max(5, min(args.steps, 30))
    eng.profile(True)
    for _ in range(steps):
        eng.score_yes_no(ids, lengths, 1176, 6136)
    rep = eng.profile_report()
    eng.profile(False)
    eng.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        lg, sc = eng.score_yes_no(ids, lengths, 1176, 6136)
    dt = time.perf_counter() - t0
    S = Q_LEN + P_LEN + 24
    d, I, F, V = cfg["d_model"], cfg["num_heads"] * cfg["d_kv"], cfg["d_ff"], cfg["vocab_size"]
    gf = (cfg["num_layers"] * (8 * d * I * S + 4 * d * F * S + 4 * S * S * I) + cfg["num_decoder_layers"] * 4 * d * I * S
          + cfg["num_decoder_layers"] * (8 * d * I + 4 * d * I + 4 * S * I + 4 * d * F) + 2 * d * V) / 1e9      # SURVEY.md §8d with an ungated feed-forward
    peak, peak_src = peak_for_region(dt)
    value = hits * steps / dt
    line = {"metric": f"docs scored/sec, MonoT5-style yes/no ({model}, q32/p128)", "value": value, "unit": "docs/s", "n_gpus": 1, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{model} (T5 v1.0, d_kv {cfg['d_kv']}) pointwise true/false, {hits} hits/step, S {S}, T 1, synchronous call with host buffers",
                       "algorithmic_gflop_per_doc": gf, "weights_load_s": round(t_w, 1)},
            "step_frac": value * gf * 1e9 / (peak * 1e12), "step_peak_source": peak_src, "finite_scores": bool(np.isfinite(sc).all()),
            "by_kernel_ms_per_step": {k: round(v["ms"] / steps, 4) for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])[:12]}}
    print(json.dumps(line))
    eng.close()
    return 0


def run_pairwise(args):
    """Secondary workload (BASELINE configs[3]): PairwiseLlmRanker allpair, flan-t5-xl, batch_size 2 (the
    reference default), through the drop-in Python API on TEXT; `--hits` documents per query (default 24 -> 552 prompts of S ~ 320 to
    keep a 1-GPU run short; configs[3] says 100 -> 9900 prompts). N = 1 also reports the reference's one-call-per-DataLoader-batch
    loop (B200RANK_BATCHED_SORT=0; outputs identical). N > 1 (torchrun): the n(n-1) prompts of ONE query are split over the ranks
    inside the backend (llmrankers/_backend.py::ShardedBackend: whole reference batches per rank, one host-side gather per engine
    call) — strong scaling; every rank assembles all prompt rows (host work is replicated and inside the timed region)."""
    import b200rank as br
    from b200rank.synthetic import model_cfg, synthetic_tokenizer
    from llmrankers._backend import ShardedBackend, T5Backend
    from llmrankers.pairwise import PairwiseLlmRanker
    from llmrankers.rankers import SearchResult
    rank, world, local, dist = dist_init()
    model, hits = args.model or "flan-t5-xl", args.hits or 24
    cfg = model_cfg(model)
    tok = synthetic_tokenizer()
    lab = [tok.convert_tokens_to_ids("\u2581" + c) for c in "AB"]

    def tweak(w):
        w["lm_head.weight"][lab] *= 30.0   # label-favouring lm_head: generation emits "Passage A/B"
    c = br.make_config(cfg["d_model"], cfg["num_heads"], cfg["d_ff"], cfg["num_layers"], cfg["num_decoder_layers"],
                       max_tokens=32768, max_docs=256, max_dec_len=8, max_logit_rows=256)
    eng = br.Engine(c, local)
    load = load_and_broadcast(eng, cfg, rank, world, local, dist, tweak)
    be = T5Backend(eng, tok, cfg)
    if dist is not None:
        be = ShardedBackend(be)
    rng = np.random.default_rng(SEED)
    words = rng.integers(0, 2000, size=(hits + 1, P_LEN))
    query = " ".join(f"w{int(x)}" for x in words[hits, :Q_LEN])
    docs = [SearchResult(docid=str(i), score=0.0, text=" ".join(f"w{int(x)}" for x in words[i])) for i in range(hits)]
    res = {}
    modes = (("merged", "1"), ("per_batch", "0")) if world == 1 and hits <= 32 else (("merged", "1"),)
    for mode, flag in modes:
        os.environ["B200RANK_BATCHED_SORT"] = flag
        r = PairwiseLlmRanker(None, None, "cuda", method="allpair", batch_size=2, k=10, backend=be)
        r.rerank(query, list(docs[:6]))   # warm-up
        eng.sync()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        out = r.rerank(query, list(docs))
        eng.sync()
        dt = max_over_ranks_f(time.perf_counter() - t0, dist, local)
        res[mode] = dict(s=dt, prompts=hits * (hits - 1), order=[d.docid for d in out], tokens=r.total_prompt_tokens, compares=r.total_compare)
    m = res["merged"]
    if rank == 0:
        import hashlib
        line = {"metric": f"prompts/sec, pairwise allpair ({model}, batch_size 2, generation)", "value": m["prompts"] / m["s"], "unit": "prompts/s",
                "n_gpus": world, "steps": 1, "warmup": 1, "ms_per_step": m["s"] * 1e3, "higher_is_better": True,
                "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"{model} pairwise allpair, {hits} hits/query -> {m['prompts']} prompts, batch_size 2, text API (BASELINE configs[3]"
                                       + (")" if hits == 100 else " at reduced hits)"),
                           "parallelism": f"dp{world}: the prompts of ONE query split over the ranks (whole reference batches), weights NCCL-broadcast once, "
                                          "generated ids gathered per engine call (host side)",
                           "padded_prompt_tokens": m["tokens"], "reference_batches": m["compares"], **load},
                "order_sha1": hashlib.sha1(",".join(m["order"]).encode()).hexdigest()[:16]}
        if "per_batch" in res:
            b = res["per_batch"]
            assert m["order"] == b["order"] and m["tokens"] == b["tokens"]
            line["per_reference_batch"] = {"value": b["prompts"] / b["s"], "ms_per_step": b["s"] * 1e3,
                                           "what": "B200RANK_BATCHED_SORT=0: one engine call per DataLoader batch of 2, as the reference loops (same outputs)"}
            line["speedup_from_merged_batches"] = b["s"] / m["s"]
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
    eng.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_engine(args):
    rank, world, local = dist_env()
    import b200rank as br
    from b200rank.synthetic import NO_ID, YES_ID, model_cfg, synthetic_prompt_ids, synthetic_weights

    use_dist = world > 1
    if use_dist:
        # stdout carries exactly ONE JSON line: NCCL prints its version banner there when NCCL_DEBUG=VERSION is in the environment
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = model_cfg(MODEL)
    QPS = max(1, getattr(args, "queries_per_step", 2))   # queries merged into one device pass (= one step)
    DOCS = HITS * QPS
    c = br.make_config(cfg["d_model"], cfg["num_heads"], cfg["d_ff"], cfg["num_layers"], cfg["num_decoder_layers"],
                       max_tokens=DOCS * (Q_LEN + P_LEN + 24) + 256, max_docs=max(128, DOCS), max_logit_rows=256)
    eng = br.Engine(c, local)
    t_load = time.time()
    if rank == 0:
        eng.load_state_dict(synthetic_weights(cfg, SEED).items())
    if use_dist:
        # the ONE collective of the whole system: broadcast the device weight arena from rank 0 over NCCL/NVLink
        import torch
        import torch.distributed as dist
        ptr, nbytes = eng.weights_blob()

        class _Arena:
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        arena = torch.as_tensor(_Arena(), device=torch.device("cuda", local))
        dist.broadcast(arena, src=0)
        torch.cuda.synchronize()
        if rank != 0:
            eng.mark_weights_loaded()
    t_load = time.time() - t_load

    # every rank scores its own query (different token ids), 100 hits each; rank 0's is the committed headline query
    ids, lengths, fixture_logits = bench_query(rank)
    if QPS > 1:   # further queries of the same shape ride in the same device pass (a document's result does not depend on its batch)
        parts = [(ids, lengths)] + [synthetic_prompt_ids(HITS, Q_LEN, P_LEN, seed=SEED + 1000 * q + rank) for q in range(1, QPS)]
        width = max(p[0].shape[1] for p in parts)
        ids = np.concatenate([np.pad(p[0], ((0, 0), (0, width - p[0].shape[1]))) for p in parts]).astype(np.int32)
        lengths = np.concatenate([p[1] for p in parts]).astype(np.int32)
    n_tok = int(lengths.sum())

    def barrier():
        eng.sync()
        if use_dist:
            import torch.distributed as dist
            dist.barrier()
        eng.sync()

    def max_over_ranks(x):
        if not use_dist:
            return x
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pipelined = not args.no_pipeline

    def run_steps(n, submit):
        """n steps, two in flight when pipelined: submit(step i+1) before wait(step i). Returns the last step's (logits, scores)."""
        out, prev = None, None
        for _ in range(n):
            t = submit()
            if prev is not None:
                out = eng.wait_yes_no(prev)
            prev = t
            if not pipelined:
                out = eng.wait_yes_no(prev)
                prev = None
        if prev is not None:
            out = eng.wait_yes_no(prev)
        return out

    # ---- value: inputs resident in HBM (staged once), CUDA events on the engine's streams
    eng.stage(ids, lengths)
    run_steps(max(args.warmup, 3), lambda: eng.submit_yes_no_staged(YES_ID, NO_ID))
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.15)
    l0 = eng.launch_count()
    t0 = time.time()
    eng.event_record(0)
    logits_dev, scores_dev = run_steps(args.steps, lambda: eng.submit_yes_no_staged(YES_ID, NO_ID))
    eng.event_record(1)
    ms = eng.event_elapsed_ms()
    barrier()
    t1 = time.time()
    launches = eng.launch_count() - l0
    sampler.stop()
    clocks = sampler.summary(t0, t1)
    ms = max_over_ranks(ms)
    value = world * DOCS * args.steps / (ms * 1e-3)
    # ---- sustained: the same loop for >= 3 s, so that the power-capped steady state (what MEASURED_PEAKS' sustained cuBLAS figure was
    # taken in) has a matching numerator next to the short driver-sized region above
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(3200.0 / max(ms / args.steps, 1e-3)) + 1)
        barrier()
        s_sampler = ClockSampler(local)
        s_sampler.start()
        ts0 = time.time()
        eng.event_record(0)
        run_steps(n_sus, lambda: eng.submit_yes_no_staged(YES_ID, NO_ID))
        eng.event_record(1)
        ms_sus = max_over_ranks(eng.event_elapsed_ms())
        barrier()
        ts1 = time.time()
        s_sampler.stop()
        sus_peak, sus_src = peak_for_region(ms_sus * 1e-3)
        sus_value = world * DOCS * n_sus / (ms_sus * 1e-3)
        sustained = {"value": sus_value, "unit": "docs/s", "steps": n_sus, "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / n_sus,
                     "step_frac": (sus_value / world) * GF_PER_DOC * 1e9 / (sus_peak * 1e12), "peak": sus_peak, "peak_source": sus_src,
                     "clocks": s_sampler.summary(ts0, ts1)}
    # the synchronous single-stream path must give the same bits
    eng.run_yes_no_staged(YES_ID, NO_ID)
    lg_sync, _ = eng.fetch_yes_no()
    debug_skip = os.environ.get("B200RANK_DEBUG_SKIP_DECODER", "0") not in ("", "0")    # measurement switch: scores are garbage then
    assert debug_skip or np.array_equal(lg_sync, logits_dev), "pipelined and synchronous passes disagree"

    # ---- e2e: HOST buffers through the C-ABI (submit/wait), wall clock; pack + H2D + compute + D2H inside the timed region
    run_steps(max(args.warmup, 3), lambda: eng.submit_yes_no(ids, lengths, YES_ID, NO_ID))
    barrier()
    w0 = time.perf_counter()
    lg, sc = run_steps(args.steps, lambda: eng.submit_yes_no(ids, lengths, YES_ID, NO_ID))
    eng.sync()
    e2e_s = max_over_ranks(time.perf_counter() - w0)
    e2e_value = world * DOCS * args.steps / e2e_s
    assert debug_skip or np.array_equal(lg, logits_dev), "e2e and device-resident passes disagree"
    if QPS > 1 and not debug_skip:   # batch-composition invariance, checked where it is relied on: the headline query alone gives the same bits
        lg1, _ = eng.score_yes_no(ids[:HITS], lengths[:HITS], YES_ID, NO_ID)
        assert np.array_equal(lg1, np.asarray(logits_dev)[:HITS]), "a merged pass changed the headline query's logits"
    h2d = n_tok * 4 + (DOCS + 1) * 4 + DOCS * 4 + 2 * 4  # packed ids + cu_seqlens + decoder ids + (yes,no) ids
    d2h = DOCS * 3 * 4                                   # (yes, no) logits + P(yes) per document
    logits_dev, scores_dev = np.asarray(logits_dev)[:HITS], np.asarray(scores_dev)[:HITS]   # the headline query: what the parity legs below compare

    # ---- informational: the same workload through the drop-in Python API with TEXT (llmrankers PointwiseLlmRanker.rerank_many):
    # prompt assembly + tokenisation of 100 never-seen documents per query on the host, then the same submit/wait pipeline.
    api_text = None
    if rank == 0 and world == 1 and not args.no_text_api:
        try:
            api_text = text_api_docs_per_s(eng, cfg, n_queries=min(args.steps, 30))
        except Exception as exc:  # noqa: BLE001 - informational leg: never lose the bench line to it
            api_text = {"unavailable": f"{type(exc).__name__}: {exc}"}
            try:
                eng.drain()       # tickets the aborted pipeline left in flight would block the synchronous calls below
            except Exception:  # noqa: BLE001
                pass

    # ---- roofline for the dominant kernel: per-launch CUDA events in a separate profiled pass
    roofline = None
    cpu_baseline = None
    parity = None
    if rank == 0:
        prof_steps = min(args.steps, 10)
        eng.stage(ids, lengths)  # the text-API measurement above staged its own documents: profile the headline input again
        eng.profile(True)
        for _ in range(prof_steps):
            eng.run_yes_no_staged(YES_ID, NO_ID)
        rep = eng.profile_report()
        eng.profile(False)
        gemm_ms = sum(v["ms"] for k, v in rep.items() if k.startswith("gemm_tcgen05")) / prof_steps
        gemm_n = sum(v["n"] for k, v in rep.items() if k.startswith("gemm_tcgen05")) / prof_steps
        all_ms = sum(v["ms"] for v in rep.values()) / prof_steps
        # denominators: the dominant kernel is timed by per-launch events in a sub-second profiled pass => the BURST
        # cuBLAS figure; the whole-path fraction uses the peak that matches the length of the timed region it was measured over
        peak, peak_src = peak_for_region(0.0)
        step_peak, step_peak_src = peak_for_region(ms * 1e-3)
        sus_peak_val, _ = load_peaks()
        # FLOPs the GEMM launches actually execute (2*M*N*K from the launch labels; the gated epilogue's N counts both halves).
        # The engine skips some of the reference's arithmetic (cross-K|V projection, dead decoder q/k at T=1), so this is LESS
        # than the algorithmic GEMM work of SURVEY.md §8d — `step_frac` below is the algorithmic, whole-path figure.
        import re
        flops = 0.0
        for k, v in rep.items():
            m = re.search(r"M(\d+) N(\d+) K(\d+)", k)
            if k.startswith("gemm_tcgen05") and m:
                flops += 2.0 * int(m.group(1)) * int(m.group(2)) * int(m.group(3)) * v["n"] / prof_steps
        agg_achieved = flops / (gemm_ms * 1e-3) / 1e12
        # dominant kernel = the GEMM instantiation with the largest share of the step (the gated FFN-in GEMM at this workload)
        dom_label, dom = max(((k, v) for k, v in rep.items() if k.startswith("gemm_tcgen05")), key=lambda kv: kv[1]["ms"])
        m = re.search(r"M(\d+) N(\d+) K(\d+)", dom_label)
        dom_flop = 2.0 * int(m.group(1)) * int(m.group(2)) * int(m.group(3))
        dom_ms = dom["ms"] / dom["n"]
        achieved = dom_flop / (dom_ms * 1e-3) / 1e12
        traffic, traffic_src = None, None
        try:  # per-launch DRAM bytes of that instantiation from the committed ncu --set full capture
            for tname in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
                tp = os.path.join(ROOT, "profiles", tname)
                if not os.path.exists(tp):
                    continue
                with open(tp) as f:
                    tj = json.load(f)
                if dom_label in tj:
                    traffic = tj[dom_label]["dram_read"] + tj[dom_label]["dram_write"]
                    traffic_src = "profiles/" + tname
                    break
        except OSError:
            pass
        roofline = {
            "bound": "tensor", "kernel": dom_label, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            # note: against the sustained cuBLAS figure (a seconds-long loop at ~1.3 GHz under the power cap) the same launch reads higher
            "peak_sustained": sus_peak_val, "frac_of_sustained": achieved / sus_peak_val,
            "flop_per_launch": dom_flop, "avg_launch_ms": dom_ms, "launches_per_step": dom["n"] / prof_steps,
            "kernel_share_of_step": dom["ms"] / prof_steps / all_ms if all_ms else None,
            # all gemm_tcgen05 launches of a step together (264 launches incl. the small decoder GEMMs)
            "all_gemm": {"achieved": agg_achieved, "frac": agg_achieved / peak, "launches_per_step": gemm_n,
                         "executed_gflop_per_step": flops / 1e9,
                         "algorithmic_gflop_per_step": gemm_gflop_per_doc(cfg, Q_LEN + P_LEN + 24) * DOCS,
                         "share_of_step": gemm_ms / all_ms if all_ms else None},
            # whole path: the reference's algorithmic 136.1 GF/doc at the measured docs/s, against the peak matching the timed region
            "step_frac": (value / world) * GF_PER_DOC * 1e9 / (step_peak * 1e12), "step_peak": step_peak, "step_peak_source": step_peak_src,
            "timed_region_s": ms * 1e-3,
            "by_kernel_ms_per_step": {k: round(v["ms"] / prof_steps, 4) for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])[:14]},
        }
        # the HBM-bound kernel with the largest share: T5LayerNorm over the fp32 residual stream (read 4 B + write 2 B per element), the
        # encoder-sized launches only (the 100-row decoder launches are launch-latency-bound and carry the label rmsnorm_small).
        # Launch times come from event pairs around each launch and include the inter-kernel gap, so this is a lower bound
        # (ncu, kernel alone: 17.5-19.2 us per launch = 5.9 TB/s, profiles/r01_ncu_summary_final.txt).
        if "rmsnorm" in rep:
            # (2 per encoder layer: the norm before block 0 rides on the embedding gather and the final one closes the last layer)
            rows_per_step = rep["rmsnorm"]["n"] / prof_steps * n_tok
            nbytes = rows_per_step * cfg["d_model"] * 6.0
            hbm_peak, hbm_src = load_hbm_peak()
            gbs = nbytes / (rep["rmsnorm"]["ms"] / prof_steps * 1e-3) / 1e9
            roofline["hbm_kernels"] = [{"kernel": "rmsnorm_kernel", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                        "frac": gbs / hbm_peak, "bytes_per_step": nbytes, "launches_per_step": rep["rmsnorm"]["n"] / prof_steps,
                                        "peak_source": hbm_src}]
        if world == 1 and not args.no_cpu_baseline:
            try:   # the reference's own CPU path (transformers fp32 on the host cores) + full-size parity of the engine against it
                cpu_baseline, parity = cpu_reference_baseline(synthetic_weights(cfg, SEED), np.asarray(logits_dev), np.asarray(scores_dev))
            except Exception as exc:  # noqa: BLE001 - never lose the bench line to the baseline leg: fall back to the numpy port
                n_sample = 4
                cb, secs = cpu_oracle_docs_per_s(n_sample, repeats=1)
                cpu_baseline = {"value": cb, "unit": "docs/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{n_sample} of the {HITS} documents of one query (S={Q_LEN + P_LEN + 24}), {secs:.1f} s, numpy fp32 oracle on all host "
                                          f"threads (transformers CPU leg failed: {type(exc).__name__}: {exc})"}

    # parity at full size over ALL 100 documents of the headline query: the engine's logits against the reference's own fp32 logits
    # (transformers on CPU, computed in the build container by tests/golden/make_headline_query.py and committed); the cpu_baseline leg
    # above re-derives the first 32 of them live on this box (`live_check`)
    if rank == 0 and fixture_logits is not None:
        from b200rank import tolerance
        n_layers = cfg["num_layers"] + cfg["num_decoder_layers"]
        live = parity
        parity = dict(ordering_report(fixture_logits, np.asarray(logits_dev), n_layers),
                      against="tests/golden/headline_query.npz: fp32 transformers logits of all 100 documents of the headline query (committed fixture)",
                      tolerance=tolerance.describe(n_layers))
        if live is not None:
            # the live re-derivation covers the first 32 documents: its logit agreement is what it adds (top-k of a 32-document
            # subset of this query is not separated by construction and says nothing)
            parity["live_check"] = {k: live[k] for k in ("against", "docs", "max_abs_logit_diff", "mean_abs_logit_diff", "within_logit_tolerance",
                                                         "inversions_beyond_tolerance", "max_abs_score_diff", "reference_bf16_yardstick") if k in live}

    hf_cuda = None
    if rank == 0 and world == 1 and not args.no_hf_cuda:
        try:
            import torch
            from oracle import hf_cpu
            model, ids64, mask64 = cpu_reference_setup(synthetic_weights(cfg, SEED))
            v, secs, hf_logits = hf_cpu.timed_docs_per_s_on_device(model, ids64, mask64, YES_ID, NO_ID, 32, f"cuda:{local}", "bfloat16")
            hf_cuda = {"value": v, "unit": "docs/s", "ms_per_query": secs * 1e3, "dtype": "bf16", "batch_size": 32,
                       "what": f"transformers {__import__('transformers').__version__} T5ForConditionalGeneration on this GPU via torch {torch.__version__} "
                               "(cuBLAS GEMMs, eager attention), the reference's loop over 32/32/32/4 documents with ids resident on the device; wall clock, best of 3",
                       "engine_speedup": value / v,
                       "max_abs_logit_diff_vs_engine": float(np.abs(hf_logits - np.asarray(logits_dev)).max())}
            del model
            torch.cuda.empty_cache()
        except Exception as exc:  # noqa: BLE001 - informational leg
            hf_cuda = {"unavailable": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        line = {
            "metric": f"docs scored/sec ({MODEL} q32/p128, pointwise yes_no)", "value": value, "unit": "docs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world, queries_per_step=QPS), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "docs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": launches, "sustained": sustained, "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity, "api_text": api_text, "hf_cuda": hf_cuda,
            "weights_load_s": round(t_load, 2), "sample_scores": [float(x) for x in scores_dev[:4]],
        }
        print(json.dumps(line))
    barrier()
    eng.close()
    if use_dist:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pointwise", choices=["pointwise", "setwise", "pairwise", "qlm", "monot5"],
                    help="pointwise = the headline (BASELINE configs[1], default); setwise / pairwise = configs[2] / configs[3] through the text API, 1 GPU")
    ap.add_argument("--model", default=None, help="qlm / pairwise workloads: synthetic model shape (qlm: default flan-t5-large, BASELINE configs[4] is flan-t5-xxl; "
                                                   "pairwise: default flan-t5-xl)")
    ap.add_argument("--hits", type=int, default=0, help="qlm / pairwise workloads: documents per query (qlm: default 100, configs[4] says 1000; pairwise: default 24, configs[3] says 100)")
    ap.add_argument("--queries-per-step", type=int, default=int(os.environ.get("B200RANK_BENCH_QUERIES_PER_STEP", "2")),
                    help="headline workload: queries (100 hits each) merged into one device pass = one step; the first is the committed headline query")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-text-api", action="store_true", help="skip the informational strings -> rerank_many measurement")
    ap.add_argument("--no-pipeline", action="store_true", help="one batch in flight (wait right after submit) instead of two")
    ap.add_argument("--hf-cuda", action="store_true", help="(default since round 2; kept for old command lines)")
    ap.add_argument("--no-hf-cuda", action="store_true",
                    help="skip the on-GPU library comparator: the reference's own GPU path (the transformers model in bf16 on this GPU through torch / cuBLAS, "
                         "batch_size 32), timed after the engine has finished: an informational `hf_cuda` object, library kernels, never part of the product")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 3 s sustained loop")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and dist_env()[1] == 1:
        # convenience: re-launch under torchrun when asked for N > 1 directly (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.workload == "setwise":
        return run_setwise(args)
    if args.workload == "pairwise":
        return run_pairwise(args)
    if args.workload == "qlm":
        return run_qlm(args)
    if args.workload == "monot5":
        return run_monot5(args)
    return run_engine(args)


if __name__ == "__main__":
    sys.exit(main())
