"""Builds the HEADLINE query of bench.py / the full-size GPU parity test: 100 documents of the BASELINE configs[1] shape whose
reference ordering is WELL-POSED at the top ("pick/record inputs whose gaps exceed tolerance").

    python tests/golden/make_headline_query.py [--pool 3000]          (CPU, ~5 min per 1000 pool documents on 8 cores)

Why. With seeded random weights the (yes - no) margins of 100 random documents have a standard deviation of ~1.2 and the adjacent
gaps among the reference's top-11 are ~0.02 (tests/golden/golden_headline_meta.json, twelve ids seeds) — smaller than what bf16
operands can resolve (b200rank/tolerance.py: 0.12 + 0.03|x| per logit at 48 layers; observed max 0.083). "Identical top-10" is then
a coin flip for ANY reduced-precision implementation, including the reference's own fp16/bf16 CUDA path. The ordering claim becomes
testable on a query whose top-11 reference margins are further apart than the tolerance allows them to move.

How. One query (32 ids) and a pool of random passages (128 ids each) of exactly the bench shape are scored by the reference's own
arithmetic (transformers fp32 on CPU through oracle/hf_cpu.py, as llmrankers/pointwise.py:117-124). From the pool the script keeps
  * 11 "head" documents: walking down from the pool's best margin, each next one at least GAP = 2 x (ATOL + RTOL*|x|) below the
    previous (GAP evaluated with the larger |logit| of the pair) — these are the reference's ranks 1..11, and
  * 89 "tail" documents drawn at random (seeded) from the pool documents at least GAP below head document #11,
then shuffles the 100 (seeded) so the input order carries no information. Only the fp32 reference decides the selection; no engine
output is read. The result (token ids, lengths, reference logits, margins, order) is committed as tests/golden/headline_query.npz
+ headline_query_meta.json (~80 KB) and is what bench.py times and what tests/test_engine_gpu.py::test_headline_query_parity checks.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "llm-rankers_b200"))
sys.path.insert(0, ROOT)

MODEL, HITS, Q_LEN, P_LEN, SEED = "flan-t5-large", 100, 32, 128, 929


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pool", type=int, default=3000)
    ap.add_argument("--out", default=os.path.join(HERE, "headline_query"))
    ap.add_argument("--cache", default="/tmp/headline_pool.npz", help="pool logits cache (not committed)")
    args = ap.parse_args()
    from b200rank.synthetic import NO_ID, TPL_A, TPL_B, TPL_C, YES_ID, model_cfg, synthetic_weights
    from b200rank.tolerance import logit_tolerance
    from oracle import hf_cpu
    cfg = model_cfg(MODEL)
    n_layers = cfg["num_layers"] + cfg["num_decoder_layers"]
    rng = np.random.default_rng(SEED)
    query = rng.integers(3, 32000, size=Q_LEN).tolist()                     # the same draw order as synthetic_prompt_ids(seed=929)
    passages = rng.integers(3, 32000, size=(args.pool, P_LEN))
    ids = np.array([TPL_A + p.tolist() + TPL_B + query + TPL_C + [1] for p in passages], np.int64)
    S = ids.shape[1]
    assert S == Q_LEN + P_LEN + 24
    if os.path.exists(args.cache) and np.load(args.cache)["logits"].shape[0] == args.pool:
        logits = np.load(args.cache)["logits"]
    else:
        t0 = time.time()
        model = hf_cpu.build_model(cfg, synthetic_weights(cfg, SEED))
        mask = np.ones_like(ids)
        parts = []
        for b0 in range(0, args.pool, 96):
            lg, _ = hf_cpu.score_yes_no(model, ids[b0:b0 + 96], mask[b0:b0 + 96], YES_ID, NO_ID, 32)
            parts.append(lg)
            print(f"{b0 + len(lg)}/{args.pool} pool documents, {time.time() - t0:.0f} s", flush=True)
        logits = np.concatenate(parts, 0).astype(np.float32)
        np.savez_compressed(args.cache, logits=logits)
    m = (logits[:, 0] - logits[:, 1]).astype(np.float64)
    tol = logit_tolerance(logits, n_layers).max(1)                           # the looser of a document's two logit bounds
    by_margin = np.argsort(-m, kind="stable")
    head = [int(by_margin[0])]
    for i in by_margin[1:]:
        gap = 2.0 * max(tol[head[-1]], tol[i])
        if m[head[-1]] - m[i] >= gap:
            head.append(int(i))
            if len(head) == 11:
                break
    if len(head) < 11:
        raise SystemExit(f"pool of {args.pool} too small: only {len(head)} head documents with the required gaps (margin range {m.min():.2f}..{m.max():.2f})")
    last = head[-1]
    tail_pool = [int(i) for i in by_margin if m[last] - m[i] >= 2.0 * max(tol[last], tol[i]) and int(i) not in head]
    sel_rng = np.random.default_rng(SEED + 1)
    tail = sel_rng.choice(np.array(tail_pool), size=HITS - 11, replace=False).tolist()
    chosen = np.array(head + tail)
    chosen = chosen[sel_rng.permutation(HITS)]
    q_ids = ids[chosen].astype(np.int32)
    q_logits = logits[chosen]
    q_m = m[chosen]
    scores = np.exp(q_logits[:, 0]) / np.exp(q_logits).sum(1)
    order = np.argsort(-scores, kind="stable")
    top = order[:11]
    gaps = q_m[top[:-1]] - q_m[top[1:]]
    need = 2.0 * np.maximum(tol[chosen][top[:-1]], tol[chosen][top[1:]])
    assert (gaps >= need - 1e-9).all(), (gaps, need)
    srt = np.sort(q_m)[::-1]
    meta = {"model": MODEL, "weights_seed": SEED, "query_and_pool_seed": SEED, "pool": args.pool, "hits": HITS, "q_len": Q_LEN, "p_len": P_LEN, "S": int(S),
            "reference": "transformers %s T5ForConditionalGeneration fp32 on CPU via oracle/hf_cpu.py (llmrankers/pointwise.py:117-124)" % __import__("transformers").__version__,
            "selection": "11 head documents with adjacent reference-margin gaps >= 2*(ATOL+RTOL*|x|), 89 seeded random tail documents at least that far below head #11, shuffled",
            "tolerance": "b200rank/tolerance.py at %d layers" % n_layers,
            "pool_margin_std": float(m.std()), "pool_margin_range": [float(m.min()), float(m.max())],
            "top11_adjacent_gaps": [float(x) for x in gaps], "top11_required_gaps": [float(x) for x in need],
            "min_adjacent_gap_all_100": float(np.min(srt[:-1] - srt[1:])), "max_abs_logit": float(np.abs(q_logits).max()),
            "order": [int(x) for x in order]}
    np.savez_compressed(args.out + ".npz", ids=q_ids, lengths=np.full((HITS,), S, np.int32), ref_logits=q_logits.astype(np.float32))
    with open(args.out + "_meta.json", "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps({k: meta[k] for k in ("pool_margin_std", "pool_margin_range", "top11_adjacent_gaps", "top11_required_gaps", "min_adjacent_gap_all_100")}, indent=1))


if __name__ == "__main__":
    main()
