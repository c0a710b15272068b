"""Full-size parity fixture for the HEADLINE workload (BASELINE configs[1]): all 100 documents of the bench query on flan-t5-large.

    python tests/golden/make_golden_headline.py [--ids-seeds 929 930 ...]      (CPU, ~1-2 min per ids seed on 8 cores)

Runs the reference's own arithmetic — `transformers` T5ForConditionalGeneration in fp32 on the CPU, called as
llmrankers/pointwise.py:117-124 does (oracle/hf_cpu.py, pinned against the reference's rerank() by tests/test_oracle_golden.py) — on
the exact token ids and seeded weights bench.py uses, and writes `golden_headline.npz` / `golden_headline_meta.json`:
the (yes, no) logits of all 100 documents, the reference order, and the margin statistics that make the ordering claim well-posed
("pick/record seeds whose gaps exceed tolerance; report min adjacent gap + tau").

Seed choice. A document's score is softmax(yes, no)[0], a monotone function of its margin m = yes - no. The engine computes in bf16
with fp32 accumulation; its logits carry |err| <= ATOL + RTOL*|x| (frozen in b200rank/tolerance.py), so a margin can move by at most
2*(ATOL + RTOL*|x|) and a PAIR of documents can swap only if their reference margins are closer than the sum of both bounds. With
random-init weights and random token ids some of the 4950 pairs of a 100-document query are always closer than that; whether the
top-10 *set and order* is well separated depends on the query. The script evaluates the candidate ids seeds, records for each the
smallest adjacent margin gap among the reference's top-11 (what decides "identical top-10"), and marks as `headline` the first seed
(starting from bench.py's historical 929) whose top-11 gaps all exceed the pair bound. Weights seed stays 929. Nothing here reads
an engine result: the choice is a property of the fp32 reference alone.
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

MODEL, HITS, Q_LEN, P_LEN, WEIGHT_SEED = "flan-t5-large", 100, 32, 128, 929


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ids-seeds", type=int, nargs="+", default=[929, 930, 931, 932, 933, 934])
    ap.add_argument("--out", default=os.path.join(HERE, "golden_headline"))
    args = ap.parse_args()
    from b200rank.synthetic import NO_ID, YES_ID, model_cfg, synthetic_prompt_ids, synthetic_weights
    from b200rank.tolerance import logit_tolerance
    from oracle import hf_cpu
    cfg = model_cfg(MODEL)
    t0 = time.time()
    model = hf_cpu.build_model(cfg, synthetic_weights(cfg, WEIGHT_SEED))
    print(f"model built in {time.time() - t0:.0f} s", flush=True)
    n_layers = cfg["num_layers"] + cfg["num_decoder_layers"]
    arrays, meta = {}, {"model": MODEL, "weights_seed": WEIGHT_SEED, "hits": HITS, "q_len": Q_LEN, "p_len": P_LEN,
                        "reference": "transformers %s T5ForConditionalGeneration fp32 on CPU via oracle/hf_cpu.py (llmrankers/pointwise.py:117-124)"
                                     % __import__("transformers").__version__,
                        "seeds": {}}
    headline = None
    for seed in args.ids_seeds:
        ids, lengths = synthetic_prompt_ids(HITS, Q_LEN, P_LEN, seed=seed)
        mask = (np.arange(ids.shape[1])[None] < lengths[:, None]).astype(np.int64)
        t0 = time.time()
        logits, scores = hf_cpu.score_yes_no(model, ids.astype(np.int64), mask, YES_ID, NO_ID, 32)
        dt = time.time() - t0
        m = logits[:, 0] - logits[:, 1]
        order = np.argsort(-scores, kind="stable")
        tol = logit_tolerance(logits, n_layers)                # per-logit bound
        mtol = tol.sum(1)                                      # a margin moves by at most the bound of both of its logits
        top = order[:11]
        gaps = m[top[:-1]] - m[top[1:]]                        # adjacent gaps among the top-11 (>= 0)
        pair_bound = mtol[top[:-1]] + mtol[top[1:]]
        all_sorted = np.sort(m)[::-1]
        info = {"seconds": round(dt, 1), "margin_min": float(m.min()), "margin_max": float(m.max()), "margin_std": float(m.std()),
                "max_abs_logit": float(np.abs(logits).max()),
                "top11_adjacent_gaps": [float(x) for x in gaps], "top11_pair_bounds": [float(x) for x in pair_bound],
                "top10_separated": bool((gaps > pair_bound).all()),
                "min_adjacent_gap_all_100": float(np.min(all_sorted[:-1] - all_sorted[1:])),
                "order": [int(x) for x in order]}
        meta["seeds"][str(seed)] = info
        arrays[f"logits_{seed}"] = logits.astype(np.float32)
        print(f"ids seed {seed}: {dt:.0f} s, margin std {info['margin_std']:.3f}, top-11 min gap {gaps.min():.4f} vs pair bound {pair_bound.max():.4f} "
              f"-> separated {info['top10_separated']}", flush=True)
        if headline is None and info["top10_separated"]:
            headline = seed
    meta["headline_ids_seed"] = headline
    np.savez_compressed(args.out + ".npz", **arrays)
    with open(args.out + "_meta.json", "w") as f:
        json.dump(meta, f, indent=1)
    print("headline ids seed:", headline)


if __name__ == "__main__":
    main()
