"""Golden fixtures for the listwise widening: the reference's own `ListwiseLlmRanker` (llmrankers/listwise.py:202-291,
T5 branch; sliding-window `rerank` inherited from `OpenAiListwiseLlmRanker`, :177-195) driven over a live transformers fp32 model.

    PYTHONPATH=/root/reference python tests/golden/make_golden_listwise.py

Build container only (needs /root/reference). Constructor bypassed with `__new__` as in make_golden.py (hub download, accelerate,
`batch_encode_plus` removed in transformers 5). Same tiny model / documents as golden_tiny (seed 1234, docs12), plus a
digit-favouring lm_head for the generation mode so that the free-form output actually contains passage numbers.
Writes golden_listwise.npz + golden_listwise_meta.json.
"""
import copy
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import TINY_VOCAB, Recorder, build_hf_model, counters, make_docs, save_calls  # noqa: E402  (also sets sys.path)

from b200rank.synthetic import model_cfg, synthetic_tokenizer, synthetic_weights  # noqa: E402
from llmrankers.listwise import ListwiseLlmRanker  # noqa: E402  (the reference)


def listwise_ranker(tok, rec, cfg, window_size, step_size, num_repeat, scoring):
    r = ListwiseLlmRanker.__new__(ListwiseLlmRanker)
    r.tokenizer, r.llm, r.config = tok, rec, cfg
    r.device, r.window_size, r.step_size, r.num_repeat, r.scoring = "cpu", window_size, step_size, num_repeat, scoring
    r.decoder_input_ids = tok.encode("<pad> Passage", return_tensors="pt", add_special_tokens=False)
    r.target_token_ids = tok([f"<pad> Passage {c}" for c in ListwiseLlmRanker.CHARACTERS], return_tensors="pt",
                             add_special_tokens=False, padding=True).input_ids[:, -1]
    r.total_compare = r.total_completion_tokens = r.total_prompt_tokens = 0
    return r


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    tok = synthetic_tokenizer()
    cfg = model_cfg("t5-tiny", TINY_VOCAB)
    seed = 1234
    weights = synthetic_weights(cfg, seed)
    model, hf_cfg = build_hf_model(cfg, weights)
    query = "w11 w23 w5 w42 w8"
    docs12 = make_docs(np.random.default_rng(8), 12, 5, 12)     # == golden_meta.json tiny.docs12
    # generation: x30 on the lm_head rows of the digit pieces and of '>' so that greedy decoding emits identifiers
    digit_ids = sorted({i for ch in "123456789" for i in tok.encode(ch, add_special_tokens=False)} |
                       {i for ch in "123456789" for i in tok.encode("x" + ch, add_special_tokens=False)[-1:]})
    w_dig = dict(weights)
    w_dig["lm_head.weight"] = weights["lm_head.weight"].copy()
    w_dig["lm_head.weight"][digit_ids] *= 30.0
    model_dig, _ = build_hf_model(cfg, w_dig)

    meta = {"transformers": __import__("transformers").__version__, "torch": torch.__version__, "model": "t5-tiny", "vocab_size": TINY_VOCAB,
            "seed": seed, "query": query, "docs12": [dict(docid=d.docid, score=d.score, text=d.text) for d in docs12],
            "digit_ids": digit_ids, "digit_boost": 30.0, "cases": {}}
    store = {}
    for name, scoring, mdl, ws, ss, rep in (("listwise_lik", "likelihood", model, 4, 2, 1),
                                            ("listwise_lik_rep2", "likelihood", model, 5, 3, 2),
                                            ("listwise_gen", "generation", model_dig, 4, 2, 1),
                                            ("listwise_gen_plain", "generation", model, 6, 3, 1)):
        rec = Recorder(mdl)
        r = listwise_ranker(tok, rec, hf_cfg, ws, ss, rep, scoring)
        compares = []
        orig = r.compare

        def spy(q, ds, _orig=orig, _log=compares):
            out = _orig(q, ds)
            _log.append(dict(docids=[d.docid for d in ds], output=out))
            return out
        r.compare = spy
        out = r.rerank(query, copy.deepcopy(docs12))
        n = save_calls(store, name, rec.calls, cols=None)
        meta["cases"][name] = dict(n_calls=n, order=[d.docid for d in out], scores=[d.score for d in out], window_size=ws, step_size=ss,
                                   num_repeat=rep, scoring=scoring, digit_favouring=mdl is model_dig, compares=compares, **counters(r))
        print(name, meta["cases"][name]["order"], counters(r), [c["output"] for c in compares][:3])
    # the reference's response parser on its own (listwise.py:110-144): well-formed, duplicated, out-of-range and junk responses
    from llmrankers.listwise import create_permutation_instruction_complete, receive_permutation
    from llmrankers.rankers import SearchResult
    rng = np.random.default_rng(11)
    perm_cases = []
    for resp, n, a, b in (("[2] > [1] > [3]", 6, 0, 3), ("[3] > [3] > [1]", 6, 2, 6), ("[9] > [2] > [0] > [1]", 5, 0, 4), ("no numbers here", 4, 0, 4),
                          ("2>1", 4, 1, 4), ("[4]>[3]>[2]>[1]", 4, 0, 4), ("12 1 2", 12, 0, 12), ("[1] > [2]", 3, 2, 10), ("", 3, 0, 3)):
        perm_cases.append((resp, n, a, b))
    for _ in range(200):
        n = int(rng.integers(2, 24)); a = int(rng.integers(0, n)); b = int(rng.integers(a + 1, n + 3))
        w = min(b, n) - a
        nums = rng.integers(0, w + 3, size=int(rng.integers(0, w + 4)))
        sep = [" > ", ">", ", ", " ", "] > [", " then "][int(rng.integers(0, 6))]
        perm_cases.append((sep.join(f"[{int(x)}]" for x in nums) + ["", ".", " done 7", "\n"][int(rng.integers(0, 4))], n, a, b))
    meta["receive_permutation"] = []
    for resp, n, a, b in perm_cases:
        ranking = [SearchResult(docid=f"p{i}", score=0.0, text="") for i in range(n)]
        out = receive_permutation(ranking, resp, a, b)
        meta["receive_permutation"].append(dict(response=resp, n=n, rank_start=a, rank_end=b, order=[d.docid for d in out]))
    meta["instruction_complete"] = dict(docs=[" Title: Content: alpha  beta\tgamma ", "delta " * 305, "x"], query="q one",
                                        text=create_permutation_instruction_complete("q one", [SearchResult(docid=str(i), score=0.0, text=t) for i, t in
                                                                                               enumerate([" Title: Content: alpha  beta\tgamma ", "delta " * 305, "x"])]))
    meta["target_token_ids"] = r.target_token_ids.tolist()
    meta["decoder_prefix"] = r.decoder_input_ids[0].tolist()
    np.savez_compressed(os.path.join(HERE, "golden_listwise.npz"), **store)
    with open(os.path.join(HERE, "golden_listwise_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    for fn in ("golden_listwise.npz", "golden_listwise_meta.json"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")


if __name__ == "__main__":
    main()
