"""Generates tests/golden/golden_setwise_perm.json: the reference's permutation-voting compare (llmrankers/setwise.py:102-157,
num_permutation > 1: the passages AND the labels of a compare are shuffled num_permutation times with the module RNG, the prompts go
through generate() as ONE batch, the winners are voted, ties broken with random.choice), plus whole rerank() runs on top of it.

What this branch adds over the single-prompt compare is host logic — RNG consumption, prompt assembly under shuffled labels, the vote,
its rejection rules and its tie-break — so `self.llm.generate` is a deterministic stand-in here (`stub_generate`: a pure function of
each prompt's token ids that answers with one of the labels present in the prompt, sometimes with a label that is not, sometimes with
the label twice — which the reference rejects, as it does for every answer of a random-init model). The model arithmetic under
generate() is pinned by the other fixtures. The reference tokenises the shuffled prompts without padding (:125), so it only works when
they are equally long: every passage here has the same number of single-token words.

    python tests/golden/make_golden_setwise_perm.py      (build container only: needs /root/reference)

tests/test_host_logic.py replays the fixtures through llm-rankers_b200/llmrankers with `random.seed(929)` (setwise.py:18) and the same
stand-in behind the backend; it imports `stub_generate` from this module (which touches /root/reference only inside main())."""
import copy
import hashlib
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def stub_generate(rows, passage_id: int, label_token_ids, eos_id: int = 1, pad_id: int = 0):
    """[[pad, ▁Passage, t1, t2]] per prompt row: t1 = a label token chosen by a hash of the row, t2 = </s>; ~20 % of the rows answer
    with the label twice (decodes to 'X X': rejected), ~10 % with a label that is not in the prompt (rejected)."""
    out = []
    for row in rows:
        row = [int(t) for t in row if int(t) != pad_id]
        present = [row[i + 1] for i in range(len(row) - 1) if row[i] == passage_id and row[i + 1] in label_token_ids]
        h = int.from_bytes(hashlib.sha1(repr(row).encode()).digest()[:8], "big")
        u = (h >> 24) % 10
        if u < 2:
            t = present[h % len(present)]
            out.append([pad_id, passage_id, t, t])
        elif u < 3:
            absent = [t for t in label_token_ids if t not in present]
            out.append([pad_id, passage_id, absent[h % len(absent)], eos_id])
        else:
            out.append([pad_id, passage_id, present[h % len(present)], eos_id])
    return out


def main():
    ROOT = os.path.dirname(os.path.dirname(HERE))
    sys.path.insert(0, os.path.join(ROOT, "llm-rankers_b200"))
    import torch
    from types import SimpleNamespace
    from b200rank.synthetic import synthetic_tokenizer
    sys.path.insert(0, "/root/reference")
    from llmrankers.rankers import SearchResult                    # the reference
    from llmrankers.setwise import SetwiseLlmRanker                # the reference
    assert "/root/reference" in sys.modules["llmrankers.setwise"].__file__

    tok = synthetic_tokenizer()
    passage_id = tok.encode("<pad> Passage", add_special_tokens=False)[1]
    label_ids = [tok.encode(f"<pad> Passage {c}", add_special_tokens=False)[-1] for c in SetwiseLlmRanker.CHARACTERS]

    class StubLLM:
        device = "cpu"

        def __init__(self):
            self.calls = []

        def generate(self, input_ids, decoder_input_ids=None, max_new_tokens=2):
            assert max_new_tokens == 2 and decoder_input_ids.shape == (input_ids.shape[0], 2)
            rows = input_ids.tolist()
            out = stub_generate(rows, passage_id, label_ids)
            self.calls.append(dict(input_ids=rows, output=out))
            return torch.tensor(out)

    def ranker(num_child, k, method, num_perm):
        r = SetwiseLlmRanker.__new__(SetwiseLlmRanker)   # constructor needs the hub / accelerate
        r.tokenizer, r.llm, r.config = tok, StubLLM(), SimpleNamespace(model_type="t5")
        r.device, r.num_child, r.k, r.scoring, r.method, r.num_permutation = "cpu", num_child, k, "generation", method, num_perm
        r.decoder_input_ids = tok.encode("<pad> Passage", return_tensors="pt", add_special_tokens=False)
        r.total_compare = r.total_completion_tokens = r.total_prompt_tokens = 0
        return r

    def counters(r):
        return dict(total_compare=int(r.total_compare), total_prompt_tokens=int(r.total_prompt_tokens),
                    total_completion_tokens=int(r.total_completion_tokens))

    rng = np.random.default_rng(21)
    docs = [SearchResult(docid=f"d{i}", score=float(30 - i), text=" ".join(f"w{int(x)}" for x in rng.integers(0, 2000, 7))) for i in range(30)]
    query = "w11 w23 w5 w42 w8"
    out = dict(query=query, passage_id=passage_id, label_token_ids=label_ids,
               docs=[dict(docid=d.docid, score=d.score, text=d.text) for d in docs], compares=[], reranks=[])
    sets = [[0, 1, 2, 3], [4, 5], [6, 7, 8], [9, 10, 11, 0], [3, 2, 1, 0], [5, 7, 9, 11], [12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22], [1, 2]]
    for num_perm in (2, 3, 5, 8):
        r = ranker(3, 3, "heapsort", num_perm)
        random.seed(929)                                   # setwise.py:18 / run.py:16
        for s in sets:
            n0 = len(r.llm.calls)
            label = r.compare(query, [docs[i] for i in s])
            call = r.llm.calls[n0]
            out["compares"].append(dict(num_permutation=num_perm, docs=s, label=label, input_ids=call["input_ids"], output=call["output"],
                                        **counters(r)))
    for n, num_child, k, method, num_perm in ((12, 3, 3, "heapsort", 3), (30, 10, 10, "heapsort", 4), (12, 3, 3, "bubblesort", 3),
                                              (30, 4, 5, "bubblesort", 2), (5, 3, 10, "heapsort", 5)):
        r = ranker(num_child, k, method, num_perm)
        random.seed(929)
        rec = dict(n=n, num_child=num_child, k=k, method=method, num_permutation=num_perm)
        try:
            res = r.rerank(query, copy.deepcopy(docs[:n]))
            rec.update(order=[d.docid for d in res], scores=[d.score for d in res])
        except Exception as e:   # noqa: BLE001 - the exception type IS the recorded behaviour
            rec["raises"] = type(e).__name__
        rec.update(n_generate_calls=len(r.llm.calls), **counters(r))
        out["reranks"].append(rec)
    path = os.path.join(HERE, "golden_setwise_perm.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes;", [c["label"] for c in out["compares"]])
    print([(r.get("order", r.get("raises")), r["total_compare"]) for r in out["reranks"]])


if __name__ == "__main__":
    main()
