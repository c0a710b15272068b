"""Generates tests/golden/golden_allpair_sweep.json: the reference's pairwise allpair rerank (llmrankers/pairwise.py:164-219, 279-290 —
2·C(n,2) prompts through Text2TextGenerationDataset + DataLoader batches, generate(max_new_tokens=2) per batch, exact-match verdicts,
win / conflict score aggregation in a defaultdict whose INSERTION order decides ties, documents that never score left to the
original-order tail, top-k assembly) over a grid of list sizes, batch sizes and k, with a deterministic stand-in for `self.llm.generate`
that answers from a hash of each row's non-pad token ids ('Passage A', 'Passage B', or something else).

    python tests/golden/make_golden_allpair_sweep.py      (build container only: needs /root/reference)

tests/test_host_logic.py replays the cases through llm-rankers_b200/llmrankers/pairwise.py with the same stand-in behind
T5Backend.generate_batches (imported from this module, which touches /root/reference only inside main())."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _h(*key) -> int:
    return int.from_bytes(hashlib.sha1(repr(key).encode()).digest()[:8], "big")


def stub_generate(rows, passage_id: int, a_id: int, b_id: int, junk_id: int, eos_id: int = 1, pad_id: int = 0):
    """[[pad, ▁Passage, t, </s>]] per prompt row, t chosen by a hash of the row without its pads: ▁A 45 %, ▁B 45 %, another token 10 %."""
    out = []
    for row in rows:
        row = [int(t) for t in row if int(t) != pad_id]
        u = _h(row) % 100
        out.append([pad_id, passage_id, a_id if u < 45 else (b_id if u < 90 else junk_id), eos_id])
    return out


def texts(n: int, seed: int):
    rng = np.random.default_rng(seed)
    return [" ".join(f"w{int(x)}" for x in rng.integers(0, 2000, int(rng.integers(1, 9)))) for _ in range(n)]


GRID = [(n, bs, k) for n in (0, 1, 2, 3, 4, 6, 9) for bs in (1, 2, 3, 4, 7, 16) for k in (1, 3, 10)]


def cases():
    for n, bs, k in GRID:
        yield dict(n=n, batch_size=bs, k=k, seed=_h(n, bs, k) % 10000, query="w11 w23 w5")


def main():
    ROOT = os.path.dirname(os.path.dirname(HERE))
    sys.path.insert(0, os.path.join(ROOT, "llm-rankers_b200"))
    import torch
    from types import SimpleNamespace
    from b200rank.synthetic import synthetic_tokenizer
    sys.path.insert(0, "/root/reference")
    from llmrankers.pairwise import PairwiseLlmRanker
    from llmrankers.rankers import SearchResult
    assert "/root/reference" in sys.modules["llmrankers.pairwise"].__file__
    tok = synthetic_tokenizer()
    passage_id = tok.encode("<pad> Passage", add_special_tokens=False)[1]
    a_id = tok.encode("<pad> Passage A", add_special_tokens=False)[-1]
    b_id = tok.encode("<pad> Passage B", add_special_tokens=False)[-1]
    junk_id = tok.encode("<pad> Passage C", add_special_tokens=False)[-1]
    assert tok.decode([0, passage_id, a_id, 1], skip_special_tokens=True) == "Passage A"

    class FakeLLM:
        device = "cpu"

        def __init__(self):
            self.batches = []

        def generate(self, input_ids, decoder_input_ids=None, max_new_tokens=2):
            assert max_new_tokens == 2 and decoder_input_ids.shape == (input_ids.shape[0], 2)
            self.batches.append([int(input_ids.shape[0]), int(input_ids.shape[1])])
            return torch.tensor(stub_generate(input_ids.tolist(), passage_id, a_id, b_id, junk_id))

    out = dict(passage_id=passage_id, a_id=a_id, b_id=b_id, junk_id=junk_id, cases=[])
    for c in cases():
        r = PairwiseLlmRanker.__new__(PairwiseLlmRanker)   # constructor needs the hub / accelerate
        r.tokenizer, r.llm, r.config = tok, FakeLLM(), SimpleNamespace(model_type="t5")
        r.device, r.method, r.batch_size, r.k = "cpu", "allpair", c["batch_size"], c["k"]
        r.prompt = """Given a query "{query}", which of the following two passages is more relevant to the query?

Passage A: "{doc1}"

Passage B: "{doc2}"

Output Passage A or Passage B:"""  # llmrankers/pairwise.py:42-48 (constructor bypassed, so restated here)
        r.decoder_input_ids = tok.encode("<pad> Passage", return_tensors="pt", add_special_tokens=False).repeat(c["batch_size"], 1)
        r.total_compare = r.total_completion_tokens = r.total_prompt_tokens = 0
        docs = [SearchResult(docid=f"d{i}", score=float(c["n"] - i), text=t) for i, t in enumerate(texts(c["n"], c["seed"]))]
        rec = dict(c, texts=[d.text for d in docs])
        try:
            res = r.rerank(c["query"], docs)
            rec.update(result=[[d.docid, d.score] for d in res], batches=r.llm.batches, total_compare=int(r.total_compare),
                       total_prompt_tokens=int(r.total_prompt_tokens), total_completion_tokens=int(r.total_completion_tokens))
        except Exception as e:   # noqa: BLE001 - the exception type IS the recorded behaviour
            rec["raises"] = type(e).__name__
        out["cases"].append(rec)
    path = os.path.join(HERE, "golden_allpair_sweep.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(f"wrote {len(out['cases'])} cases to {path}, {os.path.getsize(path)} bytes; raising:",
          sorted({(c['n'], c['raises']) for c in out['cases'] if 'raises' in c}))


if __name__ == "__main__":
    main()
