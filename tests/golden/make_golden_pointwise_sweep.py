"""Generates tests/golden/golden_pointwise_sweep.json: the reference's pointwise rankers (llmrankers/pointwise.py:36-130 yes_no / qlm,
:136-186 MonoT5) — prompt templating, Text2TextGenerationDataset, DataLoader + DataCollatorWithPadding batching, counters, score
read-out, stable sort — run over a grid of list sizes and batch sizes with a deterministic stand-in for `self.llm` whose logits are a
hash of each row's REAL (unpadded) token ids. A score therefore matches only if the drop-in class fed the engine exactly the token
rows the reference fed the model: the sweep pins prompt assembly in situ (odd whitespace, newlines, empty passages, the shared query),
the per-batch counters (short last batch, batch_size > n, n = 1) and the final order for sizes the model-driven fixtures do not reach.

    python tests/golden/make_golden_pointwise_sweep.py      (build container only: needs /root/reference)

tests/test_host_logic.py replays the cases through llm-rankers_b200/llmrankers/pointwise.py with the same stand-in behind the backend
(imported from this module, which touches /root/reference only inside main())."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
VOCAB = 6400                     # >= 6137: MonoT5 reads logits 6136 / 1176 (pointwise.py:176)


def _h(*key) -> int:
    return int.from_bytes(hashlib.sha1(repr(key).encode()).digest()[:8], "big")


def _unit(h: int) -> float:
    return ((h % 100003) / 100003.0 - 0.5) * 8.0


def yes_no_logits(row, col_a: int, col_b: int):
    """(logit[col_a], logit[col_b]) of the stand-in model for one unpadded prompt row."""
    row = [int(t) for t in row]
    return _unit(_h("a", row)), _unit(_h("b", row))


def qlm_label_logits(row, labels):
    """Per label position t: the stand-in model's logit at labels[t] (every other vocabulary entry is 0)."""
    row = [int(t) for t in row]
    return [_unit(_h("q", row, t, int(l))) for t, l in enumerate(labels)]


def full_logits_yes_no(rows, col_a, col_b):
    out = np.zeros((len(rows), 1, VOCAB), np.float32)
    for i, r in enumerate(rows):
        a, b = yes_no_logits(r, col_a, col_b)
        out[i, 0, col_a], out[i, 0, col_b] = a, b
    return out


def full_logits_qlm(rows, labels):
    out = np.zeros((len(rows), len(labels), VOCAB), np.float32)
    for i, r in enumerate(rows):
        for t, v in enumerate(qlm_label_logits(r, labels)):
            out[i, t, int(labels[t])] = v
    return out


def texts(n: int, seed: int):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        words = [f"w{int(x)}" for x in rng.integers(0, 2000, int(rng.integers(0, 14)))]
        kind = i % 5
        if kind == 1:
            t = "  ".join(words) + " "            # runs of spaces, trailing space
        elif kind == 2:
            t = "\n".join(words)                  # newlines collapse in the T5 pre-tokeniser
        elif kind == 3:
            t = "\t" + " ".join(words) + "\n\n"
        else:
            t = " ".join(words)
        out.append(t)
    return out


GRID = [(n, bs) for n in (0, 1, 2, 3, 7, 10, 33) for bs in (1, 2, 4, 32)]


def cases():
    for kind in ("yes_no", "qlm", "monot5"):
        for n, bs in GRID:
            yield dict(kind=kind, n=n, batch_size=bs, seed=_h(kind, n, bs) % 10000, query="w11  w23\nw5 w42 w8" if n % 2 else "w7 w9")


def main():
    ROOT = os.path.dirname(os.path.dirname(HERE))
    sys.path.insert(0, os.path.join(ROOT, "llm-rankers_b200"))
    import torch
    from types import SimpleNamespace
    from b200rank.synthetic import synthetic_tokenizer
    sys.path.insert(0, "/root/reference")
    from llmrankers.pointwise import MonoT5LlmRanker, PointwiseLlmRanker
    from llmrankers.rankers import SearchResult
    assert "/root/reference" in sys.modules["llmrankers.pointwise"].__file__
    tok = synthetic_tokenizer()
    yes_id, no_id = tok.encode("Yes", add_special_tokens=False)[0], tok.encode("No", add_special_tokens=False)[0]

    class FakeLLM:
        device = "cpu"
        config = SimpleNamespace(decoder_start_token_id=0)

        def __init__(self, kind):
            self.kind, self.batches = kind, []

        def __call__(self, input_ids=None, attention_mask=None, decoder_input_ids=None, labels=None):
            rows = [[int(t) for t, m in zip(r, mk) if m] for r, mk in zip(input_ids.tolist(), attention_mask.tolist())]
            self.batches.append([len(rows), int(input_ids.shape[1])])
            if self.kind == "qlm":
                assert all(l == labels[0].tolist() for l in labels.tolist())
                lg = full_logits_qlm(rows, labels[0].tolist())
            elif self.kind == "monot5":
                assert decoder_input_ids.tolist() == [[0]] * len(rows)
                lg = full_logits_yes_no(rows, 1176, 6136)
            else:
                assert decoder_input_ids.tolist() == [[0]] * len(rows)
                lg = full_logits_yes_no(rows, yes_id, no_id)
            return SimpleNamespace(logits=torch.from_numpy(lg))

    out = dict(yes_id=yes_id, no_id=no_id, cases=[])
    for c in cases():
        cls = MonoT5LlmRanker if c["kind"] == "monot5" else PointwiseLlmRanker
        r = cls.__new__(cls)                               # constructor needs the hub / accelerate
        r.tokenizer, r.llm, r.config = tok, FakeLLM(c["kind"]), SimpleNamespace(model_type="t5")
        r.device, r.method, r.batch_size = "cpu", ("qlm" if c["kind"] == "qlm" else "yes_no"), c["batch_size"]
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
    path = os.path.join(HERE, "golden_pointwise_sweep.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(f"wrote {len(out['cases'])} cases to {path}, {os.path.getsize(path)} bytes; raising:",
          [(c['kind'], c['n'], c['batch_size'], c['raises']) for c in out['cases'] if 'raises' in c])


if __name__ == "__main__":
    main()
