"""Generates the golden fixtures in this directory by running THE REFERENCE ITSELF (ielab/llm-rankers at
/root/reference, unmodified) on top of a live `transformers` T5ForConditionalGeneration in fp32 on CPU.

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

Only runs in the build container (needs /root/reference); the fixtures it writes are committed and are what the
GPU box sees. The reference constructors cannot run offline (hub download, accelerate's device_map='auto',
tokenizer.batch_encode_plus removed in transformers 5) — SURVEY.md §8c — so rankers are created with `__new__`
and the attributes their constructors would set are injected; every method that runs afterwards
(rerank / compare / heapify / heapSort / truncate) is the reference's own code.

Weights come from b200rank.synthetic.synthetic_weights (numpy PCG64, machine-independent), so tests rebuild the
identical model from (model name, vocab size, seed) without committing weight files.
"""
import copy
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "llm-rankers_b200"))
sys.path.insert(0, "/root/reference")

from transformers import T5Config, T5ForConditionalGeneration  # noqa: E402
from transformers.models.t5.modeling_t5 import T5Attention  # noqa: E402

from b200rank.synthetic import LABELS, model_cfg, synthetic_tokenizer, synthetic_weights  # noqa: E402
from llmrankers.pairwise import PairwiseLlmRanker  # noqa: E402  (the reference)
from llmrankers.pointwise import PointwiseLlmRanker  # noqa: E402
from llmrankers.rankers import SearchResult  # noqa: E402
from llmrankers.setwise import SetwiseLlmRanker  # noqa: E402

TINY_VOCAB = 2304


def build_hf_model(cfg, weights):
    hf_cfg = T5Config(vocab_size=cfg["vocab_size"], d_model=cfg["d_model"], d_kv=64, d_ff=cfg["d_ff"],
                      num_layers=cfg["num_layers"], num_decoder_layers=cfg["num_decoder_layers"], num_heads=cfg["num_heads"],
                      feed_forward_proj="gated-gelu", tie_word_embeddings=False, decoder_start_token_id=0)
    model = T5ForConditionalGeneration(hf_cfg).eval()
    # transformers 5.x ties lm_head to shared for a freshly constructed model even with tie_word_embeddings=False
    # (real Flan-T5 checkpoints carry a distinct lm_head.weight and stay untied): break the tie explicitly.
    model.lm_head.weight = torch.nn.Parameter(torch.empty_like(model.shared.weight))
    model.config.tie_word_embeddings = False
    sd = {k: torch.from_numpy(v.copy()) for k, v in weights.items()}
    sd["encoder.embed_tokens.weight"] = sd["shared.weight"]
    sd["decoder.embed_tokens.weight"] = sd["shared.weight"]
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("embed_tokens" in m for m in missing), missing
    # lm_head must be the separate tensor we loaded (untied)
    assert torch.equal(model.lm_head.weight, sd["lm_head.weight"])
    assert torch.equal(model.shared.weight, sd["shared.weight"])
    assert model.encoder.embed_tokens.weight is model.shared.weight and model.decoder.embed_tokens.weight is model.shared.weight
    assert model.config.scale_decoder_outputs is False
    return model, hf_cfg


class Recorder:
    """Stands where `self.llm` is; forwards to the HF model and records what crossed the boundary."""

    def __init__(self, model):
        self.m = model
        self.device = model.device
        self.config = model.config
        self.calls = []

    def __call__(self, **kw):
        with torch.no_grad():
            out = self.m(**kw)
        self.calls.append(dict(kind="forward", inputs={k: v.clone() for k, v in kw.items()}, logits=out.logits.clone()))
        return out

    def generate(self, input_ids, **kw):
        with torch.no_grad():
            out = self.m.generate(input_ids, **kw)
        self.calls.append(dict(kind="generate", input_ids=input_ids.clone(), output=out.clone()))
        return out


def make_docs(rng, n, lo, hi, n_words=2000):
    docs = []
    for i in range(n):
        L = int(rng.integers(lo, hi + 1))
        docs.append(SearchResult(docid=f"d{i}", score=float(n - i), text=" ".join(f"w{int(x)}" for x in rng.integers(0, n_words, L))))
    return docs


def pointwise_ranker(tok, rec, cfg, method, batch_size):
    r = PointwiseLlmRanker.__new__(PointwiseLlmRanker)
    r.tokenizer, r.llm, r.config = tok, rec, cfg
    r.device, r.method, r.batch_size = "cpu", method, batch_size
    r.total_compare = r.total_completion_tokens = r.total_prompt_tokens = 0
    return r


def setwise_ranker(tok, rec, cfg, num_child, k, scoring, method):
    r = SetwiseLlmRanker.__new__(SetwiseLlmRanker)
    r.tokenizer, r.llm, r.config = tok, rec, cfg
    r.device, r.num_child, r.k, r.scoring, r.method, r.num_permutation = "cpu", num_child, k, scoring, method, 1
    r.decoder_input_ids = tok.encode("<pad> Passage", return_tensors="pt", add_special_tokens=False)
    r.target_token_ids = tok([f"<pad> Passage {c}" for c in SetwiseLlmRanker.CHARACTERS], return_tensors="pt",
                             add_special_tokens=False, padding=True).input_ids[:, -1]
    r.total_compare = r.total_completion_tokens = r.total_prompt_tokens = 0
    return r


def pairwise_ranker(tok, rec, cfg, method, batch_size, k):
    r = PairwiseLlmRanker.__new__(PairwiseLlmRanker)
    r.tokenizer, r.llm, r.config = tok, rec, cfg
    r.device, r.method, r.batch_size, r.k = "cpu", method, batch_size, k
    r.prompt = """Given a query "{query}", which of the following two passages is more relevant to the query?

Passage A: "{doc1}"

Passage B: "{doc2}"

Output Passage A or Passage B:"""  # llmrankers/pairwise.py:42-48 (constructor bypassed, so restated here)
    r.decoder_input_ids = tok.encode("<pad> Passage", return_tensors="pt", add_special_tokens=False).repeat(batch_size, 1)
    r.total_compare = r.total_completion_tokens = r.total_prompt_tokens = 0
    return r


def counters(r):
    return dict(total_compare=int(r.total_compare), total_prompt_tokens=int(r.total_prompt_tokens),
                total_completion_tokens=int(r.total_completion_tokens))


def save_calls(store, prefix, calls, cols=None):
    for i, c in enumerate(calls):
        p = f"{prefix}/call{i}"
        if c["kind"] == "forward":
            for k, v in c["inputs"].items():
                store[f"{p}/{k}"] = v.numpy()
            lg = c["logits"].numpy()
            store[f"{p}/logits"] = lg if cols is None else lg[..., cols]
        else:
            store[f"{p}/input_ids"] = c["input_ids"].numpy()
            store[f"{p}/output"] = c["output"].numpy()
    return len(calls)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    tok = synthetic_tokenizer()
    meta = {"transformers": __import__("transformers").__version__, "torch": torch.__version__, "cases": {}}

    # ---- A. relative position buckets straight from HF
    rel = torch.arange(-300, 301)
    np.savez_compressed(os.path.join(HERE, "buckets.npz"), rel=rel.numpy(),
                        bidirectional=T5Attention._relative_position_bucket(rel, True, 32, 128).numpy(),
                        unidirectional=T5Attention._relative_position_bucket(rel, False, 32, 128).numpy())

    # ---- B. tiny model, every reference mode
    store = {}
    cfg = model_cfg("t5-tiny", TINY_VOCAB)
    seed = 1234
    weights = synthetic_weights(cfg, seed)
    model, hf_cfg = build_hf_model(cfg, weights)
    rng = np.random.default_rng(7)
    query = "w11 w23 w5 w42 w8"
    docs = make_docs(rng, 10, 6, 20)
    yes_id = tok.encode("Yes", add_special_tokens=False)[0]
    no_id = tok.encode("No", add_special_tokens=False)[0]
    meta["tiny"] = dict(model="t5-tiny", vocab_size=TINY_VOCAB, seed=seed, query=query, yes_id=yes_id, no_id=no_id,
                        docs=[dict(docid=d.docid, score=d.score, text=d.text) for d in docs])

    rec = Recorder(model)
    r = pointwise_ranker(tok, rec, hf_cfg, "yes_no", 4)
    out = r.rerank(query, copy.deepcopy(docs))
    n = save_calls(store, "yes_no", rec.calls)
    meta["cases"]["yes_no"] = dict(n_calls=n, order=[d.docid for d in out], scores={d.docid: d.score for d in out}, **counters(r))

    rec = Recorder(model)
    r = pointwise_ranker(tok, rec, hf_cfg, "qlm", 4)
    out = r.rerank(query, copy.deepcopy(docs))
    n = save_calls(store, "qlm", rec.calls)
    meta["cases"]["qlm"] = dict(n_calls=n, order=[d.docid for d in out], scores={d.docid: d.score for d in out},
                                labels=tok.encode(f"<pad> {query}", add_special_tokens=False), **counters(r))

    # label-favouring lm_head so that generation emits passage labels: x30 on the rows of ▁A..▁D
    # (only the 4 labels a num_child=3 window can hold: the reference's bubblesort raises IndexError on a label
    # beyond the window, setwise.py:259 — heapify catches it (:210-213), bubblesort does not)
    label_ids = [tok.convert_tokens_to_ids("▁" + c) for c in LABELS[:4]]
    w_lab = dict(weights)
    w_lab["lm_head.weight"] = weights["lm_head.weight"].copy()
    w_lab["lm_head.weight"][label_ids] *= 30.0
    model_lab, _ = build_hf_model(cfg, w_lab)
    meta["tiny"]["label_ids"] = label_ids
    meta["tiny"]["label_boost"] = 30.0
    docs12 = make_docs(np.random.default_rng(8), 12, 5, 12)
    meta["tiny"]["docs12"] = [dict(docid=d.docid, score=d.score, text=d.text) for d in docs12]

    for name, scoring, method, mdl in (("setwise_heap_gen", "generation", "heapsort", model_lab),
                                       ("setwise_heap_lik", "likelihood", "heapsort", model),
                                       ("setwise_bubble_lik", "likelihood", "bubblesort", model),
                                       ("setwise_bubble_gen", "generation", "bubblesort", model_lab)):
        rec = Recorder(mdl)
        r = setwise_ranker(tok, rec, hf_cfg, 3, 3, scoring, method)
        out = r.rerank(query, copy.deepcopy(docs12))
        n = save_calls(store, name, rec.calls, cols=None)
        meta["cases"][name] = dict(n_calls=n, order=[d.docid for d in out], scores=[d.score for d in out], num_child=3, k=3,
                                   scoring=scoring, method=method, label_favouring=mdl is model_lab, **counters(r))
    meta["tiny"]["target_token_ids"] = r.target_token_ids.tolist()
    meta["tiny"]["decoder_prefix"] = r.decoder_input_ids[0].tolist()

    docs6 = docs12[:6]
    for name, method, bs in (("pairwise_allpair", "allpair", 4), ("pairwise_heap", "heapsort", 2), ("pairwise_bubble", "bubblesort", 2)):
        rec = Recorder(model_lab)
        r = pairwise_ranker(tok, rec, hf_cfg, method, bs, 3)
        out = r.rerank(query, copy.deepcopy(docs6))
        n = save_calls(store, name, rec.calls)
        meta["cases"][name] = dict(n_calls=n, order=[d.docid for d in out], scores=[d.score for d in out], method=method,
                                   batch_size=bs, k=3, **counters(r))

    # truncate() (pointwise.py:132-133)
    meta["tiny"]["truncate"] = [dict(text=docs[0].text, length=L, out=r.truncate(docs[0].text, L)) for L in (1, 3, 5, 50)]
    np.savez_compressed(os.path.join(HERE, "golden_tiny.npz"), **store)

    # ---- C. BASELINE config 1: flan-t5-small shape, yes_no, 1 query x 10 passages, batch_size 4, CPU
    store = {}
    cfg = model_cfg("flan-t5-small", 32128)
    seed = 929
    weights = synthetic_weights(cfg, seed)
    model, hf_cfg = build_hf_model(cfg, weights)
    rng = np.random.default_rng(929)
    query = " ".join(f"w{int(x)}" for x in rng.integers(0, 2000, 32))
    docs = make_docs(rng, 10, 64, 128)
    rec = Recorder(model)
    r = pointwise_ranker(tok, rec, hf_cfg, "yes_no", 4)
    out = r.rerank(query, copy.deepcopy(docs))
    n = save_calls(store, "yes_no", rec.calls, cols=[yes_id, no_id])
    meta["small"] = dict(model="flan-t5-small", vocab_size=32128, seed=seed, query=query, yes_id=yes_id, no_id=no_id,
                         docs=[dict(docid=d.docid, score=d.score, text=d.text) for d in docs])
    meta["cases"]["small_yes_no"] = dict(n_calls=n, order=[d.docid for d in out], scores={d.docid: d.score for d in out}, **counters(r))
    rec = Recorder(model)
    r = pointwise_ranker(tok, rec, hf_cfg, "qlm", 4)
    out = r.rerank(query, copy.deepcopy(docs))
    for i, c in enumerate(rec.calls):
        for k, v in c["inputs"].items():
            store[f"qlm/call{i}/{k}"] = v.numpy()
    meta["cases"]["small_qlm"] = dict(n_calls=len(rec.calls), order=[d.docid for d in out], scores={d.docid: d.score for d in out},
                                      labels=tok.encode(f"<pad> {query}", add_special_tokens=False), **counters(r))
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **store)

    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    for fn in ("buckets.npz", "golden_tiny.npz", "golden_small.npz", "golden_meta.json"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")


if __name__ == "__main__":
    main()
