"""CPU tests of the listwise widening: llmrankers.listwise.ListwiseLlmRanker with the oracle standing in for the
engine must reproduce what the reference's own ListwiseLlmRanker.rerank() produced (tests/golden/make_golden_listwise.py): every
window's response string, the final order and scores, the counters; the response parser and the completion prompt are pinned to the
reference functions by their own fixtures."""
import json
import os

import numpy as np
import pytest

from fake_backend import OracleBackend
from helpers import GOLDEN, golden_npz, calls
from test_host_logic import docs_from, tokenizer

_cache = {}


def meta():
    if "m" not in _cache:
        with open(os.path.join(GOLDEN, "golden_listwise_meta.json")) as f:
            _cache["m"] = json.load(f)
    return _cache["m"]


def backend(digit_favouring=False):
    from b200rank.synthetic import model_cfg, synthetic_weights
    from oracle.t5_oracle import T5Oracle
    key = ("be", digit_favouring)
    if key not in _cache:
        m = meta()
        cfg = model_cfg(m["model"], m["vocab_size"])
        w = synthetic_weights(cfg, m["seed"])
        if digit_favouring:
            w = dict(w)
            w["lm_head.weight"] = w["lm_head.weight"].copy()
            w["lm_head.weight"][m["digit_ids"]] *= m["digit_boost"]
        _cache[key] = OracleBackend(T5Oracle(cfg, w), tokenizer(), cfg)
    return _cache[key]


def ranker(c):
    from llmrankers.listwise import ListwiseLlmRanker
    return ListwiseLlmRanker(None, None, "cuda", c["window_size"], c["step_size"], scoring=c["scoring"], num_repeat=c["num_repeat"],
                             backend=backend(c["digit_favouring"]))


@pytest.mark.parametrize("case", ["listwise_lik", "listwise_lik_rep2", "listwise_gen", "listwise_gen_plain"])
def test_listwise_matches_reference(case):
    m = meta()
    c = m["cases"][case]
    r = ranker(c)
    assert r.decoder_input_ids == m["decoder_prefix"] and r.target_token_ids == m["target_token_ids"]
    seen = []
    orig = r.compare

    def spy(q, ds):
        out = orig(q, ds)
        seen.append(dict(docids=[d.docid for d in ds], output=out))
        return out
    r.compare = spy
    docs = docs_from(m["docs12"])
    out = r.rerank(m["query"], docs)
    assert seen == c["compares"]
    assert [d.docid for d in out] == c["order"] and [d.score for d in out] == c["scores"]
    assert (r.total_compare, r.total_prompt_tokens, r.total_completion_tokens) == (c["total_compare"], c["total_prompt_tokens"], c["total_completion_tokens"])
    assert all(a is not b for a in out for b in docs)   # the reference returns deep copies (listwise.py:182)


@pytest.mark.parametrize("case", ["listwise_lik", "listwise_gen"])
def test_listwise_prompt_tokens_and_generated_ids_match_the_recorded_calls(case):
    """What crossed the model boundary in the reference run: the prompt token ids of every compare, and for generation the ids
    `generate()` returned (start token + 20 new tokens on transformers 5.5)."""
    m = meta()
    c = m["cases"][case]
    r = ranker(c)
    rec = calls(golden_npz("golden_listwise.npz"), case)
    assert len(rec) == c["n_calls"]
    for call, cmp_ in zip(rec, c["compares"]):
        by_id = {d["docid"]: d for d in m["docs12"]}
        docs = docs_from([by_id[i] for i in cmp_["docids"]])
        if c["scoring"] == "likelihood":
            row = r._likelihood_rows(m["query"], [docs])[0]
            assert row == call["input_ids"][0].tolist()
            assert call["decoder_input_ids"][0].tolist() == r.decoder_input_ids
            probs = r.backend.label_probs([row], r.decoder_input_ids, r.target_token_ids[:len(docs)])[0]
            lg = call["logits"][0, -1].astype(np.float64)
            ref = np.exp(lg - lg.max()) / np.exp(lg - lg.max()).sum()
            np.testing.assert_allclose(probs, ref[r.target_token_ids[:len(docs)]], rtol=2e-3, atol=1e-7)
        else:
            row = r._generation_row(m["query"], docs)
            assert row == call["input_ids"][0].tolist()
            assert r._generate_free(row) == call["output"][0].tolist()


def test_generation_budget_and_eos_stop(monkeypatch):
    """Chunked free-form decoding: the budget is honoured within one call of the greedy entry point and across chunks (8-token
    chunks here, so that the default budget spans three calls), and an </s> ends it."""
    m = meta()
    c = m["cases"]["listwise_gen"]
    r = ranker(c)
    one_call = r._generate_free(r._generation_row(m["query"], docs_from(m["docs12"][:3])))
    r.GREEDY_CHUNK = 8
    row = r._generation_row(m["query"], docs_from(m["docs12"][:3]))
    full = r._generate_free(row)
    assert len(full) == 21 and full[0] == 0 and full == one_call
    for budget in (1, 7, 8, 9, 16, 19):
        monkeypatch.setenv("B200RANK_LISTWISE_MAX_NEW", str(budget))
        assert r._generate_free(row) == full[:budget + 1]

    class EosAfter(OracleBackend):
        def generate_rows(self, rows, dec_prefix, max_new):
            outs = super().generate_rows(rows, dec_prefix, max_new)
            return [np.concatenate([o[:len(dec_prefix) + 2], [self.eos_id]]) if len(dec_prefix) >= 9 else o for o in outs]
    monkeypatch.setenv("B200RANK_LISTWISE_MAX_NEW", "20")
    b = backend(True)
    r.backend = EosAfter(b.oracle, b.tokenizer, b.cfg)
    got = r._generate_free(row)
    assert got[:9] == full[:9] and got[-1] == 1 and len(got) == 12


def test_response_parser_matches_reference():
    from llmrankers.listwise import clean_response, receive_permutation, remove_duplicate
    from llmrankers.rankers import SearchResult
    for t in meta()["receive_permutation"]:
        ranking = [SearchResult(docid=f"p{i}", score=0.0, text="") for i in range(t["n"])]
        out = receive_permutation(ranking, t["response"], t["rank_start"], t["rank_end"])
        assert out is ranking and [d.docid for d in out] == t["order"], t
    assert clean_response("[12] > [3]x") == "12     3"
    assert remove_duplicate([3, 1, 3, 2, 1]) == [3, 1, 2]


def test_completion_prompt_matches_reference():
    from llmrankers.listwise import create_permutation_instruction_complete
    from llmrankers.rankers import SearchResult
    t = meta()["instruction_complete"]
    docs = [SearchResult(docid=str(i), score=0.0, text=x) for i, x in enumerate(t["docs"])]
    assert create_permutation_instruction_complete(t["query"], docs) == t["text"]


def test_windows_and_degenerate_inputs():
    from llmrankers.listwise import ListwiseLlmRanker, _window_positions
    assert list(_window_positions(12, 4, 2)) == [(8, 12), (6, 10), (4, 8), (2, 6), (0, 4)]
    assert list(_window_positions(12, 5, 3)) == [(7, 12), (4, 9), (1, 6)]
    assert list(_window_positions(3, 4, 2)) == []                  # fewer documents than the window: no compare at all (listwise.py:184-185)
    assert list(_window_positions(4, 4, 10)) == [(0, 4)]
    m = meta()
    r = ListwiseLlmRanker(None, None, "cuda", 20, 10, scoring="likelihood", backend=backend())
    docs = docs_from(m["docs12"])
    out = r.rerank(m["query"], docs)
    assert [d.docid for d in out] == [d.docid for d in docs] and [d.score for d in out] == [-i for i in range(12)] and r.total_compare == 0
    assert r.rerank(m["query"], []) == []
    r24 = ListwiseLlmRanker(None, None, "cuda", 24, 10, scoring="likelihood", backend=backend())
    with pytest.raises(IndexError):                                # 24 passages, 23 labels: the reference fails the same way
        r24.rerank("w1", docs_from(m["docs12"] + [dict(d, docid="e" + d["docid"]) for d in m["docs12"]]))


@pytest.mark.parametrize("case", ["listwise_lik", "listwise_lik_rep2", "listwise_gen"])
def test_listwise_rerank_many_equals_rerank(case):
    m = meta()
    c = m["cases"][case]
    docs12 = m["docs12"]
    requests = [(m["query"], docs12), ("w7 w8", docs12[:5]), ("w1 w2 w3", docs12[3:]), (m["query"], docs12[:3]), ("w9", docs12[::-1]), ("w4", [])]
    want = []
    for q, dd in requests:
        r = ranker(c)
        out = r.rerank(q, docs_from(dd))
        want.append(([d.docid for d in out], [d.score for d in out], r.total_compare, r.total_prompt_tokens, r.total_completion_tokens))
    assert want[0][0] == c["order"]
    for window in (1, 2, 8):
        r = ranker(c)
        got = []
        for out in r.rerank_many([(q, docs_from(dd)) for q, dd in requests], window=window):
            got.append(([d.docid for d in out], [d.score for d in out], r.total_compare, r.total_prompt_tokens, r.total_completion_tokens))
        assert got == want, window


@pytest.mark.parametrize("case", ["listwise_lik", "listwise_lik_rep2", "listwise_gen"])
def test_cli_listwise_end_to_end(case, tmp_path, monkeypatch, capsys):
    """run.py with the listwise sub-command: run file and the printed counters against the reference's own rerank()."""
    import run as cli_mod
    from llmrankers import _backend
    m = meta()
    c = m["cases"][case]
    monkeypatch.setattr(_backend.T5Backend, "load", classmethod(lambda cls, *a, **k: backend(c["digit_favouring"])))
    docs = m["docs12"]
    (tmp_path / "queries.tsv").write_text(f"q1\t{m['query']}\n")
    (tmp_path / "docs.tsv").write_text("".join(f"{d['docid']}\t{d['text']}\n" for d in docs))
    (tmp_path / "run.txt").write_text("".join(f"q1 Q0 {d['docid']} {i + 1} {d['score']} bm25\n" for i, d in enumerate(docs)))
    out = tmp_path / "out.txt"
    cli_mod.cli(["run", "--model_name_or_path", "synthetic:t5-tiny", "--run_path", str(tmp_path / "run.txt"), "--save_path", str(out),
                 "--queries_tsv", str(tmp_path / "queries.tsv"), "--collection_tsv", str(tmp_path / "docs.tsv"),
                 "--query_length", "32", "--passage_length", "128", "--scoring", c["scoring"],
                 "listwise", "--window_size", str(c["window_size"]), "--step_size", str(c["step_size"]), "--num_repeat", str(c["num_repeat"])])
    lines = out.read_text().splitlines()
    assert [l.split("\t")[2] for l in lines] == c["order"]
    assert [float(l.split("\t")[4]) for l in lines] == [float(x) for x in c["scores"]]
    printed = capsys.readouterr().out
    assert f"Avg comparisons: {float(c['total_compare'])}" in printed
    assert f"Avg prompt tokens: {float(c['total_prompt_tokens'])}" in printed
    assert f"Avg completion tokens: {float(c['total_completion_tokens'])}" in printed


# ------------------------------------------------------------------------------------------- window arithmetic against the reference
def test_sliding_windows_match_the_reference_over_a_grid():
    """tests/golden/golden_listwise_windows.json: the reference's own rerank() loop and response handling (listwise.py:110-144, 177-195)
    under a deterministic stand-in for compare(), 272 configurations of list size / window / step / num_repeat with partial, duplicated,
    out-of-window, prose, junk and 'ERROR::reduce_length' responses (tests/golden/make_golden_listwise_windows.py). The drop-in class
    must ask for the same windows in the same order and return the same ranking, scores and compare count — and, like the reference,
    hand back the caller's own objects when num_repeat is 0."""
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden_listwise_windows import make_docs, response_for
    from llmrankers.listwise import ListwiseLlmRanker
    from llmrankers.rankers import SearchResult
    with open(os.path.join(GOLDEN, "golden_listwise_windows.json")) as f:
        cases = json.load(f)
    assert len(cases) == 272 and sum(len(c["calls"]) for c in cases) == 824
    for c in cases:
        r = ListwiseLlmRanker(None, None, "cuda", window_size=c["window_size"], step_size=c["step_size"], scoring="generation",
                              num_repeat=c["num_repeat"], backend=backend())
        log = []

        def compare(query, docs, c=c, log=log, r=r):
            r.total_compare += 1
            ids = [d.docid for d in docs]
            log.append(ids)
            return response_for(c["seed"], c["p_bad"], query, ids)
        r.compare = compare
        docs = make_docs(c["n"], SearchResult)
        res = r.rerank(c["query"], docs)
        assert log == c["calls"], c
        assert [[d.docid, d.score] for d in res] == c["result"], c
        assert (res is docs) == c["returns_input_objects"], c
        assert [d.score for d in docs] == c["input_scores_after"], c
        assert r.total_compare == c["total_compare"]
        got_many = list(r.rerank_many([(c["query"], make_docs(c["n"], SearchResult))]))     # generation mode: loops over rerank()
        assert [[d.docid, d.score] for d in got_many[0]] == c["result"], c
