"""CPU tests of the host side: the drop-in `llmrankers` classes (with the oracle standing in for the GPU engine) must
reproduce what the reference's own rerank() produced in tests/golden (order, scores, counters, truncate), the C-ABI
library must load and export every symbol of include/b200rank.h, and compute entry points must fail loudly without a GPU."""
import copy
import ctypes
import os
import re

import numpy as np
import pytest

from fake_backend import OracleBackend
from helpers import ROOT, golden_meta, model_and_weights, oracle_for

_tok = None


def tokenizer():
    global _tok
    if _tok is None:
        from b200rank.synthetic import synthetic_tokenizer
        _tok = synthetic_tokenizer()
    return _tok


def backend(which="tiny", lab=False):
    cfg, _ = model_and_weights(which, lab)
    return OracleBackend(oracle_for(which, lab), tokenizer(), cfg)


def docs_from(meta_docs):
    from llmrankers.rankers import SearchResult
    return [SearchResult(docid=d["docid"], score=d["score"], text=d["text"]) for d in meta_docs]


def check_counters(r, c):
    assert (r.total_compare, r.total_prompt_tokens, r.total_completion_tokens) == \
           (c["total_compare"], c["total_prompt_tokens"], c["total_completion_tokens"])


@pytest.mark.parametrize("which,method,case", [("tiny", "yes_no", "yes_no"), ("tiny", "qlm", "qlm"), ("small", "yes_no", "small_yes_no")])
def test_pointwise_matches_reference(which, method, case):
    from llmrankers.pointwise import PointwiseLlmRanker
    meta = golden_meta()
    m, c = meta[which], meta["cases"][case]
    r = PointwiseLlmRanker(None, None, "cuda", method=method, batch_size=4, backend=backend(which))
    docs = docs_from(m["docs"])
    out = r.rerank(m["query"], docs)
    assert [d.docid for d in out] == c["order"]
    for d in out:
        assert d.score == pytest.approx(c["scores"][d.docid], rel=2e-5, abs=2e-3 if method == "qlm" else 1e-5)
        assert d.text is not None  # pointwise returns the input objects, text intact
    check_counters(r, c)


def v10_backend(which="tiny"):
    from helpers import v10_model_and_weights, v10_oracle_for
    return OracleBackend(v10_oracle_for(which), tokenizer(), v10_model_and_weights(which)[0])


@pytest.mark.parametrize("which,case", [("tiny", "mono"), ("small", "small_mono")])
def test_monot5_matches_reference(which, case):
    """MonoT5LlmRanker.rerank (pointwise.py:136-186) on a T5 v1.0 model: order, scores, counters."""
    from helpers import golden_v10_meta
    from llmrankers.pointwise import MonoT5LlmRanker
    meta = golden_v10_meta()
    m, c = meta[which], meta["cases"][case]
    r = MonoT5LlmRanker(None, None, "cuda", batch_size=4, backend=v10_backend(which))
    out = r.rerank(m["query"], docs_from(m["docs"]))
    assert [d.docid for d in out] == c["order"]
    for d in out:
        assert d.score == pytest.approx(c["scores"][d.docid], abs=1e-5) and d.text is not None
    check_counters(r, c)


def test_monot5_rerank_many_equals_rerank():
    from helpers import golden_v10_meta
    from llmrankers.pointwise import MonoT5LlmRanker
    meta = golden_v10_meta()
    m, c = meta["tiny"], meta["cases"]["mono"]
    r = MonoT5LlmRanker(None, None, "cuda", batch_size=4, backend=v10_backend("tiny"))
    reqs = [(m["query"], docs_from(m["docs"])), ("w1 w2", docs_from(m["docs"][:3])), (m["query"], [])]
    outs = list(r.rerank_many(iter(reqs)))
    assert [d.docid for d in outs[0]] == c["order"] and len(outs[1]) == 3 and outs[2] == []
    for d in outs[0]:
        assert d.score == pytest.approx(c["scores"][d.docid], abs=1e-5)


def test_duot5_matches_reference():
    """DuoT5LlmRanker.rerank heapsort (pairwise.py:296-352): compare sequence, order, scores, counters."""
    from helpers import golden_v10_meta
    from llmrankers.pairwise import DuoT5LlmRanker
    meta = golden_v10_meta()
    m, c = meta["tiny"], meta["cases"]["duo_heap"]
    r = DuoT5LlmRanker(None, None, "cuda", method="heapsort", k=c["k"], backend=v10_backend("tiny"))
    out = r.rerank(m["query"], docs_from(m["docs"][:7]))
    assert [d.docid for d in out] == c["order"] and [d.score for d in out] == c["scores"]
    assert all(d.text is None for d in out)
    check_counters(r, c)
    with pytest.raises(NotImplementedError):
        DuoT5LlmRanker(None, None, "cuda", method="allpair", backend=v10_backend("tiny")).rerank(m["query"], docs_from(m["docs"][:3]))


def test_duot5_batched_sequential_and_cross_query_drivers_agree(monkeypatch):
    """DuoT5's level-parallel heap construction (one engine call per round of compares) and rerank_many (several queries in lockstep)
    against the sequential compare-at-a-time driver the reference fixture pins: same order, scores and counters."""
    from helpers import golden_v10_meta
    from llmrankers.pairwise import DuoT5LlmRanker
    meta = golden_v10_meta()
    m, c = meta["tiny"], meta["cases"]["duo_heap"]
    assert DuoT5LlmRanker(None, None, "cuda", method="heapsort", k=2, backend=v10_backend("tiny"))._has_batched_compares()
    requests = [(m["query"], m["docs"][:7]), ("w3 w4", m["docs"][2:9]), (m["query"], m["docs"][:1]), ("w1 w2 w3", m["docs"][::-1]), ("w5", [])]
    monkeypatch.setenv("B200RANK_BATCHED_SORT", "0")
    seq = DuoT5LlmRanker(None, None, "cuda", method="heapsort", k=c["k"], backend=v10_backend("tiny"))
    want = []
    for q, dd in requests:
        out = seq.rerank(q, docs_from(dd))
        want.append(([(d.docid, d.score) for d in out], (seq.total_compare, seq.total_prompt_tokens, seq.total_completion_tokens)))
    assert want[0][0] == list(zip(c["order"], c["scores"]))
    monkeypatch.setenv("B200RANK_BATCHED_SORT", "1")
    bat = DuoT5LlmRanker(None, None, "cuda", method="heapsort", k=c["k"], backend=v10_backend("tiny"))
    calls = []
    orig = bat.backend.score_yes_no
    bat.backend.score_yes_no = lambda rows, y, n: (calls.append(len(rows)), orig(rows, y, n))[1]
    for (q, dd), w in zip(requests, want):
        out = bat.rerank(q, docs_from(dd))
        assert ([(d.docid, d.score) for d in out], (bat.total_compare, bat.total_prompt_tokens, bat.total_completion_tokens)) == w
    assert max(calls) > 2            # several pairs really shared an engine call
    for window in (1, 3):
        got = []
        for out in bat.rerank_many([(q, docs_from(dd)) for q, dd in requests], window=window):
            got.append(([(d.docid, d.score) for d in out], (bat.total_compare, bat.total_prompt_tokens, bat.total_completion_tokens)))
        assert got == want


@pytest.mark.parametrize("case", ["setwise_heap_gen", "setwise_heap_lik", "setwise_bubble_lik", "setwise_bubble_gen"])
def test_setwise_matches_reference(case, capsys):
    from llmrankers.setwise import SetwiseLlmRanker
    meta = golden_meta()
    m, c = meta["tiny"], meta["cases"][case]
    r = SetwiseLlmRanker(None, None, "cuda", num_child=c["num_child"], k=c["k"], scoring=c["scoring"], method=c["method"],
                         backend=backend("tiny", c["label_favouring"]))
    assert r.decoder_input_ids == m["decoder_prefix"] and r.target_token_ids == m["target_token_ids"]
    out = r.rerank(m["query"], docs_from(m["docs12"]))
    assert [d.docid for d in out] == c["order"]
    assert [d.score for d in out] == c["scores"]
    assert all(d.text is None for d in out)
    check_counters(r, c)


@pytest.mark.parametrize("case", ["pairwise_allpair", "pairwise_heap", "pairwise_bubble"])
def test_pairwise_matches_reference(case):
    from llmrankers.pairwise import PairwiseLlmRanker
    meta = golden_meta()
    m, c = meta["tiny"], meta["cases"][case]
    r = PairwiseLlmRanker(None, None, "cuda", method=c["method"], batch_size=c["batch_size"], k=c["k"], backend=backend("tiny", True))
    out = r.rerank(m["query"], docs_from(m["docs12"][:6]))
    assert [d.docid for d in out] == c["order"]
    assert [d.score for d in out] == c["scores"]
    check_counters(r, c)


def test_truncate_matches_reference():
    from llmrankers.pointwise import PointwiseLlmRanker
    r = PointwiseLlmRanker(None, None, "cuda", backend=backend())
    for t in golden_meta()["tiny"]["truncate"]:
        assert r.truncate(t["text"], t["length"]) == t["out"]


def test_pointwise_sort_is_stable_and_empty_ok():
    from llmrankers.pointwise import PointwiseLlmRanker
    from llmrankers.rankers import SearchResult
    r = PointwiseLlmRanker(None, None, "cuda", method="yes_no", batch_size=3, backend=backend())
    assert r.rerank("w1", []) == []
    same = [SearchResult(f"d{i}", 0.0, "w5 w6 w7") for i in range(5)]  # identical text => identical scores => input order kept
    assert [d.docid for d in r.rerank("w1 w2", same)] == ["d0", "d1", "d2", "d3", "d4"]
    assert r.total_compare == 2


@pytest.mark.parametrize("method", ["yes_no", "qlm"])
def test_rerank_many_equals_rerank(method):
    from llmrankers.pointwise import PointwiseLlmRanker
    m = golden_meta()["tiny"]
    r = PointwiseLlmRanker(None, None, "cuda", method=method, batch_size=4, backend=backend())
    reqs = [(m["query"], docs_from(m["docs"])), ("w3 w4", docs_from(m["docs"][:3])), ("w9", []), (m["query"], docs_from(m["docs"][::-1]))]
    want = [[(d.docid, d.score) for d in r.rerank(q, copy.deepcopy(rk))] for q, rk in reqs]
    got = [[(d.docid, d.score) for d in out] for out in r.rerank_many([(q, copy.deepcopy(rk)) for q, rk in reqs])]
    assert got == want
    assert r.total_compare == 3  # counters describe the last query (10 docs, batch_size 4)


def test_rerank_many_merges_queries_into_passes():
    """queries_per_pass: consecutive queries share one submit (backends that state batch-composition invariance only), every query gets
    its own scores / order / counters back, and a declined merged pass falls back to one query per pass for good."""
    from llmrankers.pointwise import PointwiseLlmRanker
    m = golden_meta()["tiny"]

    class Merging(type(backend())):
        batch_invariant = True
        calls = []
        max_rows = 10 ** 6

        def submit_yes_no(self, rows, yes_id, no_id):
            if len(rows) > self.max_rows:
                self.calls.append(("declined", len(rows)))
                return None
            self.calls.append(("submit", len(rows)))
            return super().submit_yes_no(rows, yes_id, no_id)

    b = backend()
    b.__class__ = Merging
    r = PointwiseLlmRanker(None, None, "cuda", method="yes_no", batch_size=4, backend=b)
    reqs = [(m["query"], docs_from(m["docs"])), ("w3 w4", docs_from(m["docs"][:3])), ("w9", []), (m["query"], docs_from(m["docs"][::-1])),
            ("w1", docs_from(m["docs"][:2]))]
    want = [[(d.docid, d.score) for d in r.rerank(q, copy.deepcopy(rk))] for q, rk in reqs]

    def check(got):
        assert [[d for d, _ in q] for q in got] == [[d for d, _ in q] for q in want]
        for g, w in zip(got, want):
            np.testing.assert_allclose([s for _, s in g], [s for _, s in w], rtol=0, atol=1e-5)   # the numpy stand-in is not bit-invariant
    for qpp, expect in ((2, [("submit", 13), ("submit", 10), ("submit", 2)]), (3, [("submit", 13), ("submit", 12)]), (1, None)):
        Merging.calls = []
        check([[(d.docid, d.score) for d in out] for out in r.rerank_many([(q, copy.deepcopy(rk)) for q, rk in reqs], queries_per_pass=qpp)])
        if expect is not None:
            assert Merging.calls == expect
    assert r.total_compare == 1   # counters describe the last query (2 docs, batch_size 4)
    # a merged pass that does not fit: the first query goes alone, the others are handed back in order, no further merging
    Merging.calls, Merging.max_rows = [], 12
    check([[(d.docid, d.score) for d in out] for out in r.rerank_many([(q, copy.deepcopy(rk)) for q, rk in reqs], queries_per_pass=2)])
    assert Merging.calls == [("declined", 13), ("submit", 10), ("submit", 3), ("submit", 10), ("submit", 2)]


def test_rerank_many_falls_back_for_batches_the_pipeline_declines():
    """b200rank_submit_yes_no only takes batches that fit the pipelined pass (documents of <= 240 tokens, one device pass); for the
    others T5Backend.submit_yes_no returns None and rerank_many must drain the query in flight (the engine refuses synchronous calls
    next to pipelined batches) and score the query synchronously — same results, same order of results as rerank()."""
    from llmrankers.pointwise import PointwiseLlmRanker
    m = golden_meta()["tiny"]

    class Picky(type(backend())):
        in_flight = 0

        def submit_yes_no(self, rows, yes_id, no_id):
            if len(rows) > 4:            # stands in for "does not fit the pipelined pass"
                return None
            self.in_flight += 1
            return super().submit_yes_no(rows, yes_id, no_id)

        def wait_yes_no(self, ticket):
            self.in_flight -= 1
            return super().wait_yes_no(ticket)

        def score_yes_no(self, rows, yes_id, no_id):
            if getattr(self, "_strict", False):
                assert self.in_flight == 0, "synchronous call while a pipelined batch is in flight"
            return super().score_yes_no(rows, yes_id, no_id)

    b = backend()
    b.__class__ = Picky
    r = PointwiseLlmRanker(None, None, "cuda", method="yes_no", batch_size=4, backend=b)
    reqs = [("w3 w4", docs_from(m["docs"][:3])), (m["query"], docs_from(m["docs"])), ("w9", []), ("w1", docs_from(m["docs"][:2])),
            (m["query"], docs_from(m["docs"][::-1])), ("w3", docs_from(m["docs"][:4]))]
    want = [[(d.docid, d.score) for d in r.rerank(q, copy.deepcopy(rk))] for q, rk in reqs]
    b._strict = True
    # OracleBackend.submit_yes_no computes through score_yes_no itself: count only the calls rerank_many makes directly
    orig_submit = Picky.submit_yes_no

    def submit(self, rows, yes_id, no_id):
        self._strict = False
        try:
            return orig_submit(self, rows, yes_id, no_id)
        finally:
            self._strict = True
    Picky.submit_yes_no = submit
    got = [[(d.docid, d.score) for d in out] for out in r.rerank_many([(q, copy.deepcopy(rk)) for q, rk in reqs])]
    assert got == want
    assert b.in_flight == 0


def test_rerank_many_leaves_no_ticket_in_flight_when_abandoned():
    """A consumer that stops after the first result (or an exception between submit and wait) must not leave a pipelined batch in
    flight: the engine's synchronous entry points refuse to run until every ticket has been waited for."""
    from llmrankers.pointwise import PointwiseLlmRanker
    m = golden_meta()["tiny"]
    b = backend()
    flight = set()
    sub, wait = b.submit_yes_no, b.wait_yes_no

    def submit(rows, y, n):
        t = sub(rows, y, n)
        flight.add(id(t))
        return t

    def waited(t):
        flight.discard(id(t))
        return wait(t)
    b.submit_yes_no, b.wait_yes_no = submit, waited
    r = PointwiseLlmRanker(None, None, "cuda", method="yes_no", batch_size=4, backend=b)
    reqs = [(m["query"], docs_from(m["docs"])), ("w3 w4", docs_from(m["docs"][:3])), ("w5", docs_from(m["docs"][:5]))]
    gen = r.rerank_many(reqs)
    first = next(gen)
    assert len(first) == len(m["docs"]) and flight            # the second query is in flight while the first is handed out
    gen.close()
    assert not flight
    assert [d.docid for d in r.rerank(m["query"], docs_from(m["docs"]))] == [d.docid for d in first]


# ------------------------------------------------------------------------------------------- permutation voting (setwise.py:102-157)
def _perm_fixture():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden_setwise_perm.json")) as f:
        return json.load(f)


def _perm_backend(g, log):
    """The oracle-backed test backend with generate() replaced by the stand-in the fixture generator put behind the REFERENCE's
    `self.llm.generate` (tests/golden/make_golden_setwise_perm.py::stub_generate)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_setwise_perm import stub_generate
    b = backend()

    def generate(padded_ids, dec_prefix, max_new):
        rows = np.asarray(padded_ids).tolist()
        assert max_new == 2 and list(dec_prefix) == [0, g["passage_id"]]
        out = stub_generate(rows, g["passage_id"], g["label_token_ids"])
        log.append(dict(input_ids=rows, output=out))
        return np.asarray(out, np.int64)
    b.generate = generate
    return b


def test_setwise_permutation_voting_matches_reference(capsys):
    """num_permutation > 1: the shuffles drawn from the module RNG (seed 929), the prompt of every shuffle token for token, the vote
    with its rejections ('X X', labels that are not in the prompt), the tie-break and the counters, compare by compare."""
    import random
    from llmrankers.setwise import SetwiseLlmRanker
    g = _perm_fixture()
    all_docs = docs_from(g["docs"])
    by_perm = {}
    for c in g["compares"]:
        by_perm.setdefault(c["num_permutation"], []).append(c)
    assert sorted(by_perm) == [2, 3, 5, 8]
    for num_perm, compares in by_perm.items():
        log = []
        r = SetwiseLlmRanker(None, None, "cuda", num_child=3, k=3, scoring="generation", method="heapsort", num_permutation=num_perm,
                             backend=_perm_backend(g, log))
        random.seed(929)
        for c in compares:
            label = r.compare(g["query"], [all_docs[i] for i in c["docs"]])
            assert log[-1]["input_ids"] == c["input_ids"], (num_perm, c["docs"])
            assert log[-1]["output"] == c["output"]
            assert label == c["label"], (num_perm, c["docs"])
            check_counters(r, c)
    capsys.readouterr()
    assert any(c["label"] == "Unexpected voting." for c in g["compares"]) and any(len(c["label"]) == 1 for c in g["compares"])


def test_setwise_rerank_with_permutation_voting_matches_reference(capsys):
    import random
    from llmrankers.setwise import SetwiseLlmRanker
    g = _perm_fixture()
    for c in g["reranks"]:
        log = []
        r = SetwiseLlmRanker(None, None, "cuda", num_child=c["num_child"], k=c["k"], scoring="generation", method=c["method"],
                             num_permutation=c["num_permutation"], backend=_perm_backend(g, log))
        random.seed(929)
        if "raises" in c:
            with pytest.raises(Exception) as ei:
                r.rerank(g["query"], docs_from(g["docs"][:c["n"]]))
            assert type(ei.value).__name__ == c["raises"]
        else:
            out = r.rerank(g["query"], docs_from(g["docs"][:c["n"]]))
            assert [d.docid for d in out] == c["order"], c
            assert [d.score for d in out] == c["scores"]
        assert len(log) == c["n_generate_calls"]
        check_counters(r, c)
    capsys.readouterr()


# ------------------------------------------------------------------------------------------- level-parallel heap build (§8f-2)
@pytest.mark.parametrize("n,c,k", [(100, 10, 10), (100, 3, 10), (37, 2, 5), (12, 3, 3), (5, 10, 3), (1, 3, 1), (64, 4, 64)])
def test_batched_heap_equals_sequential_heap(n, c, k):
    """Same final array, same multiset of compares as the reference's sequential heapify order, for a NON-transitive, noisy
    comparison (the sort must not rely on the comparator being an order) — and far fewer sequential rounds."""
    from llmrankers._sorting import heap_top_k, heap_top_k_batched
    rng = np.random.default_rng(n * 100 + c)
    noise = rng.standard_normal((n, n))
    vals = rng.permutation(n)

    def best(docs):   # deliberately context dependent: the winner depends on who else is in the set
        return int(np.argmax([vals[d] + noise[d, docs[0]] for d in docs]))

    seq_calls, bat_calls, rounds = [], [], []
    a = list(range(n))
    heap_top_k(a, c, k, lambda docs, inds: (seq_calls.append(tuple(docs)), inds[best(docs)])[1])
    b = list(range(n))

    def many(reqs):
        rounds.append(len(reqs))
        bat_calls.extend(tuple(docs) for docs, _ in reqs)
        return [inds[best(docs)] for docs, inds in reqs]
    heap_top_k_batched(b, c, k, many)
    assert a == b
    assert sorted(seq_calls) == sorted(bat_calls)
    assert len(rounds) <= len(seq_calls) and (n < 20 or len(rounds) < len(seq_calls))


@pytest.mark.parametrize("n,k", [(100, 10), (37, 5), (6, 3), (2, 1), (1, 1), (33, 33)])
def test_batched_binary_heap_equals_sequential(n, k):
    from llmrankers._sorting import binary_heap_top_k, binary_heap_top_k_batched
    rng = np.random.default_rng(n)
    noise = rng.standard_normal((n, n))
    vals = rng.permutation(n).astype(float)
    greater = lambda a, b: bool(vals[a] + noise[a, b] > vals[b])   # noisy, non-transitive
    seq, bat, rounds = [], [], []
    a = list(range(n))
    binary_heap_top_k(a, k, lambda x, y: (seq.append((x, y)), greater(x, y))[1])
    b = list(range(n))

    def many(pairs):
        rounds.append(len(pairs))
        bat.extend(pairs)
        return [greater(x, y) for x, y in pairs]
    binary_heap_top_k_batched(b, k, many)
    assert a == b and sorted(seq) == sorted(bat) and len(rounds) <= len(seq)


def test_pairwise_batched_and_sequential_heapsort_agree(monkeypatch):
    from llmrankers.pairwise import PairwiseLlmRanker
    meta = golden_meta()
    m, c = meta["tiny"], meta["cases"]["pairwise_heap"]
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("B200RANK_BATCHED_SORT", flag)
        r = PairwiseLlmRanker(None, None, "cuda", method="heapsort", batch_size=c["batch_size"], k=c["k"], backend=backend("tiny", True))
        out = r.rerank(m["query"], docs_from(m["docs12"][:6]))
        outs.append(([d.docid for d in out], [d.score for d in out], r.total_compare, r.total_prompt_tokens, r.total_completion_tokens))
    assert outs[0] == outs[1]
    assert outs[0][0] == c["order"] and outs[0][2:] == (c["total_compare"], c["total_prompt_tokens"], c["total_completion_tokens"])


@pytest.mark.parametrize("sort,case", [("heapsort", "pairwise_heap"), ("bubblesort", "pairwise_bubble")])
def test_pairwise_rerank_many_equals_rerank(sort, case):
    from llmrankers.pairwise import PairwiseLlmRanker
    meta = golden_meta()
    m, c = meta["tiny"], meta["cases"][case]
    mk = lambda method=sort: PairwiseLlmRanker(None, None, "cuda", method=method, batch_size=c["batch_size"], k=c["k"], backend=backend("tiny", True))
    d12 = m["docs12"]
    requests = [(m["query"], d12[:6]), ("w7 w8", d12[4:9]), ("w1", d12[:1]), (m["query"], d12[::-1][:7]), ("w4", [])]
    want = []
    for q, dd in requests:
        r = mk()
        out = r.rerank(q, docs_from(dd))
        want.append(([d.docid for d in out], [d.score for d in out], r.total_compare, r.total_prompt_tokens, r.total_completion_tokens))
    assert want[0][0] == c["order"] and want[0][2] == c["total_compare"]
    for window in (1, 2, 8):
        r = mk()
        got = []
        for out in r.rerank_many([(q, docs_from(dd)) for q, dd in requests], window=window):
            got.append(([d.docid for d in out], [d.score for d in out], r.total_compare, r.total_prompt_tokens, r.total_completion_tokens))
        assert got == want, window
    r = mk("allpair")   # falls back to rerank()
    outs = list(r.rerank_many([(m["query"], docs_from(d12[:4]))]))
    assert [d.docid for d in outs[0]] == [d.docid for d in mk("allpair").rerank(m["query"], docs_from(d12[:4]))]


@pytest.mark.parametrize("case", ["setwise_heap_gen", "setwise_heap_lik", "setwise_bubble_lik", "setwise_bubble_gen"])
def test_setwise_rerank_many_equals_rerank(case, capsys):
    """Cross-query lockstep (rerank_many) must reproduce rerank() per query: order, scores, counters — including queries of
    different sizes entering and leaving the window at different times."""
    from llmrankers.setwise import SetwiseLlmRanker
    meta = golden_meta()
    m, c = meta["tiny"], meta["cases"][case]
    mk = lambda: SetwiseLlmRanker(None, None, "cuda", num_child=c["num_child"], k=c["k"], scoring=c["scoring"], method=c["method"],
                                  backend=backend("tiny", c["label_favouring"]))
    docs12 = m["docs12"]
    requests = [(m["query"], docs12), ("w7 w8", docs12[:5]), ("w1 w2 w3", docs12[3:]), (m["query"], docs12[:4]), ("w9", docs12[::-1])]
    if c["method"] == "heapsort":
        requests += [(m["query"], docs12[:1]), ("w4", [])]
    want = []
    for q, dd in requests:
        r = mk()
        out = r.rerank(q, docs_from(dd))
        want.append(([d.docid for d in out], [d.score for d in out], r.total_compare, r.total_prompt_tokens, r.total_completion_tokens))
    assert want[0][0] == c["order"] and want[0][2] == c["total_compare"]
    for window in (1, 3, 8):
        r = mk()
        got = []
        for out in r.rerank_many([(q, docs_from(dd)) for q, dd in requests], window=window):
            got.append(([d.docid for d in out], [d.score for d in out], r.total_compare, r.total_prompt_tokens, r.total_completion_tokens))
        assert got == want, window


@pytest.mark.parametrize("case", ["setwise_heap_gen", "setwise_heap_lik"])
def test_setwise_batched_and_sequential_heapsort_agree(case, monkeypatch, capsys):
    from llmrankers.setwise import SetwiseLlmRanker
    meta = golden_meta()
    m, c = meta["tiny"], meta["cases"][case]
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("B200RANK_BATCHED_SORT", flag)
        r = SetwiseLlmRanker(None, None, "cuda", num_child=c["num_child"], k=c["k"], scoring=c["scoring"], method="heapsort",
                             backend=backend("tiny", c["label_favouring"]))
        out = r.rerank(m["query"], docs_from(m["docs12"]))
        outs.append(([d.docid for d in out], [d.score for d in out], r.total_compare, r.total_prompt_tokens, r.total_completion_tokens))
    assert outs[0] == outs[1]
    assert outs[0][0] == c["order"] and outs[0][2:] == (c["total_compare"], c["total_prompt_tokens"], c["total_completion_tokens"])


@pytest.mark.parametrize("mode", ["ones", "infer"])
def test_merged_generate_batches_equals_batch_by_batch(mode, monkeypatch):
    """T5Backend.generate_batches (many reference batches in one engine call, per-row lengths carrying each batch's own padding
    semantics) against T5Backend.generate called batch by batch — on an engine stand-in that runs the oracle row by row."""
    from llmrankers._backend import T5Backend
    monkeypatch.setenv("B200RANK_GENERATE_MASK", mode)
    orc = oracle_for("tiny", True)
    cfg, _ = model_and_weights("tiny", True)

    class RowWiseEngine:   # same contract as b200rank.Engine.greedy: right-padded ids + true lengths, rows independent
        def greedy(self, ids, lengths, dec_prefix, max_new):
            out = np.zeros((ids.shape[0], max_new), np.int32)
            for r in range(ids.shape[0]):
                row = np.asarray(ids[r:r + 1, : int(lengths[r])], np.int64)
                out[r] = orc.greedy(row, np.ones_like(row), dec_prefix, max_new, 1, 0)[0]
            return out

    be = T5Backend(RowWiseEngine(), tokenizer(), cfg)
    rng = np.random.default_rng(11)
    prefix = tokenizer().encode("<pad> Passage", add_special_tokens=False)
    batches = []
    for n, lo, hi in ((2, 5, 9), (2, 12, 12), (1, 7, 7), (3, 4, 15)):
        rows = [rng.integers(3, 2000, int(rng.integers(lo, hi + 1))).tolist() + [1] for _ in range(n)]
        batches.append(be.pad_rows(rows, 0)[0])
    want = [be.generate(b, prefix, 2) for b in batches]
    got = be.generate_batches(batches, prefix, 2)
    assert len(got) == len(want) and all(np.array_equal(g, w) for g, w in zip(got, want))
    assert be.generate_batches([], prefix, 2) == []


# ------------------------------------------------------------------------------------------- token-level prompt assembly (§8f-1)
def test_prompt_assembler_equals_whole_string_tokenisation():
    from llmrankers._prompts import PromptAssembler
    from llmrankers.pairwise import PAIRWISE_PROMPT
    from llmrankers.pointwise import MONOT5_PROMPT, QLM_PROMPT, YES_NO_PROMPT
    tok = tokenizer()
    rng = np.random.default_rng(3)
    texts = [" ".join(f"w{int(x)}" for x in rng.integers(0, 2000, int(rng.integers(1, 40)))) for _ in range(30)]
    texts += ["", " ", "  leading and trailing  ", "new\nline\tand tab", "Punct, here! (yes?) 'q' \"dq\"", "w1  w2   w3", "unknown-\u00e9\u4e2d words"]
    query = "w5  w6 Query: Passage"
    for template, fields in ((YES_NO_PROMPT, [dict(text=t, query=query) for t in texts]), (QLM_PROMPT, [dict(text=t) for t in texts]),
                             (MONOT5_PROMPT, [dict(query=query, document=t) for t in texts])):
        a = PromptAssembler(tok, template, verify=0)   # verify=0: no safety net, the assembly itself must be right
        assert a.eligible
        want = tok([template.format(**f) for f in fields])["input_ids"]
        assert a.rows(fields) == want
        assert a.rows(fields) == want and a.hits > 0          # second time from the cache
    # punctuation glued to a field (the quoted passages of the pairwise / setwise prompts) joins the cached unit
    a = PromptAssembler(tok, PAIRWISE_PROMPT, verify=0)
    assert a.eligible
    f = [dict(query=query, doc1=texts[i], doc2=texts[j]) for i in range(len(texts)) for j in (0, len(texts) - 1, len(texts) - 3, 31, 32)]
    assert a.rows(f) == tok([PAIRWISE_PROMPT.format(**x) for x in f])["input_ids"]
    assert len(a.tc.data) <= len(set(texts)) + 1              # one cached unit per distinct quoted passage + the quoted query
    assert a.rows(f[:10]) == tok([PAIRWISE_PROMPT.format(**x) for x in f[:10]])["input_ids"] and a.hits >= 30
    setwise_like = 'Given a query "{query}", which?\n\nPassage A: "{d0}"\n\nPassage B: "{d1}"\n\nOutput only the passage label:'
    shared = PromptAssembler(tok, setwise_like, verify=0, cache=a.tc)    # shares the pairwise assembler's token cache
    f = [dict(query=query, d0=texts[i], d1=texts[-1 - i]) for i in range(len(texts))]
    assert shared.rows(f) == tok([setwise_like.format(**x) for x in f])["input_ids"]
    # not separable: two fields joined by punctuation only, or a format spec -> whole-string path (still correct)
    for template in ("Pair: {a}-{b} end", "Value: {a:>8} end"):
        bad = PromptAssembler(tok, template)
        assert not bad.eligible
        ff = [dict(a="w1 w2", b="w3")] if "{b}" in template else [dict(a="w1")]
        assert bad.rows(ff) == tok([template.format(**ff[0])])["input_ids"]
    assert PromptAssembler(tok, YES_NO_PROMPT).rows([]) == []
    warm = PromptAssembler(tok, PAIRWISE_PROMPT, verify=0)
    warm.warm("doc1", texts)
    n0 = len(warm.tc.data)
    warm.rows([dict(query=query, doc1=t, doc2=texts[0]) for t in texts[1:5]])
    assert len(warm.tc.data) == n0 + 1     # only the quoted query was new: doc1 / doc2 units are the same strings


def test_prompt_assembler_falls_back_when_the_tokenizer_is_not_word_local():
    """A tokenizer whose segmentation depends on the neighbouring word breaks the assembly argument: the built-in check must
    notice on the first prompts and switch to whole-string tokenisation for good."""
    from llmrankers._prompts import PromptAssembler

    class Contextual:
        eos_token_id = 1
        is_fast = False

        def _ids(self, text, specials):
            words = text.split()
            ids = [10 + (len(w) + (len(words[i - 1]) if i else 0)) % 50 for i, w in enumerate(words)]   # depends on the previous word
            return ids + ([1] if specials else [])

        def encode(self, text, add_special_tokens=True):
            return self._ids(text, add_special_tokens)

        def __call__(self, texts, add_special_tokens=True):
            return {"input_ids": [self._ids(t, add_special_tokens) for t in texts]}

    tok = Contextual()
    a = PromptAssembler(tok, "Passage: {text}\nQuery: {query}", verify=4)
    fields = [dict(text=f"abc de{'f' * i}", query="gh ijk") for i in range(6)]
    want = tok(["Passage: {text}\nQuery: {query}".format(**f) for f in fields])["input_ids"]
    assert a.rows(fields) == want and not a.eligible
    assert a.rows(fields) == want


def test_prompt_assembler_is_thread_safe():
    from concurrent.futures import ThreadPoolExecutor
    from llmrankers._prompts import PromptAssembler
    from llmrankers.pointwise import YES_NO_PROMPT
    tok = tokenizer()
    a = PromptAssembler(tok, YES_NO_PROMPT, cache_size=64)   # small cache: evictions race with lookups
    rng = np.random.default_rng(4)
    texts = [" ".join(f"w{int(x)}" for x in rng.integers(0, 2000, 20)) for _ in range(200)]
    jobs = [[dict(text=texts[int(i)], query=f"w{q}") for i in rng.integers(0, 200, 40)] for q in range(16)]
    want = [tok([YES_NO_PROMPT.format(**f) for f in job])["input_ids"] for job in jobs]
    with ThreadPoolExecutor(8) as pool:
        got = list(pool.map(a.rows, jobs))
    assert got == want


def test_unsupported_variants_fail_loudly():
    from llmrankers.listwise import OpenAiListwiseLlmRanker
    from llmrankers.setwise import OpenAiSetwiseLlmRanker, SetwiseLlmRanker
    for cls in (OpenAiListwiseLlmRanker, OpenAiSetwiseLlmRanker):
        with pytest.raises(NotImplementedError):
            cls("x", "y")
    r = SetwiseLlmRanker(None, None, "cuda", method="quicksort", backend=backend())
    with pytest.raises(NotImplementedError):
        r.rerank("w1", docs_from(golden_meta()["tiny"]["docs12"]))


def test_device_cpu_is_refused():
    from llmrankers.pointwise import PointwiseLlmRanker
    with pytest.raises(RuntimeError, match="no CPU path"):
        PointwiseLlmRanker("synthetic:t5-tiny", None, "cpu")


# ------------------------------------------------------------------------------------------- CLI
def test_cli_pointwise_end_to_end(tmp_path, monkeypatch, capsys):
    import run as cli_mod
    from llmrankers import _backend
    meta = golden_meta()
    m, c = meta["tiny"], meta["cases"]["yes_no"]
    monkeypatch.setattr(_backend.T5Backend, "load", classmethod(lambda cls, *a, **k: backend("tiny")))
    (tmp_path / "queries.tsv").write_text(f"q1\t{m['query']}\n")
    (tmp_path / "docs.tsv").write_text("".join(f"{d['docid']}\t{d['text']}\n" for d in m["docs"]))
    (tmp_path / "run.txt").write_text("".join(f"q1 Q0 {d['docid']} {i + 1} {d['score']} bm25\n" for i, d in enumerate(m["docs"])))
    out = tmp_path / "out.txt"
    cli_mod.cli(["run", "--model_name_or_path", "synthetic:t5-tiny", "--run_path", str(tmp_path / "run.txt"), "--save_path", str(out),
                 "--queries_tsv", str(tmp_path / "queries.tsv"), "--collection_tsv", str(tmp_path / "docs.tsv"),
                 "--query_length", "32", "--passage_length", "128", "pointwise", "--method", "yes_no", "--batch_size", "4"])
    lines = out.read_text().splitlines()
    assert [l.split("\t")[2] for l in lines] == c["order"]
    assert all(re.fullmatch(r"q1\tQ0\td\d+\t\d+\t[-0-9.e]+\tLLMRankers", l) for l in lines)
    printed = capsys.readouterr().out
    assert f"Avg comparisons: {float(c['total_compare'])}" in printed and "Avg time per query:" in printed


@pytest.mark.parametrize("sub,case,extra", [("setwise", "setwise_heap_lik", ["--num_child", "3", "--method", "heapsort", "--k", "3"]),
                                            ("setwise", "setwise_bubble_lik", ["--num_child", "3", "--method", "bubblesort", "--k", "3"]),
                                            ("pairwise", "pairwise_heap", ["--method", "heapsort", "--batch_size", "2", "--k", "3"]),
                                            ("pairwise", "pairwise_allpair", ["--method", "allpair", "--batch_size", "4", "--k", "3"])])
def test_cli_sort_rankers_end_to_end(sub, case, extra, tmp_path, monkeypatch, capsys):
    """run.py with the setwise / pairwise sub-commands (rerank_many drives the loop): run file and summary counters against the
    reference's own rerank() outputs for the same candidates."""
    import run as cli_mod
    from llmrankers import _backend
    meta = golden_meta()
    m, c = meta["tiny"], meta["cases"][case]
    lab = c.get("label_favouring", sub == "pairwise")
    monkeypatch.setattr(_backend.T5Backend, "load", classmethod(lambda cls, *a, **k: backend("tiny", lab)))
    docs = m["docs12"] if sub == "setwise" else m["docs12"][:6]
    (tmp_path / "queries.tsv").write_text(f"q1\t{m['query']}\n")
    (tmp_path / "docs.tsv").write_text("".join(f"{d['docid']}\t{d['text']}\n" for d in docs))
    (tmp_path / "run.txt").write_text("".join(f"q1 Q0 {d['docid']} {i + 1} {d['score']} bm25\n" for i, d in enumerate(docs)))
    out = tmp_path / "out.txt"
    scoring = ["--scoring", c["scoring"]] if sub == "setwise" else []
    cli_mod.cli(["run", "--model_name_or_path", "synthetic:t5-tiny", "--run_path", str(tmp_path / "run.txt"), "--save_path", str(out),
                 "--queries_tsv", str(tmp_path / "queries.tsv"), "--collection_tsv", str(tmp_path / "docs.tsv"),
                 "--query_length", "32", "--passage_length", "128"] + scoring + [sub] + extra)
    assert [l.split("\t")[2] for l in out.read_text().splitlines()] == c["order"]
    printed = capsys.readouterr().out
    assert f"Avg comparisons: {float(c['total_compare'])}" in printed
    assert f"Avg prompt tokens: {float(c['total_prompt_tokens'])}" in printed
    assert f"Avg completion tokens: {float(c['total_completion_tokens'])}" in printed


def test_cli_grammar_errors():
    import run as cli_mod
    with pytest.raises(ValueError):
        cli_mod.cli(["run", "--model_name_or_path", "x"])                      # no method sub-command
    with pytest.raises(ValueError):
        cli_mod.cli(["run", "--model_name_or_path", "x", "pointwise", "setwise"])  # two methods
    parser, commands = cli_mod.make_parser()
    args = cli_mod.parse_args(parser, commands, ["run", "--hits", "7", "setwise"])
    assert args.run.hits == 7 and args.run.query_length == 128 and args.run.device == "cuda" and args.run.scoring == "generation"
    assert (args.setwise.num_child, args.setwise.method, args.setwise.k, args.setwise.num_permutation) == (3, "heapsort", 10, 1)
    assert args.pointwise is None and args.pairwise is None


# ------------------------------------------------------------------------------------------- C-ABI (no compute without a GPU)
def test_library_exports_every_declared_symbol():
    import b200rank as br
    header = open(os.path.join(ROOT, "include", "b200rank.h")).read()
    declared = set(re.findall(r"\b(b200rank_[a-z0-9_]+)\s*\(", header))
    lib = br.load_library()
    assert declared == set(br.EXPORTED_SYMBOLS), declared ^ set(br.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert b"sm_100a" in lib.b200rank_version()
    assert ctypes.sizeof(br.Config) == 18 * 4


@pytest.mark.skipif(__import__("conftest").HAS_GPU, reason="checks the no-GPU failure mode")
def test_compute_entry_points_fail_loudly_without_gpu():
    import b200rank as br
    c = br.make_config(128, 2, 256, 2, 2, vocab_size=2304)
    with pytest.raises(br.B200RankError, match="no CUDA device|no CPU fallback"):
        br.Engine(c, 0)
    with pytest.raises(br.B200RankError):
        br.test_gemm(np.zeros((128, 64), np.float32), np.zeros((64, 64), np.float32))


def test_config_validation_messages():
    import b200rank as br
    lib = br.load_library()
    h = ctypes.c_void_p()
    bad = br.make_config(128, 2, 256, 2, 2, vocab_size=2304, d_kv=96)
    assert lib.b200rank_create(ctypes.byref(bad), 0, ctypes.byref(h)) == -1
    assert b"d_kv" in lib.b200rank_last_error()
    # every malformed configuration is refused by argument validation (B200RANK_ERR_ARG = -1) before the device is even looked at
    ok = dict(d_model=128, num_heads=2, d_ff=256, num_layers=2, num_decoder_layers=2)
    for bad_kw in (dict(d_model=0), dict(d_model=-64), dict(d_model=100), dict(d_model=8192), dict(d_ff=0), dict(d_ff=7), dict(num_heads=0),
                   dict(num_layers=0), dict(num_layers=100000), dict(num_decoder_layers=-3), dict(vocab_size=0), dict(vocab_size=-8),
                   dict(vocab_size=2 ** 31 - 8), dict(d_kv=0), dict(d_kv=32), dict(rel_buckets=0), dict(rel_buckets=33),
                   dict(rel_max_distance=4096), dict(rel_max_distance=0), dict(max_tokens=-1), dict(max_docs=-1), dict(max_dec_len=-1),
                   dict(max_logit_rows=-1), dict(layer_norm_eps=0.0), dict(pad_id=-1), dict(eos_id=10 ** 6)):
        kw = dict(ok, vocab_size=2304)
        kw.update(bad_kw)
        cfg = br.make_config(kw.pop("d_model"), kw.pop("num_heads"), kw.pop("d_ff"), kw.pop("num_layers"), kw.pop("num_decoder_layers"), **kw)
        assert lib.b200rank_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -1, bad_kw
        assert not h.value
    assert lib.b200rank_create(None, 0, None) == -1
    assert lib.b200rank_score_yes_no(None, None, None, 0, 0, 0, 0, None, None) == -1 and lib.b200rank_wait_yes_no(None, 0, None, None) == -1
    lib.b200rank_destroy(None)      # tolerated, like free(NULL)
    assert br.rel_bucket(-200, True) == 15 and br.rel_bucket(200, True) == 31 and br.rel_bucket(-16, False) == 16


def test_rel_bucket_table_matches_hf_golden():
    import b200rank as br
    from helpers import golden_npz
    g = golden_npz("buckets.npz")
    assert [br.rel_bucket(int(r), True) for r in g["rel"]] == g["bidirectional"].tolist()
    assert [br.rel_bucket(int(r), False) for r in g["rel"]] == g["unidirectional"].tolist()


def test_prompt_assembler_with_a_trained_subword_vocabulary():
    """§8f-1's open caveat, as far as it can be closed offline: a REAL sentencepiece unigram vocabulary (trained by
    tests/golden/make_spm_vocab.py; subword pieces, punctuation, digits — the model family of Flan-T5's spiece.model) behind the same
    transformers.T5Tokenizer class. Token-level assembly must equal whole-string tokenisation for every prompt template of the rankers
    on natural prose, with verify=0 (no safety net), including text with quotes, brackets, unicode and odd whitespace."""
    import json
    from transformers import T5Tokenizer
    from helpers import GOLDEN
    from llmrankers._prompts import PromptAssembler
    from llmrankers.pairwise import PAIRWISE_PROMPT
    from llmrankers.pointwise import MONOT5_PROMPT, QLM_PROMPT, YES_NO_PROMPT
    from llmrankers.setwise import SetwiseLlmRanker
    with open(os.path.join(GOLDEN, "spm_unigram_vocab.json")) as f:
        fx = json.load(f)
    tok = T5Tokenizer(vocab=[(p, s) for p, s in fx["pieces"]])
    assert (tok.pad_token_id, tok.eos_token_id, tok.unk_token_id) == (0, 1, 2)
    texts = list(fx["sample_sentences"])
    assert any(len(tok.tokenize(w)) > 1 for t in texts[:5] for w in t.split())        # the vocabulary really splits words
    texts += ['He said: "quoted, (bracketed) text" -- and left.', "it's 3.14159, isn't it? 100% sure; e.g. x=1/2", "café naïve 中文 — dash",
              "  leading, trailing   and   inner   runs  ", "tab\tand\nnewline\r\nmix", "", " ", "\"", "a\"b \"c\" d\"", "UPPER lower MiXeD snake_case camelCase"]
    query = 'how to "sort" a list, in-place?'
    cases = [(YES_NO_PROMPT, [dict(text=t, query=query) for t in texts]), (QLM_PROMPT, [dict(text=t) for t in texts]),
             (MONOT5_PROMPT, [dict(query=query, document=t) for t in texts]),
             (PAIRWISE_PROMPT, [dict(query=query, doc1=texts[i], doc2=texts[-1 - i]) for i in range(len(texts))]),
             (SetwiseLlmRanker._template(3, ["A", "B", "C"]), [dict(query=query, d0=texts[i], d1=texts[(i + 7) % len(texts)], d2=texts[-1 - i])
                                                               for i in range(len(texts))])]
    for template, fields in cases:
        a = PromptAssembler(tok, template, verify=0)
        assert a.eligible, template
        want = tok([template.format(**f) for f in fields])["input_ids"]
        got = a.rows(fields)
        bad = [i for i, (g, w) in enumerate(zip(got, want)) if g != w]
        assert not bad, (template[:30], fields[bad[0]])
        assert a.rows(fields) == want     # from the cache


@pytest.mark.parametrize("source", ["ir_datasets", "pyserini"])
def test_cli_reference_data_sources_with_stub_modules(source, tmp_path, monkeypatch, capsys):
    """run.py keeps the reference's two data sources (run.py:135-149, 165-172) behind lazy imports; neither library exists offline,
    so stand-ins with the same call surface are injected: queries from dataset.queries_iter() / get_topics(index + '-test'), document
    text from docs_store().get(docid) (title prepended when the record has one) / LuceneSearcher.doc(docid).raw() JSON."""
    import json
    import sys
    import types
    import run as cli_mod
    from llmrankers import _backend
    meta = golden_meta()
    m, c = meta["tiny"], meta["cases"]["yes_no"]
    monkeypatch.setattr(_backend.T5Backend, "load", classmethod(lambda cls, *a, **k: backend("tiny")))
    docs = {d["docid"]: d["text"] for d in m["docs"]}
    # split every passage into a "title" (first word) and the rest: title + ' ' + text reproduces the fixture's passage
    split = {k: v.split(" ", 1) for k, v in docs.items()}
    if source == "ir_datasets":
        class Doc:
            def __init__(self, title, text):
                self.title, self.text = title, text

        class Store:
            def get(self, docid):
                return Doc(*split[docid])

        class Dataset:
            def queries_iter(self):
                yield types.SimpleNamespace(query_id="q1", text=m["query"])

            def docs_store(self):
                return Store()
        mod = types.ModuleType("ir_datasets")
        mod.load = lambda name: Dataset()
        monkeypatch.setitem(sys.modules, "ir_datasets", mod)
        src_args = ["--ir_dataset_name", "stub/dataset"]
    else:
        class Searcher:
            @classmethod
            def from_prebuilt_index(cls, name):
                assert name == "stub-index.flat"
                return cls()

            def doc(self, docid):
                return types.SimpleNamespace(raw=lambda: json.dumps({"title": split[docid][0], "text": split[docid][1]}))
        pkg, search = types.ModuleType("pyserini"), types.ModuleType("pyserini.search")
        base, lucene = types.ModuleType("pyserini.search._base"), types.ModuleType("pyserini.search.lucene")
        base.get_topics = lambda name: {"q1": {"title": m["query"]}} if name == "stub-index-test" else {}
        lucene.LuceneSearcher = Searcher
        for name, mod in (("pyserini", pkg), ("pyserini.search", search), ("pyserini.search._base", base), ("pyserini.search.lucene", lucene)):
            monkeypatch.setitem(sys.modules, name, mod)
        src_args = ["--pyserini_index", "stub-index"]
    (tmp_path / "run.txt").write_text("".join(f"q1 Q0 {d['docid']} {i + 1} {d['score']} bm25\n" for i, d in enumerate(m["docs"])))
    out = tmp_path / "out.txt"
    cli_mod.cli(["run", "--model_name_or_path", "synthetic:t5-tiny", "--run_path", str(tmp_path / "run.txt"), "--save_path", str(out)] + src_args +
                ["--query_length", "32", "--passage_length", "128", "pointwise", "--method", "yes_no", "--batch_size", "4"])
    assert [l.split("\t")[2] for l in out.read_text().splitlines()] == c["order"]
    assert f"Avg comparisons: {float(c['total_compare'])}" in capsys.readouterr().out


# ------------------------------------------------------------------------------------------- run.py against the reference's run.py
def _cli_fixture():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden_cli.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", [s["name"] for s in _cli_fixture()["scenarios"]])
def test_cli_matches_the_reference_run_py(name, tmp_path, monkeypatch, capsys):
    """tests/golden/golden_cli.json holds what the REFERENCE's run.py did, executed as __main__ with recording fakes for the ranker classes
    and stand-ins for ir_datasets / pyserini (tests/golden/make_golden_cli.py): which class it built with which keywords, the exact
    (query, [(docid, score, text)]) it handed to every rerank() — truncation, title prefix, --hits, --shuffle_ranking drawn from the
    module RNG interleaved with the ranker's own draws — its summary prints, its TREC file, or the error it raised. llm-rankers_b200/run.py
    must do the same under the same stand-ins."""
    import random
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_cli as G
    import run as cli_mod
    sc = next(s for s in _cli_fixture()["scenarios"] if s["name"] == name)
    for mod_name, mod in G.stub_source_modules().items():
        monkeypatch.setitem(sys.modules, mod_name, mod)
    log = []
    for names in G.RANKER_CLASSES.values():
        for n in names:
            monkeypatch.setattr(cli_mod, n, G.make_fake(n, log))
    (tmp_path / "first_stage.txt").write_text("\n".join(G.RUN_LINES) + "\n")
    save = tmp_path / "out.trec"
    argv = ["run", "--run_path", str(tmp_path / "first_stage.txt"), "--save_path", str(save)] + sc["argv"]
    random.seed(929)
    if "raises" in sc:
        with pytest.raises(Exception) as ei:
            cli_mod.cli(argv)
        assert [type(ei.value).__name__, str(ei.value)] == sc["raises"]
        return
    cli_mod.cli(argv)
    assert log == sc["log"]
    assert G.scrub(capsys.readouterr().out) == sc["stdout"]
    assert save.read_text() == sc["trec"]


# ------------------------------------------------------------------------------------------- pointwise rankers over a grid
def test_pointwise_rankers_match_the_reference_over_a_grid():
    """tests/golden/golden_pointwise_sweep.json: the reference's PointwiseLlmRanker (yes_no, qlm) and MonoT5LlmRanker run over list sizes
    1..33 x batch sizes 1..32 with a stand-in model whose logits are a hash of each row's real token ids
    (tests/golden/make_golden_pointwise_sweep.py). Scores only agree if the drop-in classes hand the engine exactly the token rows the
    reference handed the model (prompt templating over odd whitespace / newlines / empty passages); counters follow the reference's
    per-batch accounting; the order is the reference's stable sort. The reference raises IndexError on an EMPTY ranking
    (`tokenizer([])`); the drop-in returns [] instead — the one deliberate difference, asserted here."""
    import json
    import sys
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_pointwise_sweep as G
    from llmrankers.pointwise import MonoT5LlmRanker, PointwiseLlmRanker
    from llmrankers.rankers import SearchResult
    with open(os.path.join(ROOT, "tests", "golden", "golden_pointwise_sweep.json")) as f:
        fx = json.load(f)
    b = backend()
    seen_rows = []

    def score_yes_no(rows, col_a, col_b):
        seen_rows.append([len(r) for r in rows])
        lg = torch.tensor([G.yes_no_logits(r, col_a, col_b) for r in rows], dtype=torch.float32).reshape(-1, 2)
        return lg.numpy(), torch.softmax(lg, dim=1)[:, 0].numpy()

    def score_qlm(rows, labels):
        lg = torch.from_numpy(G.full_logits_qlm(rows, list(labels)))
        lab = torch.tensor([list(labels)] * len(rows))
        ce = torch.nn.CrossEntropyLoss(reduction="none")(lg.view(-1, lg.size(-1)), lab.view(-1))
        return (-1 * ce.view(-1, lab.size(-1)).sum(dim=1)).numpy()
    b.score_yes_no, b.score_qlm = score_yes_no, score_qlm
    assert (b.tokenizer.encode("Yes", add_special_tokens=False)[0], b.tokenizer.encode("No", add_special_tokens=False)[0]) == (fx["yes_id"], fx["no_id"])
    assert len(fx["cases"]) == 84
    for c in fx["cases"]:
        cls = MonoT5LlmRanker if c["kind"] == "monot5" else PointwiseLlmRanker
        r = cls(None, None, "cuda", method="qlm" if c["kind"] == "qlm" else "yes_no", batch_size=c["batch_size"], backend=b)
        docs = [SearchResult(docid=f"d{i}", score=float(c["n"] - i), text=t) for i, t in enumerate(c["texts"])]
        out = r.rerank(c["query"], docs)
        if "raises" in c:
            assert c["n"] == 0 and c["raises"] == "IndexError" and out == []
            continue
        assert [d.docid for d in out] == [x[0] for x in c["result"]], c
        assert np.allclose([d.score for d in out], [x[1] for x in c["result"]], rtol=1e-6, atol=1e-6), c
        assert all(d.text is not None for d in out)           # pointwise returns the input objects, text intact (pointwise.py:125-129)
        check_counters(r, c)


# ------------------------------------------------------------------------------------------- pairwise allpair over a grid
@pytest.mark.parametrize("batched", ["1", "0"])
def test_pairwise_allpair_matches_the_reference_over_a_grid(batched, monkeypatch):
    """tests/golden/golden_allpair_sweep.json: the reference's allpair rerank over list sizes x batch sizes x k with a stand-in generate()
    hashing each row's non-pad token ids (tests/golden/make_golden_allpair_sweep.py): the DataLoader batch shapes, the verdict
    aggregation (wins, conflicts, the defaultdict's insertion order deciding ties, never-scoring documents in the original-order tail),
    scores, order and counters — with the reference batches merged into large engine calls (default) and one engine call per batch.
    With fewer than two documents the reference raises IndexError (`tokenizer([])`); the drop-in returns the trivial ranking."""
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_allpair_sweep as G
    from llmrankers.pairwise import PairwiseLlmRanker
    from llmrankers.rankers import SearchResult
    monkeypatch.setenv("B200RANK_BATCHED_SORT", batched)
    with open(os.path.join(ROOT, "tests", "golden", "golden_allpair_sweep.json")) as f:
        fx = json.load(f)
    assert len(fx["cases"]) == 126
    b = backend()
    shapes = []

    def generate_batches(batches, dec_prefix, max_new):
        assert max_new == 2 and list(dec_prefix) == [0, fx["passage_id"]]
        outs = []
        for ids in batches:
            shapes.append([int(ids.shape[0]), int(ids.shape[1])])
            outs.append(np.asarray(G.stub_generate(np.asarray(ids).tolist(), fx["passage_id"], fx["a_id"], fx["b_id"], fx["junk_id"]), np.int64))
        return outs
    b.generate_batches = generate_batches
    for c in fx["cases"]:
        r = PairwiseLlmRanker(None, None, "cuda", method="allpair", batch_size=c["batch_size"], k=c["k"], backend=b)
        docs = [SearchResult(docid=f"d{i}", score=float(c["n"] - i), text=t) for i, t in enumerate(c["texts"])]
        del shapes[:]
        out = r.rerank(c["query"], docs)
        if "raises" in c:
            assert c["n"] < 2 and c["raises"] == "IndexError"
            assert [(d.docid, d.score) for d in out] == [(f"d{i}", -(i + 1)) for i in range(c["n"])]
            continue
        assert [[d.docid, d.score] for d in out] == c["result"], c
        assert shapes == c["batches"], c
        check_counters(r, c)
