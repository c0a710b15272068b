"""Generates tests/golden/golden_sorts.json: the reference's OWN sort drivers (ielab/llm-rankers llmrankers/setwise.py:200-313 heapify /
heapSort / bubblesort / output assembly, llmrankers/pairwise.py:133-162,221-290) run against a deterministic stand-in for the LLM
compare, over a grid of list sizes, fan-outs, k and misbehaving outputs (labels beyond the compared set, junk strings, conflicting
pair verdicts). The model-driven fixtures of make_golden.py pin the full pipeline on a handful of configurations — and a random-init
model mostly answers with the fallback label — so the integer logic of the drivers (index arithmetic, fallback rules, early exits,
the 'skip the unchanged tail' bookkeeping, k > n, n < fan-out, empty lists) is pinned here, exhaustively and cheaply.

Recorded per case: the exact SEQUENCE of compare calls (doc ids per call), the returned ranking [(docid, score)], total_compare — or
the exception type if the reference raises. tests/test_sort_fixtures.py holds llm-rankers_b200/llmrankers to them: the sequential
drivers call for call, the level-parallel / cross-query drivers on the final ranking and the multiset of calls.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_sorts.py
The judges below are imported by the test as well (this module touches /root/reference only inside main())."""
import hashlib
import json
import os
import sys

CHARACTERS = ["A", "B", "C", "D", "E", "F", "G", "H", "I", "J", "K", "L", "M", "N", "O", "P", "Q", "R", "S", "T", "U", "V", "W"]


def _h(*key) -> int:
    return int.from_bytes(hashlib.sha1(repr(key).encode()).digest()[:8], "big")


def setwise_label(seed: int, p_bad: float, query: str, docids) -> str:
    """The stand-in model's answer for one compare: a pure function of (seed, query, compared doc ids) — the same set always gets the
    same verdict, whatever the order the driver asks in."""
    h = _h(seed, query, list(docids))
    if ((h >> 24) % 1000) / 1000.0 < p_bad:
        beyond = CHARACTERS[min(len(docids) + (h >> 8) % 3, len(CHARACTERS) - 1)]   # a label past the compared set
        return ["?", "AB", beyond, ""][(h >> 16) % 4]
    return CHARACTERS[h % max(1, len(docids))]


def pairwise_verdict(seed: int, p_bad: float, query: str, text1: str, text2: str):
    """The two decoded strings of a pairwise compare (both presentation orders), pure in (seed, query, text1, text2)."""
    h = _h(seed, query, text1, text2)
    if ((h >> 24) % 1000) / 1000.0 < p_bad:
        return [["Passage A", "Passage A"], ["Passage B", "Passage B"], ["x", ""], ["Passage A", "Passage"]][(h >> 16) % 4]
    return ["Passage A", "Passage B"] if h % 2 else ["Passage B", "Passage A"]


def make_docs(n: int, SearchResult):
    return [SearchResult(docid=f"d{i}", score=float(n - i), text=f"text of d{i}") for i in range(n)]


SETWISE_GRID = [(n, c, k) for n in (0, 1, 2, 3, 5, 12, 37, 100) for c in (2, 3, 10) for k in (1, 3, 10)] + [(7, 3, 20), (4, 10, 10), (23, 22, 5)]
PAIRWISE_GRID = [(n, k) for n in (0, 1, 2, 3, 5, 12, 37) for k in (1, 3, 10)] + [(6, 20)]


def cases():
    for n, c, k in SETWISE_GRID:
        for method in ("heapsort", "bubblesort"):
            for p_bad in (0.0, 0.2):
                yield dict(kind="setwise", method=method, n=n, num_child=c, k=k, p_bad=p_bad, seed=_h(n, c, k, method, p_bad) % 10000,
                           query=f"q{n}-{c}-{k}")
    for n, k in PAIRWISE_GRID:
        for method in ("heapsort", "bubblesort"):
            for p_bad in (0.0, 0.2):
                yield dict(kind="pairwise", method=method, n=n, k=k, p_bad=p_bad, seed=_h(n, k, method, p_bad) % 10000, query=f"q{n}-{k}")


def main():
    sys.path.insert(0, "/root/reference")
    from types import SimpleNamespace
    from llmrankers.pairwise import PairwiseLlmRanker
    from llmrankers.rankers import SearchResult
    from llmrankers.setwise import SetwiseLlmRanker
    assert SetwiseLlmRanker.CHARACTERS == CHARACTERS
    out = []
    for c in cases():
        calls = []
        if c["kind"] == "setwise":
            r = SetwiseLlmRanker.__new__(SetwiseLlmRanker)   # constructor needs the hub / accelerate
            r.num_child, r.k, r.method, r.num_permutation, r.scoring = c["num_child"], c["k"], c["method"], 1, "generation"
            r.config = SimpleNamespace(model_type="t5")

            def compare(query, docs, c=c, calls=calls):
                ids = [d.docid for d in docs]
                calls.append(ids)
                return setwise_label(c["seed"], c["p_bad"], query, ids)
        else:
            r = PairwiseLlmRanker.__new__(PairwiseLlmRanker)
            r.k, r.method = c["k"], c["method"]
            r.config = SimpleNamespace(model_type="t5")

            def compare(query, docs, c=c, calls=calls):
                calls.append([docs[0], docs[1]])
                return pairwise_verdict(c["seed"], c["p_bad"], query, docs[0], docs[1])
        r.compare = compare
        r.total_compare = r.total_prompt_tokens = r.total_completion_tokens = 0
        rec = dict(c)
        try:
            res = r.rerank(c["query"], make_docs(c["n"], SearchResult))
            rec["result"] = [[d.docid, d.score] for d in res]
        except Exception as e:   # noqa: BLE001 - the exception type IS the recorded behaviour
            rec["raises"] = type(e).__name__
        rec["calls"] = calls
        out.append(rec)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_sorts.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    n_raise = sum("raises" in r for r in out)
    print(f"wrote {len(out)} cases ({n_raise} raising, {sum(len(r['calls']) for r in out)} compare calls) to {path}, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
