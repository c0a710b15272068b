"""Generates tests/golden/golden_listwise_windows.json: the reference's sliding-window listwise rerank (llmrankers/listwise.py:177-195,
inherited by the T5 ListwiseLlmRanker) and its response handling (clean_response / remove_duplicate / receive_permutation, :110-144) run
against a deterministic stand-in for compare() over a grid of list sizes, window sizes, step sizes and repeat counts — including windows
larger than the list, steps larger than the window, num_repeat 0 and 2, and misbehaving responses (partial permutations, duplicated and
out-of-window identifiers, junk text, the 'ERROR::reduce_length' marker of the OpenAI branch).

The model-driven listwise fixtures (make_golden_listwise.py) pin the prompts, the forward and the parser on four configurations; this
sweep pins the window arithmetic and the permutation bookkeeping exhaustively. Recorded per case: the SEQUENCE of compare calls (doc
ids of every window), the returned [(docid, score)], total_compare.

    python tests/golden/make_golden_listwise_windows.py     (build container only: needs /root/reference)

tests/test_listwise.py replays the cases through llm-rankers_b200/llmrankers/listwise.py with the same stand-in (imported from here;
this module touches /root/reference only inside main())."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def _h(*key) -> int:
    return int.from_bytes(hashlib.sha1(repr(key).encode()).digest()[:8], "big")


def response_for(seed: int, p_bad: float, query: str, docids) -> str:
    """The stand-in model's answer for one window: a pure function of (seed, query, doc ids of the window)."""
    n = len(docids)
    h = _h(seed, query, list(docids))
    order = sorted(range(1, n + 1), key=lambda i: _h(h, i))
    full = " > ".join(f"[{i}]" for i in order)
    if ((h >> 24) % 1000) / 1000.0 >= p_bad:
        return full
    kind = (h >> 16) % 6
    if kind == 0:
        return " > ".join(f"[{i}]" for i in order[: max(1, n // 2)])                      # partial permutation
    if kind == 1:
        return " > ".join(f"[{i}]" for i in (order + order[:2]))                          # duplicated identifiers
    if kind == 2:
        return " > ".join(f"[{i}]" for i in ([n + 3, 0] + order))                         # identifiers outside the window
    if kind == 3:
        return "The most relevant passages are " + ", ".join(str(i) for i in order) + "."  # prose with bare numbers
    if kind == 4:
        return "no idea"                                                                  # nothing usable
    return "ERROR::reduce_length"


GRID = ([(n, w, s, 1) for n in (0, 1, 2, 3, 5, 10, 20) for w in (2, 3, 4, 10) for s in (1, 2, 3, 5)]
        + [(n, w, s, r) for n in (5, 12) for w in (3, 4) for s in (1, 2) for r in (0, 2, 3)])


def cases():
    for n, w, s, r in GRID:
        for p_bad in (0.0, 0.3):
            yield dict(n=n, window_size=w, step_size=s, num_repeat=r, p_bad=p_bad, seed=_h(n, w, s, r, p_bad) % 10000, query=f"q{n}-{w}-{s}-{r}")


def make_docs(n: int, SearchResult):
    return [SearchResult(docid=f"d{i}", score=float(n - i), text=f"text of d{i}") for i in range(n)]


def main():
    sys.path.insert(0, "/root/reference")
    from llmrankers.listwise import ListwiseLlmRanker
    from llmrankers.rankers import SearchResult
    assert "/root/reference" in sys.modules["llmrankers.listwise"].__file__
    out = []
    for c in cases():
        calls = []
        r = ListwiseLlmRanker.__new__(ListwiseLlmRanker)   # constructor needs the hub / accelerate
        r.window_size, r.step_size, r.num_repeat, r.scoring = c["window_size"], c["step_size"], c["num_repeat"], "generation"

        def compare(query, docs, c=c, calls=calls, r=r):
            r.total_compare += 1
            ids = [d.docid for d in docs]
            calls.append(ids)
            return response_for(c["seed"], c["p_bad"], query, ids)
        r.compare = compare
        rec = dict(c)
        docs = make_docs(c["n"], SearchResult)
        try:
            res = r.rerank(c["query"], docs)
            rec["result"] = [[d.docid, d.score] for d in res]
            rec["returns_input_objects"] = bool(res is docs)
            rec["input_scores_after"] = [d.score for d in docs]
            rec["total_compare"] = r.total_compare
        except Exception as e:   # noqa: BLE001 - the exception type IS the recorded behaviour
            rec["raises"] = type(e).__name__
        rec["calls"] = calls
        out.append(rec)
    path = os.path.join(HERE, "golden_listwise_windows.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(f"wrote {len(out)} cases ({sum('raises' in r for r in out)} raising, {sum(len(r['calls']) for r in out)} windows) to {path}, "
          f"{os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
