"""Generates tests/golden/golden_cli.json by executing THE REFERENCE'S run.py (/root/reference/run.py, unmodified, as __main__) for a set
of command lines, with
  * stand-ins for the two data-source libraries that do not exist offline (`ir_datasets`, `pyserini`: same call surface as run.py:135-149,
    165-172 uses, over a small in-memory corpus), and
  * recording fakes in place of the ranker classes (constructors need the hub): they log the constructor keywords, every
    rerank(query, ranking) call with the texts exactly as run.py prepared them (truncate(), title prefix, --hits cut, --shuffle_ranking,
    drawn from the module RNG seeded with 929), consume the RNG the way a permutation-voting setwise ranker does, and return a
    deterministic ranking and counters.
What is pinned is everything run.py itself does: the CLI grammar and defaults, which class is built with which keywords (monot5 / duot5 /
openai routing, the pairwise batch_size override), the data path from run file + sources to rerank() inputs, the interleaving of
random.shuffle with the ranker's own RNG use, the three summary prints and the TREC output file (run.py:41-49, 52-201, 204-258).

    python tests/golden/make_golden_cli.py        (build container only: needs /root/reference)

tests/test_host_logic.py replays every scenario through llm-rankers_b200/run.py with the same stand-ins (imported from this module,
which touches /root/reference only inside main())."""
import contextlib
import hashlib
import io
import json
import os
import random
import runpy
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))

QUERIES = {"q1": "alpha beta gamma delta epsilon zeta", "q2": "one two three", "7": "seven eight nine ten eleven twelve thirteen"}
CORPUS = {f"D{i}": dict(title=(f"Title{i}" if i % 3 else None), text=" ".join(f"t{i}_{j}" for j in range(4 + (i * 7) % 9))) for i in range(12)}
RUN_LINES = ([f"q1 Q0 D{i} {r + 1} {20.0 - r} bm25" for r, i in enumerate((3, 1, 4, 0, 5, 9, 2, 6))]
             + [f"q2 Q0 D{i} {r + 1} {9.5 - r} bm25" for r, i in enumerate((7, 8, 10, 11, 0))]
             + [f"7 Q0 D{i} {r + 1} {3.25 - r} bm25" for r, i in enumerate((2, 3))])

RANKER_CLASSES = {"llmrankers.pointwise": ["PointwiseLlmRanker", "MonoT5LlmRanker"],
                  "llmrankers.setwise": ["SetwiseLlmRanker", "OpenAiSetwiseLlmRanker"],
                  "llmrankers.pairwise": ["PairwiseLlmRanker", "DuoT5LlmRanker", "OpenAiPairwiseLlmRanker"],
                  "llmrankers.listwise": ["ListwiseLlmRanker", "OpenAiListwiseLlmRanker"]}

SCENARIOS = [
    dict(name="pointwise_defaults_ir", argv=["--ir_dataset_name", "stub/ds", "--model_name_or_path", "google/flan-t5-large", "pointwise"]),
    dict(name="pointwise_qlm_pyserini", argv=["--pyserini_index", "stub-index", "--model_name_or_path", "m", "--tokenizer_name_or_path", "tk",
                                               "--device", "cuda:1", "--cache_dir", "/c", "--hits", "3", "--query_length", "4",
                                               "--passage_length", "5", "pointwise", "--method", "qlm", "--batch_size", "16"]),
    dict(name="monot5_routing", argv=["--ir_dataset_name", "stub/ds", "--model_name_or_path", "castorini/monot5-base-msmarco", "pointwise",
                                       "--batch_size", "8"]),
    dict(name="setwise_perm_shuffle_random", argv=["--pyserini_index", "stub-index", "--model_name_or_path", "m", "--hits", "4", "--scoring",
                                                    "likelihood", "--shuffle_ranking", "random", "setwise", "--num_child", "2", "--k", "3",
                                                    "--num_permutation", "3", "--method", "bubblesort"]),
    dict(name="setwise_defaults_shuffle_random", argv=["--ir_dataset_name", "stub/ds", "--model_name_or_path", "m", "--shuffle_ranking", "random",
                                                        "setwise"]),
    dict(name="setwise_openai", argv=["--ir_dataset_name", "stub/ds", "--model_name_or_path", "gpt-x", "--openai_key", "sk-test", "setwise",
                                       "--num_child", "5"]),
    dict(name="pairwise_heapsort_inverse", argv=["--ir_dataset_name", "stub/ds", "--model_name_or_path", "m", "--hits", "3", "--shuffle_ranking",
                                                  "inverse", "pairwise", "--method", "heapsort", "--batch_size", "16", "--k", "2"]),
    dict(name="pairwise_allpair_batch", argv=["--pyserini_index", "stub-index", "--model_name_or_path", "m", "pairwise", "--batch_size", "8"]),
    dict(name="duot5_routing", argv=["--ir_dataset_name", "stub/ds", "--model_name_or_path", "castorini/duot5-base-msmarco", "pairwise", "--method",
                                      "bubblesort"]),
    dict(name="pairwise_openai", argv=["--pyserini_index", "stub-index", "--model_name_or_path", "gpt-x", "--openai_key", "sk", "pairwise"]),
    dict(name="listwise_defaults", argv=["--ir_dataset_name", "stub/ds", "--model_name_or_path", "m", "listwise"]),
    dict(name="listwise_flags", argv=["--pyserini_index", "stub-index", "--model_name_or_path", "m", "--scoring", "likelihood", "listwise",
                                       "--window_size", "4", "--step_size", "2", "--num_repeat", "3"]),
    dict(name="listwise_openai", argv=["--ir_dataset_name", "stub/ds", "--model_name_or_path", "gpt-x", "--openai_key", "sk", "listwise",
                                        "--window_size", "5"]),
    dict(name="error_both_sources", argv=["--ir_dataset_name", "a", "--pyserini_index", "b", "--model_name_or_path", "m", "pointwise"]),
    dict(name="error_two_methods", argv=["--ir_dataset_name", "a", "--model_name_or_path", "m", "pointwise", "setwise"]),
    dict(name="error_no_method", argv=["--ir_dataset_name", "a", "--model_name_or_path", "m"]),
]


def _h(*key) -> int:
    return int.from_bytes(hashlib.sha1(repr(key).encode()).digest()[:8], "big")


def make_fake(class_name: str, log: list):
    """A ranker class that records what run.py does with it."""

    class Fake:
        def __init__(self, **kw):
            self.kw = kw
            log.append(dict(event="construct", cls=class_name, kwargs={k: kw[k] for k in sorted(kw)}))
            self.total_compare = self.total_prompt_tokens = self.total_completion_tokens = 0

        def truncate(self, text, length):
            return " ".join(text.split()[:length])

        def rerank(self, query, ranking):
            log.append(dict(event="rerank", query=query, ranking=[[d.docid, d.score, d.text] for d in ranking]))
            draws = 0
            if self.kw.get("num_permutation", 1) > 1:   # a permutation-voting ranker draws from the module RNG inside rerank()
                draws = [random.random() for _ in range(self.kw["num_permutation"])]
                log[-1]["rng_draws"] = draws
            out = sorted(ranking, key=lambda d: _h(class_name, query, d.docid))
            res = [type(d)(docid=d.docid, score=-(i + 1), text=None) for i, d in enumerate(out)]
            self.total_compare = len(ranking)
            self.total_prompt_tokens = 10 * len(ranking) + len(query.split())
            self.total_completion_tokens = 2 * len(ranking)
            return res
    Fake.__name__ = class_name
    return Fake


def stub_source_modules():
    """{module name: module} standing in for ir_datasets and pyserini over CORPUS / QUERIES."""
    class Doc:
        def __init__(self, title, text):
            self.text = text
            if title is not None:
                self.title = title

    class Store:
        def get(self, docid):
            return Doc(CORPUS[docid]["title"], CORPUS[docid]["text"])

    class Dataset:
        def queries_iter(self):
            for qid, text in QUERIES.items():
                yield types.SimpleNamespace(query_id=qid, text=text)

        def docs_store(self):
            return Store()

    class Searcher:
        @classmethod
        def from_prebuilt_index(cls, name):
            assert name == "stub-index.flat", name
            return cls()

        def doc(self, docid):
            rec = {"text": CORPUS[docid]["text"]}
            if CORPUS[docid]["title"] is not None:
                rec["title"] = CORPUS[docid]["title"]
            return types.SimpleNamespace(raw=lambda: json.dumps(rec))

    def get_topics(name):
        assert name == "stub-index-test", name
        return {(int(q) if q.isdigit() else q): {"title": t} for q, t in QUERIES.items()}   # pyserini topic ids may be ints (run.py:147)
    ir = types.ModuleType("ir_datasets")
    ir.load = lambda name: Dataset()
    pkg, search = types.ModuleType("pyserini"), types.ModuleType("pyserini.search")
    base, lucene = types.ModuleType("pyserini.search._base"), types.ModuleType("pyserini.search.lucene")
    base.get_topics = get_topics
    lucene.LuceneSearcher = Searcher
    return {"ir_datasets": ir, "pyserini": pkg, "pyserini.search": search, "pyserini.search._base": base, "pyserini.search.lucene": lucene}


def scrub(stdout: str):
    """The summary prints without the wall-clock line."""
    return [l for l in stdout.splitlines() if l.startswith("Avg ") and not l.startswith("Avg time per query")]


def main():
    sys.path.insert(0, "/root/reference")
    import importlib
    for name, mod in stub_source_modules().items():
        sys.modules[name] = mod
    mods = {m: importlib.import_module(m) for m in RANKER_CLASSES}
    assert all("/root/reference" in m.__file__ for m in mods.values())
    out = dict(queries=QUERIES, corpus=CORPUS, run_lines=RUN_LINES, scenarios=[])
    with tempfile.TemporaryDirectory() as tmp:
        run_path = os.path.join(tmp, "first_stage.txt")
        with open(run_path, "w") as f:
            f.write("\n".join(RUN_LINES) + "\n")
        for sc in SCENARIOS:
            log = []
            for m, names in RANKER_CLASSES.items():
                for n in names:
                    setattr(mods[m], n, make_fake(n, log))
            save_path = os.path.join(tmp, sc["name"] + ".trec")
            argv = ["run.py", "run", "--run_path", run_path, "--save_path", save_path] + sc["argv"]
            rec = dict(name=sc["name"], argv=sc["argv"])
            random.seed(929)            # what `import random; random.seed(929)` at the top of run.py does on a fresh interpreter
            buf, old_argv = io.StringIO(), sys.argv
            try:
                sys.argv = argv
                with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
                    runpy.run_path("/root/reference/run.py", run_name="__main__")
                rec["stdout"] = scrub(buf.getvalue())
                with open(save_path) as f:
                    rec["trec"] = f.read()
            except Exception as e:   # noqa: BLE001 - the exception IS the recorded behaviour
                rec["raises"] = [type(e).__name__, str(e)]
            finally:
                sys.argv = old_argv
            rec["log"] = log
            out["scenarios"].append(rec)
    path = os.path.join(HERE, "golden_cli.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes")
    for s in out["scenarios"]:
        print(s["name"], s.get("raises") or (s["log"][0]["cls"], len(s["log"]) - 1, s["stdout"][0]))


if __name__ == "__main__":
    main()
