"""world_size-2 gloo tests of the N>1 host logic (sharding + gather); the NCCL weight broadcast itself needs GPUs and is
exercised by bench.py --gpus N on the box."""
import subprocess
import sys
import textwrap


from helpers import ROOT


def test_shard_bounds_cover_exactly_once():
    from b200rank.dist import shard_bounds
    for n in (0, 1, 7, 100, 425700):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(r"{root}", "llm-rankers_b200")); sys.path.insert(0, r"{root}"); sys.path.insert(0, os.path.join(r"{root}", "tests"))
    from b200rank.dist import score_sharded, shard_bounds
    from helpers import calls, golden_meta, golden_npz, oracle_for
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    m = golden_meta()["tiny"]
    orc = oracle_for("tiny")
    rows = []
    for call in calls(golden_npz("golden_tiny.npz"), "yes_no"):
        for r, k in zip(call["input_ids"], call["attention_mask"]):
            rows.append(r[: int(k.sum())].tolist())
    def score(chunk):
        from oracle.t5_oracle import pad_batch
        ids, mask = pad_batch(chunk)
        return orc.score_yes_no(ids, mask, m["yes_id"], m["no_id"])[1]
    got = score_sharded(score, rows)
    want = score(rows)
    assert got.shape == want.shape and np.allclose(got, want, atol=1e-6), (got, want)
    # ragged per-rank lengths through the variable all-gather
    lo, hi = shard_bounds(len(rows), dist.get_rank(), 2)
    assert hi - lo == 5
    dist.destroy_process_group()
    print("RANK_OK", sys.argv[1])
""")


def test_two_rank_sharded_scoring_matches_single_rank(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=29533))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"RANK_OK {r}" in o, o[-2000:]


def _free_port() -> int:
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _same_run(a: str, b: str):
    """Two TREC run files agree: same (qid, docid, rank) lines; scores to 1e-5 — the numpy oracle that stands in for the engine in this
    CPU test is not bit-invariant to how rows are batched (BLAS blocking), the engine is (tests/test_engine_gpu.py)."""
    la, lb = a.splitlines(), b.splitlines()
    assert len(la) == len(lb) and la
    for x, y in zip(la, lb):
        fx, fy = x.split("\t"), y.split("\t")
        assert fx[:4] == fy[:4] and fx[5] == fy[5], (x, y)
        assert abs(float(fx[4]) - float(fy[4])) < 1e-5, (x, y)


CLI_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, os.path.join(r"{root}", "llm-rankers_b200")); sys.path.insert(0, r"{root}"); sys.path.insert(0, os.path.join(r"{root}", "tests"))
    os.environ.update(RANK=sys.argv[1], LOCAL_RANK=sys.argv[1], WORLD_SIZE=sys.argv[2], MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[3])
    import torch
    torch.cuda.is_available = lambda: False            # CPU box / CPU test: rendezvous on gloo
    import run as cli_mod
    from llmrankers import _backend
    from fake_backend import OracleBackend
    from helpers import model_and_weights, oracle_for
    from b200rank.synthetic import synthetic_tokenizer
    cfg, _ = model_and_weights("tiny")
    be = OracleBackend(oracle_for("tiny"), synthetic_tokenizer(), cfg)
    _backend.T5Backend.load = classmethod(lambda cls, *a, **k: be)    # the oracle stands in for the GPU engine (host logic under test)
    d = r"{tmp}"
    cli_mod.cli(["run", "--model_name_or_path", "synthetic:t5-tiny", "--run_path", d + "/run.txt", "--save_path", d + "/out_w" + sys.argv[2] + ".txt",
                 "--queries_tsv", d + "/queries.tsv", "--collection_tsv", d + "/docs.tsv", "--query_length", "32", "--passage_length", "128"]
                + {ranker_args})
    print("RANK_OK", sys.argv[1])
""")


def test_flat_document_pieces_tile_the_list_on_batch_boundaries():
    import run as cli_mod
    rng = __import__("numpy").random.default_rng(0)
    for world in (1, 2, 3, 8):
        for bs in (1, 4, 32):
            sizes = [int(x) for x in rng.integers(0, 120, size=9)]
            rankings = [list(range(n)) for n in sizes]
            seen = []
            per_rank = []
            for r in range(world):
                pieces = cli_mod.flat_document_pieces(rankings, bs, r, world)
                per_rank.append(sum(s1 - s0 for _, s0, s1 in pieces))
                for qi, s0, s1 in pieces:
                    assert s0 % bs == 0 and (s1 % bs == 0 or s1 == sizes[qi])      # cuts only between reference batches
                    seen += [(qi, i) for i in range(s0, s1)]
            assert seen == [(qi, i) for qi, n in enumerate(sizes) for i in range(n)]   # every document exactly once, in order
            assert max(per_rank) - min(per_rank) <= 2 * bs                            # balanced to within a batch at each cut


def test_cli_shards_queries_over_ranks(tmp_path):
    """run.py under a 2-rank rendezvous. pointwise: DOCUMENT-level sharding — the flattened (query, document) list is cut in two on a
    reference-batch boundary (a query straddles the cut), scores are gathered once, rank 0 sorts and writes; the run file and the
    summary counters equal the single-process run."""
    from helpers import golden_meta
    m = golden_meta()["tiny"]
    queries = [("q1", m["query"]), ("q2", "w3 w4 w5"), ("q3", "w100 w7"), ("q4", "w9"), ("q5", "w1 w2 w3 w4")]
    (tmp_path / "queries.tsv").write_text("".join(f"{q}\t{t}\n" for q, t in queries))
    (tmp_path / "docs.tsv").write_text("".join(f"{d['docid']}\t{d['text']}\n" for d in m["docs"]))
    lines = []
    for qi, (q, _) in enumerate(queries):
        docs = m["docs"][qi:] + m["docs"][:qi]
        lines += [f"{q} Q0 {d['docid']} {i + 1} {10 - i} bm25\n" for i, d in enumerate(docs[: 10 - qi])]
    (tmp_path / "run.txt").write_text("".join(lines))
    script = tmp_path / "cli_worker.py"
    script.write_text(CLI_WORKER.format(root=ROOT, port=29541, tmp=str(tmp_path), ranker_args=repr(["pointwise", "--method", "yes_no", "--batch_size", "4"])))
    outs = {}
    for world in (1, 2):
        port = str(_free_port())
        procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), port], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                 for r in range(world)]
        logs = [p.communicate(timeout=900)[0] for p in procs]
        for r, (p, o) in enumerate(zip(procs, logs)):
            assert p.returncode == 0 and f"RANK_OK {r}" in o, o[-3000:]
        outs[world] = ((tmp_path / f"out_w{world}.txt").read_text(), [l for l in logs[0].splitlines() if l.startswith("Avg ") and "time" not in l])
        assert not any(l.startswith("Avg ") for o in logs[1:] for l in o.splitlines())   # only rank 0 reports
    _same_run(outs[1][0], outs[2][0])
    assert len(outs[1][0].splitlines()) == sum(10 - i for i in range(5))
    assert outs[1][1] == outs[2][1] and len(outs[1][1]) == 3


def _cli_world_1_vs_2(tmp_path, ranker_args, port, n_queries=3, hits=6):
    from helpers import golden_meta
    m = golden_meta()["tiny"]
    queries = [("q1", m["query"]), ("q2", "w3 w4 w5"), ("q3", "w100 w7")][:n_queries]
    (tmp_path / "queries.tsv").write_text("".join(f"{q}\t{t}\n" for q, t in queries))
    (tmp_path / "docs.tsv").write_text("".join(f"{d['docid']}\t{d['text']}\n" for d in m["docs"]))
    lines = []
    for qi, (q, _) in enumerate(queries):
        docs = m["docs"][qi:] + m["docs"][:qi]
        lines += [f"{q} Q0 {d['docid']} {i + 1} {10 - i} bm25\n" for i, d in enumerate(docs[: hits - qi])]
    (tmp_path / "run.txt").write_text("".join(lines))
    script = tmp_path / "cli_worker.py"
    script.write_text(CLI_WORKER.format(root=ROOT, port=port, tmp=str(tmp_path), ranker_args=repr(ranker_args)))
    outs = {}
    for world in (1, 2):
        port = str(_free_port())
        procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), port], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                 for r in range(world)]
        logs = [p.communicate(timeout=900)[0] for p in procs]
        for r, (p, o) in enumerate(zip(procs, logs)):
            assert p.returncode == 0 and f"RANK_OK {r}" in o, o[-3000:]
        outs[world] = ((tmp_path / f"out_w{world}.txt").read_text(), [l for l in logs[0].splitlines() if l.startswith("Avg ") and "time" not in l])
        assert not any(l.startswith("Avg ") for o in logs[1:] for l in o.splitlines())
    _same_run(outs[1][0], outs[2][0])
    assert outs[1][1] == outs[2][1] and len(outs[1][1]) == 3


def test_cli_pairwise_allpair_shards_prompts_inside_the_backend(tmp_path):
    """pairwise allpair under 2 ranks: every rank walks all queries, each query's n(n-1) prompts are split over the ranks by
    ShardedBackend.generate_batches (whole reference batches per rank) and gathered per query; identical run file and counters."""
    _cli_world_1_vs_2(tmp_path, ["pairwise", "--method", "allpair", "--batch_size", "2", "--k", "3"], 29547)


def test_cli_setwise_heapsort_keeps_query_level_sharding(tmp_path):
    _cli_world_1_vs_2(tmp_path, ["setwise", "--method", "heapsort", "--num_child", "2", "--k", "3"], 29551)


def test_cli_pointwise_qlm_flat_split(tmp_path):
    _cli_world_1_vs_2(tmp_path, ["pointwise", "--method", "qlm", "--batch_size", "2"], 29555)
