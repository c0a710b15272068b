"""world_size-2 gloo tests of the N>1 host logic (sharding + gather); the NCCL weight broadcast itself needs GPUs and is
exercised by bench.py --gpus N on the box."""
import os
import subprocess
import sys
import textwrap

import numpy as np

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
