#!/bin/bash
# GPU call: the tests that touch the decoder-step graphs of b200rank_greedy / b200rank_logits_at, then the setwise and pairwise workloads
# with the graphs on and off (same box).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tests/gpu_call_stepgraph.sh r02g'
set -u
TAG=${1:-r02g}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 15 500 python -m pytest tests -m gpu -x -q -k "greedy or graph or variants or setwise or pairwise or likelihood or listwise or generation or duot5 or merged" > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 2), {k: (round(v["ms_per_query"], 2) if k == "rerank_many" else round(v["ms_per_step"], 2)) for k, v in d.items() if k in ("rerank_many", "sequential_order")}, d.get("order_sha1"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
for v in 1 0; do
  B200RANK_DEC_GRAPH=$v timeout -k 15 300 python bench.py --workload setwise --steps 8 --warmup 1 > $OUT/${TAG}_setwise_graph$v.json 2> $OUT/${TAG}_setwise_graph$v.err; echo "setwise graph=$v rc=$?"
  line $OUT/${TAG}_setwise_graph$v.json
done
for v in 1 0; do
  B200RANK_DEC_GRAPH=$v timeout -k 15 300 python bench.py --workload pairwise --hits 24 > $OUT/${TAG}_pairwise_graph$v.json 2> $OUT/${TAG}_pairwise_graph$v.err; echo "pairwise graph=$v rc=$?"
  line $OUT/${TAG}_pairwise_graph$v.json
done
