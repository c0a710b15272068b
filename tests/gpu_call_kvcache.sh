#!/bin/bash
# GPU call of round 2 (second session): the parity suite with the KV-cached greedy / fused embedding norm / single-pass vocabulary
# reductions, then same-box A/B lines for each of them and for the encoder-above-decoder stream priority.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tests/gpu_call_kvcache.sh r02k'
set -u
TAG=${1:-r02k}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 15 700 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log
cp $OUT/parity_report.json $OUT/${TAG}_parity_report.json 2>/dev/null
bash tests/gpu_call_ab.sh ${TAG}_head "B200RANK_PIPE_PRIORITY=0" "B200RANK_PIPE_PRIORITY=2" "B200RANK_FUSE_EMBED_NORM=0"
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 2), {k: v for k, v in d.items() if k in ("rerank_many", "sequential_order", "scores_sha1", "order_sha1")},
          {k[:40]: v for k, v in list(d.get("by_kernel_ms_per_step", {}).items())[-4:]})
except Exception as e:
    print(sys.argv[1], "unreadable:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
for v in regs multipass; do
  B200RANK_VOCAB_ROW=$v timeout -k 15 300 python bench.py --workload qlm --hits 1000 --steps 4 --warmup 3 > $OUT/${TAG}_qlm_$v.json 2> $OUT/${TAG}_qlm_$v.err; echo "qlm $v rc=$?"
  line $OUT/${TAG}_qlm_$v.json
done
for v in 1 0; do
  B200RANK_KV_CACHE=$v timeout -k 15 300 python bench.py --workload setwise --steps 8 --warmup 1 > $OUT/${TAG}_setwise_kv$v.json 2> $OUT/${TAG}_setwise_kv$v.err; echo "setwise kv=$v rc=$?"
  line $OUT/${TAG}_setwise_kv$v.json
done
ls -la $OUT | tail -12
