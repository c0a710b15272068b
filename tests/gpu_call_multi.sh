#!/bin/bash
# Multi-GPU call (gpurun --gpus N): strong-scaling legs of the secondary workloads + the headline weak-scaling line.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tests/gpu_call_multi.sh r02m 2 48 1000'        (N = 1 and N = 2 legs on one box)
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tests/gpu_call_multi.sh r02n 8 100 1000 only'   (only the N = 8 legs: the N = 1 legs
#                                                                                                              come from a 1-GPU call, 8x cheaper)
set -u
TAG=$1; N=$2; PW_HITS=${3:-100}; QLM_HITS=${4:-1000}; ONLY=${5:-}; OUT=gpurun_out; mkdir -p $OUT
run() { name=$1; shift; timeout -k 15 600 python bench.py "$@" > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err; echo "$name rc=$?"; tail -c 600 $OUT/${TAG}_$name.json | cut -c1-600; echo; }
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for g in $([ -n "$ONLY" ] && echo $N || echo 1 $N); do
  run pairwise_n$g --workload pairwise --gpus $g --hits $PW_HITS
  run qlm_n$g --workload qlm --gpus $g --hits $QLM_HITS --steps 5 --warmup 3
done
[ $N -gt 1 ] && run headline_n$N --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-text-api --no-hf-cuda
python - <<PY
import json
for name in ("pairwise", "qlm"):
    try:
        a = json.loads(open("$OUT/${TAG}_%s_n1.json" % name).read().strip().splitlines()[-1])
        b = json.loads(open("$OUT/${TAG}_%s_n$N.json" % name).read().strip().splitlines()[-1])
        key = "order_sha1" if name == "pairwise" else "scores_sha1"
        print(name, "N=1", round(a["value"], 1), a["unit"], "N=$N", round(b["value"], 1), "speed-up", round(b["value"] / a["value"], 2), "efficiency", round(b["value"] / a["value"] / $N, 3),
              "identical results:", a[key] == b[key], b["config"].get("broadcast"))
    except Exception as e:
        print(name, "unreadable:", e)
PY
