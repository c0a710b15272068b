#!/bin/bash
# GPU call: A/B a list of environment settings on the headline bench (short runs, kernel table only).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tests/gpu_call_ab.sh r02f "B200RANK_ATTN=tc5 B200RANK_ATTN_WAIT=0" "B200RANK_ATTN=tc5 B200RANK_ATTN_WAIT=1"'
set -u
TAG=$1; shift; OUT=gpurun_out; mkdir -p $OUT
i=0
for setting in "$@"; do
  i=$((i+1))
  env $setting timeout -k 15 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-text-api --no-hf-cuda --no-sustained > $OUT/${TAG}_ab$i.json 2> $OUT/${TAG}_ab$i.err; rc=$?
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_ab$i.json").read().strip().splitlines()[-1])
    k = d["roofline"]["by_kernel_ms_per_step"]
    print("[$setting] rc=$rc docs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {x.split(' M')[0][:28] + (' M' + x.split(' M')[1] if ' M' in x else ''): v for x, v in list(k.items())[:7]}, "clocks", d["clocks"]["sm_mhz"])
except Exception as e:
    print("[$setting] rc=$rc unreadable:", e); print(open("$OUT/${TAG}_ab$i.err").read()[-1500:])
PY
done
