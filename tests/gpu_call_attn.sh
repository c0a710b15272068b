#!/bin/bash
# GPU call: validate + A/B + profile a candidate encoder-attention kernel (B200RANK_ATTN=<mode>) against the default.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tests/gpu_call_attn.sh r02b tc2 tc5'      (A/B default vs tc2, ncu --set full of tc5)
set -u
TAG=${1:-rXX}; MODE=${2:-tc2}; PROF=${3:-tc5}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 15 400 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -k "enc_attention or headline or checkpoint_directory" > $OUT/${TAG}_pytest_attn.log 2>&1; echo "pytest attention rc=$?"
tail -15 $OUT/${TAG}_pytest_attn.log
for m in default $MODE; do
  if [ $m = default ]; then unset B200RANK_ATTN; else export B200RANK_ATTN=$m; fi
  timeout -k 15 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-text-api --no-hf-cuda --no-sustained > $OUT/${TAG}_bench_$m.json 2> $OUT/${TAG}_bench_$m.err; echo "bench $m rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$m.json").read().strip().splitlines()[-1])
    k = d["roofline"]["by_kernel_ms_per_step"]
    print("$m docs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {x: v for x, v in k.items() if "attention" in x}, "clocks", d["clocks"]["sm_mhz"])
except Exception as e:
    print("bench line unreadable:", e); print(open("$OUT/${TAG}_bench_$m.err").read()[-2000:])
PY
done
export B200RANK_ATTN=$PROF
timeout -k 15 300 ncu --set full --clock-control none --import-source on -k regex:enc_attention_ -s 30 -c 1 -f -o $OUT/${TAG}_attn_$PROF \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-text-api --no-hf-cuda --no-sustained > $OUT/${TAG}_ncu_attn_$PROF.log 2>&1; echo "ncu rc=$?"
ls -la $OUT | tail -8
