#!/bin/bash
# Weak-scaling run of the headline bench on N GPUs of one box (charged N x box time):
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tests/gpu_scaling_call.sh r02 8'
# Runs N = 1, 2, 4, ... up to the given count back to back, exactly as the driver launches them, and writes
# gpurun_out/<tag>_bench_n<N>.json (+ .err). Weights are generated on rank 0 and NCCL-broadcast once; no collective in the loop.
set -u
TAG=${1:-rXX}
MAXN=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_scaling_gpus.csv 2>&1
for N in 1 2 4 8; do
  [ "$N" -gt "$MAXN" ] && break
  if [ "$N" -eq 1 ]; then
    timeout -k 15 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-text-api > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
  else
    timeout -k 15 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
        bench.py --gpus $N --steps 50 --warmup 5 > $OUT/${TAG}_bench_n${N}.json 2> $OUT/${TAG}_bench_n${N}.err
  fi
  echo "N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n${N}.json").read().strip().splitlines()[-1])
    print("  N=%d  %.0f docs/s  (%.0f per GPU)  e2e %.0f  %.2f ms/step  sm %s MHz %s" % (d["n_gpus"], d["value"], d["value"] / d["n_gpus"],
          d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print("  unreadable:", e)
PY
done
