#!/bin/bash
# One gpurun call that re-establishes the measured state of the repo on a fresh B200 (first call of a round):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tests/gpu_first_call.sh r03'          (~12 GPU-minutes)
# (kernel experiments have their own scripts: tests/gpu_call_attn.sh, tests/gpu_call_ab.sh)
# Writes everything under gpurun_out/<tag>_*; copy what should be judged into profiles/.
#   1. the GPU parity suite                      -> <tag>_pytest_gpu.log, parity_report.json
#   2. smoke()                                   -> <tag>_smoke.log
#   3. (last in the core part, bounded) compute-sanitizer memcheck + racecheck of smoke()
#   4. bench.py (headline, N=1)                  -> <tag>_bench_n1.json   (never under a profiler)
#   5. reference arm                             -> <tag>_bench_reference.json
#   6. ncu launch list of a 2-step bench run     -> <tag>_ncu_launches.csv (per-launch times are cold-cache: compare shares)
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
core_part() {
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/${TAG}_gpu.csv 2>&1
timeout -k 15 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -3 $OUT/${TAG}_pytest_gpu.log
timeout -k 15 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log
timeout -k 15 900 python bench.py --steps 100 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("docs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "step_frac", round(r["step_frac"], 3), "dom", r["kernel"], round(r["frac"], 3),
          "traffic", r["traffic"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    print("cpu_baseline", d["cpu_baseline"]["kind"], round(d["cpu_baseline"]["value"], 2), d["cpu_baseline"]["cores"], "parity", d.get("parity"))
    print("hf_cuda", d.get("hf_cuda"))
except Exception as e:
    print("bench line unreadable:", e)
PY
timeout -k 15 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "reference rc=$?"
timeout -k 15 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/${TAG}_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-text-api > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu rc=$?"
# the sanitizer passes come LAST and are bounded to 5 minutes each: they are the slowest and least predictable step, and nothing above
# may be lost to them when the call's own timeout strikes
for tool in memcheck racecheck; do
  timeout -k 15 300 compute-sanitizer --tool $tool --error-exitcode 3 python __graft_entry__.py smoke > $OUT/${TAG}_sanitizer_${tool}.log 2>&1
  echo "sanitizer $tool rc=$?" | tee -a $OUT/${TAG}_sanitizer_${tool}.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/${TAG}_sanitizer_${tool}.log | tail -2
done
}
core_part
ls -la $OUT | tail -20
