#!/bin/bash
# GPU call: `ncu --set full` of the BN = 256 encoder GEMM instantiations at the bench's shapes (8 consecutive launches = two encoder layers:
# QKV, O-projection, FFN-in, FFN-out) -> gpurun_out/<tag>_gemm.ncu-rep; profiles/ncu_traffic.py turns it into profiles/r02_ncu_traffic.json
# (`roofline.traffic` of the bench line). Then one full headline bench line.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tests/gpu_call_ncu_gemm.sh r02n'
set -u
TAG=${1:-r02n}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 15 420 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm_tcgen05_kernel<\(int\)256' -s 40 -c 8 -f -o $OUT/${TAG}_gemm \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-text-api --no-hf-cuda --no-sustained > $OUT/${TAG}_ncu_gemm.log 2>&1; echo "ncu rc=$?"
ls -la $OUT/${TAG}_gemm.ncu-rep
[ "${2:-}" = "ncu-only" ] && exit 0
timeout -k 15 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-hf-cuda > $OUT/${TAG}_bench_steps20.json 2> $OUT/${TAG}_bench_steps20.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_steps20.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("docs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2), "step_frac", round(r["step_frac"], 3), r["step_peak_source"][:40], "dom", r["kernel"], round(r["frac"], 3),
          "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "sustained", round(d["sustained"]["value"]), round(d["sustained"]["step_frac"], 3), "api_text", d.get("api_text"))
    print("parity", {k: d["parity"][k] for k in ("max_abs_logit_diff", "top10_identical", "inversions_beyond_tolerance", "kendall_tau")} if d.get("parity") else None)
except Exception as e:
    print("bench line unreadable:", e); print(open("$OUT/${TAG}_bench_steps20.err").read()[-2000:])
PY
