#!/bin/bash
# One gpurun call that re-establishes the measured state of the repo on a fresh B200 (first call of a round):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tests/gpu_first_call.sh r02 core'          (steps 1-6, ~15 GPU-minutes)
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tests/gpu_first_call.sh r02 experimental'  (steps 7, 6b-6e, ~15 GPU-minutes)
# (no second argument: everything in one call, ~30 GPU-minutes: give gpurun --timeout 2400)
# Writes everything under gpurun_out/<tag>_*; copy what should be judged into profiles/.
#   1. the GPU parity suite                      -> <tag>_pytest_gpu.log, parity_report.json
#   2. smoke()                                   -> <tag>_smoke.log
#   3. (last in the core part, bounded) compute-sanitizer memcheck + racecheck of smoke()
#   4. bench.py (headline, N=1)                  -> <tag>_bench_n1.json   (never under a profiler)
#   5. reference arm                             -> <tag>_bench_reference.json
#   6. ncu launch list of a 2-step bench run     -> <tag>_ncu_launches.csv (per-launch times are cold-cache: compare shares)
#   7. experiments/epi_probe.cu                  -> <tag>_epi_probe.txt
set -u
TAG=${1:-rXX}
PART=${2:-all}
OUT=gpurun_out
mkdir -p $OUT
core_part() {
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/${TAG}_gpu.csv 2>&1
timeout -k 15 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -3 $OUT/${TAG}_pytest_gpu.log
timeout -k 15 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log
timeout -k 15 900 python bench.py --steps 100 --warmup 5 --hf-cuda > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"
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
timeout -k 15 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-text-api > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu rc=$?"
# the sanitizer passes come LAST and are bounded to 5 minutes each: they are the slowest and least predictable step, and nothing above
# may be lost to them when the call's own timeout strikes
for tool in memcheck racecheck; do
  timeout -k 15 300 compute-sanitizer --tool $tool --error-exitcode 3 python __graft_entry__.py smoke > $OUT/${TAG}_sanitizer_${tool}.log 2>&1
  echo "sanitizer $tool rc=$?" | tee -a $OUT/${TAG}_sanitizer_${tool}.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/${TAG}_sanitizer_${tool}.log | tail -2
done
}
experimental_part() {
# 7. stand-alone design probes (experiments/README.md)
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -o /tmp/epi_probe experiments/epi_probe.cu > $OUT/${TAG}_epi_probe.txt 2>&1 \
  && timeout -k 15 120 /tmp/epi_probe >> $OUT/${TAG}_epi_probe.txt 2>&1; echo "epi_probe rc=$?"; tail -9 $OUT/${TAG}_epi_probe.txt
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -o /tmp/tmem_probe experiments/tmem_probe.cu > $OUT/${TAG}_tmem_probe.txt 2>&1 \
  && timeout -k 15 120 /tmp/tmem_probe >> $OUT/${TAG}_tmem_probe.txt 2>&1; echo "tmem_probe rc=$?"; tail -16 $OUT/${TAG}_tmem_probe.txt
# 6b. experimental kernel variants (compiled in round 1, not yet run): agreement test + A/B of the headline bench
B200RANK_TEST_EXPERIMENTAL=1 timeout -k 15 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -k variants_agree > $OUT/${TAG}_pytest_experimental.log 2>&1; echo "experimental variants rc=$?"
tail -3 $OUT/${TAG}_pytest_experimental.log
B200RANK_EPI_PIPE=1 timeout -k 15 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-text-api > $OUT/${TAG}_bench_n1_epi_pipe.json 2> $OUT/${TAG}_bench_n1_epi_pipe.err; echo "bench epi_pipe rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n1_epi_pipe.json").read().strip().splitlines()[-1])
    print("EPI_PIPE docs/s", round(d["value"]), {k: v for k, v in d["roofline"]["by_kernel_ms_per_step"].items() if "epi1" in k})
except Exception as e:
    print("epi_pipe bench line unreadable:", e)
PY
B200RANK_EPI_PIPE=3 timeout -k 15 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-text-api > $OUT/${TAG}_bench_n1_epi_pipe3.json 2> $OUT/${TAG}_bench_n1_epi_pipe3.err; echo "bench epi_pipe=3 (fp32 residual + bf16 epilogues pipelined) rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n1_epi_pipe3.json").read().strip().splitlines()[-1])
    print("EPI_PIPE=3 docs/s", round(d["value"]), {k: v for k, v in d["roofline"]["by_kernel_ms_per_step"].items() if "epi0" in k or "epi1" in k})
except Exception as e:
    print("epi_pipe=3 bench line unreadable:", e)
PY
B200RANK_EPI_PIPE=1 B200RANK_EPI_HINT=last timeout -k 15 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-text-api > $OUT/${TAG}_bench_n1_epi_pipe_evict_last.json 2> $OUT/${TAG}_bench_n1_epi_pipe_evict_last.err; echo "bench epi_pipe+evict_last rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n1_epi_pipe_evict_last.json").read().strip().splitlines()[-1])
    k = d["roofline"]["by_kernel_ms_per_step"]
    print("EPI_PIPE+EVICT_LAST docs/s", round(d["value"]), {x: v for x, v in k.items() if "epi1" in x or x == "rmsnorm"})
except Exception as e:
    print("epi_pipe+evict_last bench line unreadable:", e)
PY
# 6c. two encoder streams (B200RANK_PIPE_DUAL=1, experimental): bit-identity of the pipelined path, then the A/B
B200RANK_PIPE_DUAL=1 timeout -k 15 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "pipelined_submit or large_yes_no" > $OUT/${TAG}_pytest_pipe_dual.log 2>&1; echo "pipe_dual tests rc=$?"
tail -3 $OUT/${TAG}_pytest_pipe_dual.log
B200RANK_PIPE_DUAL=1 timeout -k 15 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-text-api > $OUT/${TAG}_bench_n1_pipe_dual.json 2> $OUT/${TAG}_bench_n1_pipe_dual.err; echo "bench pipe_dual rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n1_pipe_dual.json").read().strip().splitlines()[-1])
    print("PIPE_DUAL docs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "clocks", d["clocks"]["sm_mhz"])
except Exception as e:
    print("pipe_dual bench line unreadable:", e)
PY
# 6d. one-pass softmax attention (B200RANK_ATTN=tc4, experimental): kernel against numpy incl. the exact-maximum redo, then the A/B
B200RANK_TEST_EXPERIMENTAL=1 timeout -k 15 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "onepass" > $OUT/${TAG}_pytest_tc4.log 2>&1; echo "tc4 tests rc=$?"
tail -3 $OUT/${TAG}_pytest_tc4.log
B200RANK_ATTN=tc4 timeout -k 15 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-text-api > $OUT/${TAG}_bench_n1_tc4.json 2> $OUT/${TAG}_bench_n1_tc4.err; echo "bench tc4 rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n1_tc4.json").read().strip().splitlines()[-1])
    k = d["roofline"]["by_kernel_ms_per_step"]
    print("ATTN=tc4 docs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {x: v for x, v in k.items() if "attention" in x}, "clocks", d["clocks"]["sm_mhz"])
except Exception as e:
    print("tc4 bench line unreadable:", e)
PY
# 6f. CUDA graph of the pipelined decoder chain (B200RANK_DEC_GRAPH=1, experimental): bit-identity, then the A/B
B200RANK_TEST_EXPERIMENTAL=1 timeout -k 15 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "decoder_graph" > $OUT/${TAG}_pytest_dec_graph.log 2>&1; echo "dec_graph tests rc=$?"
tail -3 $OUT/${TAG}_pytest_dec_graph.log
B200RANK_DEC_GRAPH=1 timeout -k 15 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-text-api > $OUT/${TAG}_bench_n1_dec_graph.json 2> $OUT/${TAG}_bench_n1_dec_graph.err; echo "bench dec_graph rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n1_dec_graph.json").read().strip().splitlines()[-1])
    print("DEC_GRAPH docs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "clocks", d["clocks"]["sm_mhz"])
except Exception as e:
    print("dec_graph bench line unreadable:", e)
PY
# 6d'. everything together: pipelined epilogues + two encoder streams + one-pass attention (only meaningful if each passed above)
B200RANK_EPI_PIPE=3 B200RANK_PIPE_DUAL=1 B200RANK_ATTN=tc4 B200RANK_DEC_GRAPH=1 timeout -k 15 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-text-api > $OUT/${TAG}_bench_n1_all_experimental.json 2> $OUT/${TAG}_bench_n1_all_experimental.err; echo "bench all experimental rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n1_all_experimental.json").read().strip().splitlines()[-1])
    print("ALL EXPERIMENTAL docs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "step_frac", round(d["roofline"]["step_frac"], 3), "clocks", d["clocks"]["sm_mhz"],
          "parity", (d.get("parity") or {}).get("within_logit_tolerance"))
except Exception as e:
    print("all-experimental bench line unreadable:", e)
PY
# 6e. d_kv = 128 (monot5-3b / duot5-3b head shape) on the generic-width attention (experimental): every entry point against the oracle
B200RANK_TEST_EXPERIMENTAL=1 timeout -k 15 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "wide_heads" > $OUT/${TAG}_pytest_dkv128.log 2>&1; echo "d_kv 128 tests rc=$?"
tail -3 $OUT/${TAG}_pytest_dkv128.log
# 8. source-level ncu captures of the attention kernel, shipped (tc2) and one-pass (tc4): one launch each, stall reasons per line
#    (read here with: ncu -i gpurun_out/<tag>_attn_tc2.ncu-rep --page source --csv; summarise with profiles/ncu_summarize.py)
for variant in tc2 tc4; do
  B200RANK_ATTN=$variant timeout -k 15 300 ncu --set full --clock-control none --import-source on -k regex:enc_attention_tc2_kernel -s 30 -c 1 -f \
      -o $OUT/${TAG}_attn_${variant} python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-text-api > $OUT/${TAG}_ncu_attn_${variant}.log 2>&1
  echo "ncu attention $variant rc=$?"
done
}
case "$PART" in
  core) core_part ;;
  experimental) experimental_part ;;
  *) core_part; experimental_part ;;
esac
ls -la $OUT | tail -20
