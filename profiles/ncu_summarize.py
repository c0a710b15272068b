"""Turns the raw ncu outputs that gpurun brings back (gpurun_out/) into the text summaries committed next to this file.

    python profiles/ncu_summarize.py launches gpurun_out/launches_final.csv
    python profiles/ncu_summarize.py report   gpurun_out/prof_gemm_cg2.ncu-rep [metric substrings ...]

`launches` groups an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid) and prints each group's share
of the serialised time; `report` prints selected metrics of every launch in a `--set full` capture side by side.
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

DEFAULT_METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, g, b, v = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size"), hdr.index("Metric Value")
    groups = OrderedDict()
    for r in rows[1:]:
        name = r[k].split("(")[0].replace("void ", "")
        groups.setdefault((name, r[g], r[b]), []).append(float(r[v].replace(",", "")) / 1e3)
    total = sum(sum(x) for x in groups.values())
    print(f"{len(rows) - 1} launches, total {total:.1f} us (serialised, cold cache, unloaded clocks: compare SHARES, not absolutes)")
    for (name, grid, block), x in sorted(groups.items(), key=lambda kv: -sum(kv[1])):
        print(f"{sum(x):9.1f} us {100 * sum(x) / total:5.1f}%  n={len(x):4d} avg={sum(x) / len(x):7.1f} us  {name} grid={grid} block={block}")


def report(path, wanted):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    print("kernels: " + " | ".join(f"{r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '')} grid={r[hdr.index('Grid Size')]}" for r in body))
    for m in wanted:
        for i, h in enumerate(hdr):
            if h == m or (m not in hdr and m in h):
                print(f"{h:78s} [{units[i]:14s}] " + " | ".join(f"{r[i]:>12s}" for r in body))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[2], sys.argv[3:] or DEFAULT_METRICS)
