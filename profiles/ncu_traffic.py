"""profiles/r02_ncu_traffic.json (what bench.py reports as `roofline.traffic`) from the `ncu --set full` capture of the encoder GEMMs:

    python profiles/ncu_traffic.py gpurun_out/r02n_gemm.ncu-rep 36800 > profiles/r02_ncu_traffic.json

The capture (tests/gpu_call_ncu_gemm.sh) holds consecutive launches of the BN = 256 instantiations, i.e. whole encoder layers in the order
QKV (epilogue 0), O-projection (1), gated FFN-in (2), FFN-out (1); the two residual-epilogue GEMMs are told apart by their position in
the layer. M = tokens of the bench's device pass (queries per step x 100 documents x 184). One entry per GEMM, first occurrence."""
import csv
import io
import json
import re
import subprocess
import sys


def main(path, M, d=1024, inner=1024, F=2816):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, body = rows[0], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    res = {"_source": f"ncu --set full --clock-control none ({path}, summarised in profiles/r02_ncu_gemm_summary.txt): dram__bytes_read.sum + dram__bytes_write.sum "
                      "of ONE launch, caches flushed before the launch (ncu default), so A/W reads are cold and outputs that still sit in L2 at kernel end are not counted"}
    prev_epi = None
    for r in body:
        m = re.search(r"gemm_tcgen05_kernel<(\d+), (\d+)", r[col["Kernel Name"]])
        if not m:
            continue
        bn, epi = int(m.group(1)), int(m.group(2))
        if epi == 0:
            N, K, alg = 3 * inner, d, M * d * 2 + 3 * inner * d * 2 + M * 3 * inner * 2
        elif epi == 2:
            N, K, alg = 2 * F, d, M * d * 2 + 2 * F * d * 2 + M * F * 2
        elif epi == 1 and prev_epi == 0:     # right after QKV (attention is not a GEMM): the O-projection; fp32 residual read-modify-write
            N, K, alg = d, inner, M * inner * 2 + d * inner * 2 + 2 * M * d * 4
        elif epi == 1 and prev_epi == 2:
            N, K, alg = d, F, M * F * 2 + d * F * 2 + 2 * M * d * 4
        else:
            prev_epi = epi
            continue
        prev_epi = epi
        label = f"gemm_tcgen05<bn{bn},epi{epi}> M{M} N{N} K{K}"
        if label in res:
            continue
        f = lambda name: float(r[col[name]].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        units = rows[1]
        res[label] = {"dram_read": int(f("dram__bytes_read.sum") * scale[units[col["dram__bytes_read.sum"]]]),
                      "dram_write": int(f("dram__bytes_write.sum") * scale[units[col["dram__bytes_write.sum"]]]),
                      "algorithmic_bytes": int(alg), "gpu_time_us": f("gpu__time_duration.sum"),
                      "tensor_active_pct_of_elapsed": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
