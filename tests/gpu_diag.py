"""GPU bring-up diagnostics for the sm_100a kernels (run on the B200 box through gpurun; not a pytest).

Each case runs in its own subprocess under a timeout so that a hung mbarrier pipeline or a sticky CUDA
error cannot take the other cases down. Results go to stdout and gpurun_out/diag.json; arrays of the first
failing GEMM case are saved to gpurun_out/ for offline analysis.

    python tests/gpu_diag.py            # all cases
    python tests/gpu_diag.py --case gemm:128:256:64:0:256
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llm-rankers_b200"))
OUT = os.path.join(ROOT, "gpurun_out")


def gelu_new(x):
    return 0.5 * x * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * x ** 3)))


def gemm_reference(a, w, epi, resid):
    import b200rank as br
    a64 = br.bf16_bits_to_f32(br.f32_to_bf16_bits(a)).astype(np.float64)
    w64 = br.bf16_bits_to_f32(br.f32_to_bf16_bits(w)).astype(np.float64)
    acc = a64 @ w64.T
    if epi == br.EPI_GATED_BF16:
        n = acc.shape[1]
        t = acc.reshape(acc.shape[0], n // 256, 2, 128)
        return (gelu_new(t[:, :, 0, :]) * t[:, :, 1, :]).reshape(acc.shape[0], n // 2)
    if epi == br.EPI_RESID_F32:
        return acc + resid.astype(np.float64)
    return acc


def run_gemm_case(M, N, K, epi, bn, simt, pattern="rand"):
    import b200rank as br
    rng = np.random.default_rng(1234 + M + 7 * N + 13 * K + epi)
    if pattern == "ident":  # W = first N rows of I_K: out[:, n] = A[:, n]
        a = rng.standard_normal((M, K)).astype(np.float32)
        w = np.eye(N, K, dtype=np.float32)
    else:
        a = rng.standard_normal((M, K)).astype(np.float32)
        w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    resid = rng.standard_normal((M, N)).astype(np.float32) if epi == br.EPI_RESID_F32 else None
    out, ms = br.test_gemm(a, w, epi=epi, block_n=bn, use_simt=bool(simt), resid=resid)
    out2, ms2 = br.test_gemm(a, w, epi=epi, block_n=bn, use_simt=bool(simt), resid=resid)
    ref = gemm_reference(a, w, epi, resid)
    err = np.abs(out.astype(np.float64) - ref)
    tol = 0.02 * (np.abs(ref).max() + 1e-6) if epi in (br.EPI_BF16, br.EPI_GATED_BF16) else 2e-3 * (np.abs(ref).max() + 1e-6)
    ok = bool(err.max() <= tol) and bool(np.array_equal(out, out2))
    res = dict(kind="gemm", M=M, N=N, K=K, epi=epi, bn=bn, simt=simt, pattern=pattern, max_err=float(err.max()),
               mean_err=float(err.mean()), tol=float(tol), ok=ok, ms=ms2, deterministic=bool(np.array_equal(out, out2)),
               tflops=2.0 * M * N * K / (ms2 * 1e-3) / 1e12 if ms2 > 0 else 0.0)
    if not ok:
        bad = np.argwhere(err > tol)
        res["n_bad"] = int(bad.shape[0])
        res["first_bad"] = bad[:8].tolist()
        res["bad_rows"] = np.unique(bad[:, 0])[:16].tolist()
        res["bad_cols"] = np.unique(bad[:, 1])[:16].tolist()
        tag = f"gemm_{M}_{N}_{K}_{epi}_{bn}_{pattern}"
        os.makedirs(OUT, exist_ok=True)
        if M * N <= 1 << 20:
            np.save(os.path.join(OUT, tag + "_out.npy"), out)
            np.save(os.path.join(OUT, tag + "_ref.npy"), ref.astype(np.float32))
    return res


def attention_reference(qkv, cu, H, bias):
    import b200rank as br
    qkv = br.bf16_bits_to_f32(br.f32_to_bf16_bits(qkv)).astype(np.float64)
    inner = H * 64
    out = np.zeros((qkv.shape[0], inner))
    for d in range(len(cu) - 1):
        s, e = cu[d], cu[d + 1]
        L = e - s
        idx = np.arange(L)
        rel = np.clip(idx[None, :] - idx[:, None], -128, 128) + 128
        for h in range(H):
            q = qkv[s:e, h * 64:(h + 1) * 64]
            k = qkv[s:e, inner + h * 64: inner + (h + 1) * 64]
            v = qkv[s:e, 2 * inner + h * 64: 2 * inner + (h + 1) * 64]
            sc = q @ k.T + bias[h][rel]
            sc -= sc.max(axis=1, keepdims=True)
            p = np.exp(sc)
            p /= p.sum(axis=1, keepdims=True)
            out[s:e, h * 64:(h + 1) * 64] = p @ v
    return out


def run_attn_case(lens, H, mode=0):
    import b200rank as br
    rng = np.random.default_rng(99 + sum(lens) + H)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    tokens = int(cu[-1])
    qkv = rng.standard_normal((tokens, 3 * H * 64)).astype(np.float32)
    qkv[:, : H * 64] *= 0.35  # keep scores O(few) like a trained model (no 1/sqrt(d) in T5)
    bias = rng.standard_normal((H, br.ATTN_BIAS_LEN)).astype(np.float32)
    out = br.test_enc_attention(qkv, cu, H, bias, mode=mode)
    ref = attention_reference(qkv, cu, H, bias)
    err = np.abs(out - ref)
    tol = 0.03 * np.abs(ref).max()
    res = dict(kind="attn", lens=list(map(int, lens))[:16], n_docs=len(lens), H=H, mode=mode, finite=bool(np.isfinite(out).all()), max_err=float(err.max()), mean_err=float(err.mean()), tol=float(tol),
               ok=bool(err.max() <= tol))
    if not res["ok"]:
        bad = np.argwhere(err > tol)
        res["n_bad"] = int(bad.shape[0])
        res["first_bad"] = bad[:8].tolist()
        res["bad_rows"] = np.unique(bad[:, 0])[:24].tolist()
        res["bad_cols"] = np.unique(bad[:, 1])[:24].tolist()
        os.makedirs(OUT, exist_ok=True)
        tag = f"attn_m{mode}_{'_'.join(map(str, lens))}_{H}"[:80]
        if out.size <= 1 << 20:
            np.save(os.path.join(OUT, tag + "_out.npy"), out)
            np.save(os.path.join(OUT, tag + "_ref.npy"), ref.astype(np.float32))
    return res


def all_cases():
    cases = []
    # harness self-check with the CUDA-core debug kernel
    cases += ["gemm:128:256:64:0:256:1:rand", "gemm:200:512:128:2:256:1:rand"]
    # smallest tcgen05 cases first: one tile, one k-block; identity weight exposes layout permutations
    for bn in (64, 32, 128, 256):
        cases.append(f"gemm:128:{bn}:64:3:{bn}:0:ident")
    cases += ["gemm:128:64:64:3:64:0:rand", "gemm:128:256:64:3:256:0:rand", "gemm:128:256:256:3:256:0:rand",
              "gemm:128:256:1024:3:256:0:rand", "gemm:256:512:512:3:256:0:rand"]
    # all epilogues, ragged M, multiple waves (persistent loop + both TMEM buffers)
    for epi in (0, 1, 2, 3):
        cases.append(f"gemm:1000:1024:1024:{epi}:256:0:rand")
    for bn in (32, 64, 128):
        cases.append(f"gemm:100:1024:1024:0:{bn}:0:rand")
        cases.append(f"gemm:100:1024:1024:1:{bn}:0:rand")
    cases += ["gemm:5888:3072:1024:0:256:0:rand", "gemm:5888:1024:1024:1:256:0:rand", "gemm:5888:5632:1024:2:256:0:rand",
              "gemm:5888:1024:2816:1:256:0:rand", "gemm:18400:3072:1024:0:256:0:rand", "gemm:18400:5632:1024:2:256:0:rand",
              "gemm:18400:1024:2816:1:256:0:rand", "gemm:18400:1024:1024:1:256:0:rand", "gemm:100:32128:1024:3:256:0:rand",
              "gemm:1000:1152:512:0:0:0:rand"]
    cases += ["attn:64:1", "attn:184:2", "attn:7,64,65,128,129,184,200:3", "attn:1536:1", "attn:184,184,184,184:16"]
    # tcgen05 attention (mode 3): one tile, two tiles, ragged, 4 key blocks, many heads
    cases += ["attn:64:1:3", "attn:128:1:3", "attn:184:2:3", "attn:7,64,65,128,129,184,192:3:3", "attn:250,256,130,193:2:3",
              "attn:184,184,184,184:16:3", "attn:184:2:2"]
    # scores-in-registers mma.sync attention (mode 4)
    cases += ["attn:64:1:4", "attn:128:1:4", "attn:184:2:4", "attn:7,64,65,128,129,184,192:3:4", "attn:250,256,130,193:2:4",
              "attn:184,184,184,184:16:4", "attn:1,2,3,8,9,15,16,17:2:4"]
    # persistent tcgen05 attention (mode 5): single item, two tiles, ragged, more items than SMs (the pipelined phases)
    cases += ["attn:64:1:5", "attn:128:1:5", "attn:184:2:5", "attn:7,64,65,128,129,184,192:3:5", "attn:1,2,3,8,9,15,16,17,33,100,150:2:5",
              "attn:184,184,184,184:16:5", "attn:rand300:16:5", "attn:rand37:5:5", "attn:rand20:32:5", "attn:rand40:12:5"]
    # column-split softmax variant of the persistent kernel (mode 6)
    cases += ["attn:64:1:6", "attn:128:1:6", "attn:184:2:6", "attn:7,64,65,128,129,184,192:3:6", "attn:1,2,3,8,9,15,16,17,33,95,96,97,100,150:2:6",
              "attn:184,184,184,184:16:6", "attn:rand300:16:6", "attn:rand37:5:6", "attn:rand20:32:6"]
    return cases


def run_one(spec):
    parts = spec.split(":")
    if parts[0] == "gemm":
        M, N, K, epi, bn, simt = map(int, parts[1:7])
        return run_gemm_case(M, N, K, epi, bn, simt, parts[7] if len(parts) > 7 else "rand")
    if parts[0] == "attn":
        if parts[1].startswith("rand"):  # randN: N ragged documents of 1..192 tokens
            lens = np.random.default_rng(int(parts[1][4:])).integers(1, 193, size=int(parts[1][4:])).tolist()
        else:
            lens = [int(x) for x in parts[1].split(",")]
        return run_attn_case(lens, int(parts[2]), int(parts[3]) if len(parts) > 3 else 0)
    raise ValueError(spec)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    ap.add_argument("--timeout", type=int, default=120)
    ap.add_argument("--only", default=None, help="substring filter on case specs, e.g. attn")
    args = ap.parse_args()
    if args.case:
        try:
            res = run_one(args.case)
        except Exception as exc:  # noqa: BLE001 - report everything to the parent
            res = dict(case=args.case, ok=False, error=repr(exc))
        print("RESULT " + json.dumps(res))
        return
    os.makedirs(OUT, exist_ok=True)
    results = []
    for spec in all_cases():
        if args.only and args.only not in spec:
            continue
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", spec], capture_output=True, text=True,
                               timeout=args.timeout)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            res = json.loads(line[-1][7:]) if line else dict(ok=False, error="no result", rc=p.returncode, stderr=p.stderr[-600:])
        except subprocess.TimeoutExpired:
            res = dict(ok=False, error=f"timeout {args.timeout}s (hang)")
        res["case"] = spec
        res["wall_s"] = round(time.time() - t0, 1)
        results.append(res)
        brief = {k: (round(v, 5) if isinstance(v, float) else v) for k, v in res.items() if k not in ("first_bad", "bad_rows", "bad_cols")}
        print(("PASS " if res.get("ok") else "FAIL ") + json.dumps(brief), flush=True)
        with open(os.path.join(OUT, "diag.json"), "w") as f:
            json.dump(results, f, indent=1)
    n_ok = sum(1 for r in results if r.get("ok"))
    print(f"SUMMARY {n_ok}/{len(results)} passed")


if __name__ == "__main__":
    main()
