import sys, os, io, contextlib, json, time
sys.path.insert(0, "/root/repo/llm-rankers_b200"); sys.path.insert(0, "/root/repo")
import numpy as np
import b200rank as br
from b200rank.synthetic import LABELS, model_cfg, synthetic_tokenizer, synthetic_weights
from llmrankers._backend import T5Backend
from llmrankers.rankers import SearchResult
from llmrankers.setwise import SetwiseLlmRanker
cfg = model_cfg("flan-t5-large"); tok = synthetic_tokenizer(); w = synthetic_weights(cfg, 929)
lab = [tok.convert_tokens_to_ids("▁" + c) for c in LABELS[:11]]; w["lm_head.weight"][lab] *= 30.0
c = br.make_config(cfg["d_model"], cfg["num_heads"], cfg["d_ff"], cfg["num_layers"], cfg["num_decoder_layers"], max_tokens=20480, max_docs=64, max_dec_len=8, max_logit_rows=256)
eng = br.Engine(c, 0); eng.load_state_dict(w.items()); be = T5Backend(eng, tok, cfg)
rng = np.random.default_rng(1)
words = rng.integers(0, 2000, size=(101, 128))
q = " ".join(f"w{int(x)}" for x in words[100, :32])
docs = [SearchResult(docid=str(i), score=0.0, text=" ".join(f"w{int(x)}" for x in words[i])) for i in range(100)]
os.environ["B200RANK_BATCHED_SORT"] = "0"
r = SetwiseLlmRanker(None, None, "cuda", num_child=10, k=10, scoring="generation", method="heapsort", backend=be)
with contextlib.redirect_stdout(io.StringIO()):
    r.rerank(q, list(docs)); eng.sync()
    eng.profile(True)
    t0 = time.perf_counter(); r.rerank(q, list(docs)); eng.sync(); dt = time.perf_counter() - t0
rep = eng.profile_report(); eng.profile(False)
n = r.total_compare
tot = sum(v["ms"] for v in rep.values())
print(json.dumps({"compares": n, "wall_ms_per_compare_profiled": dt / n * 1e3, "gpu_ms_per_compare": tot / n,
                  "by_kernel_ms_per_compare": {k: [round(v["ms"] / n, 4), round(v["n"] / n, 1)] for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])[:24]}}))
