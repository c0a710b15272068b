"""The reference's own CPU path for pointwise yes_no, timed as a baseline: `transformers` T5ForConditionalGeneration in fp32 on
the host cores, called the way llmrankers/pointwise.py:117-124 calls it.

TEST / BASELINE INFRASTRUCTURE ONLY (like everything under oracle/): imported by bench.py's `cpu_baseline` leg and
`--impl reference` arm, and by tests. The product path never touches it.

Why this exists next to t5_oracle.py: the reference is pure Python and every FLOP of its hot path is executed by the un-vendored
dependency `transformers` (reference setup.py:18-20, README pins 4.31.0; this image ships 5.5.0) on torch's CPU kernels
(`pointwise.py:22-23`: fp32 when device == 'cpu'). `/root/reference` does not travel to the GPU box, but `transformers` and
`torch` are part of the image, so the arithmetic the reference would run there can be timed there: the model object the
reference's constructor would hold in `self.llm` (same architecture, the synthetic weights of b200rank.synthetic instead of a hub
checkpoint), the same three keyword arguments, the same 2-logit softmax. The numpy oracle stays the parity checker; this module
is the honest speed baseline (torch's threaded oneDNN/MKL GEMMs are several times faster than the numpy restatement).

Model construction follows tests/golden/make_golden.py::build_hf_model (which pins the same object against the reference's
own rerank()): transformers 5.x ties lm_head to the embedding for a freshly built model, real Flan-T5 checkpoints do not.
"""
from __future__ import annotations

import os
import time
from typing import Dict, Optional, Tuple

import numpy as np


def build_model(cfg: Dict, weights: Dict[str, np.ndarray], threads: Optional[int] = None):
    """fp32 T5ForConditionalGeneration on CPU holding `weights` (HF tensor names). Built on the meta device so no time is spent on
    a random init that load_state_dict would overwrite (~40 s for flan-t5-large)."""
    import torch
    from transformers import T5Config, T5ForConditionalGeneration

    torch.set_num_threads(threads or os.cpu_count() or 1)
    gated = bool(cfg.get("gated_gelu", True))
    tied = "lm_head.weight" not in weights
    hf_cfg = T5Config(vocab_size=cfg["vocab_size"], d_model=cfg["d_model"], d_kv=cfg.get("d_kv", 64), d_ff=cfg["d_ff"], num_layers=cfg["num_layers"],
                      num_decoder_layers=cfg["num_decoder_layers"], num_heads=cfg["num_heads"],
                      feed_forward_proj="gated-gelu" if gated else "relu", tie_word_embeddings=tied, decoder_start_token_id=0)
    with torch.device("meta"):
        model = T5ForConditionalGeneration(hf_cfg)
    sd = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)) for k, v in weights.items()}
    shared = torch.nn.Parameter(sd["shared.weight"], requires_grad=False)
    model.shared.weight = shared
    model.encoder.embed_tokens.weight = shared
    model.decoder.embed_tokens.weight = shared
    model.lm_head.weight = shared if tied else torch.nn.Parameter(sd["lm_head.weight"], requires_grad=False)
    rest = {k: v for k, v in sd.items() if k not in ("shared.weight", "lm_head.weight")}
    missing, unexpected = model.load_state_dict(rest, strict=False, assign=True)
    assert not unexpected, unexpected
    assert all(("embed_tokens" in m) or m in ("shared.weight", "lm_head.weight") for m in missing), missing
    assert not any(p.is_meta for p in model.parameters()), "a parameter was left on the meta device"
    model.config.tie_word_embeddings = tied
    assert bool(model.config.scale_decoder_outputs) == tied
    return model.eval()


def score_yes_no(model, ids: np.ndarray, mask: np.ndarray, yes_id: int, no_id: int, batch_size: int = 32) -> Tuple[np.ndarray, np.ndarray]:
    """The body of the reference's yes_no loop (pointwise.py:100-124) on token ids: batches of `batch_size` rows, decoder input =
    one pad token per row, logits of (yes, no) at decoder position 0, softmax over the two. Returns ([n,2] logits, [n] P(yes))."""
    import torch

    ids_t = torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64))
    mask_t = torch.from_numpy(np.ascontiguousarray(mask, dtype=np.int64))
    dec = torch.zeros((batch_size, 1), dtype=torch.long)   # tokenizer.pad_token_id == config.decoder_start_token_id == 0
    out_logits, out_scores = [], []
    with torch.no_grad():
        for b0 in range(0, ids_t.shape[0], batch_size):
            rows = slice(b0, min(b0 + batch_size, ids_t.shape[0]))
            n = rows.stop - rows.start
            logits = model(input_ids=ids_t[rows], attention_mask=mask_t[rows], decoder_input_ids=dec[:n]).logits
            two = torch.cat((logits[:, :, yes_id], logits[:, :, no_id]), dim=1)
            out_logits.append(two.numpy().copy())
            out_scores.append(torch.nn.functional.softmax(two, dim=1)[:, 0].numpy().copy())
    return np.concatenate(out_logits, 0), np.concatenate(out_scores, 0)


def timed_docs_per_s(model, ids: np.ndarray, mask: np.ndarray, yes_id: int, no_id: int, batch_size: int, repeats: int = 1):
    """(docs/s, best seconds, logits) of score_yes_no over `ids`."""
    best, logits = None, None
    for _ in range(max(1, repeats)):
        t0 = time.perf_counter()
        logits, _ = score_yes_no(model, ids, mask, yes_id, no_id, batch_size)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return ids.shape[0] / best, best, logits


def timed_docs_per_s_on_device(model, ids: np.ndarray, mask: np.ndarray, yes_id: int, no_id: int, batch_size: int, device: str,
                               dtype: str = "bfloat16", repeats: int = 3, warmup: int = 2):
    """Optional second comparator: the SAME transformers model the reference would hold, moved to `device` in a reduced
    precision and run by torch's stock kernels (cuBLAS GEMMs + eager attention on CUDA) — the reference's own GPU path
    (pointwise.py:20-24 loads fp16 on 'cuda'; bf16 is used here because a random-init model overflows fp16 without the fp32 `wo`
    carve-out that from_pretrained applies). Inputs are moved once; a pass = the reference's loop over batches of `batch_size` with one
    device->host copy of the scores per batch (its per-document .item() calls would only be slower). Returns (docs/s, best seconds,
    [n,2] logits as float32). Library code on the hot path — a baseline to beat, never part of the product."""
    import torch

    dt = getattr(torch, dtype)
    m = model.to(device=device, dtype=dt)
    ids_t = torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64)).to(device)
    mask_t = torch.from_numpy(np.ascontiguousarray(mask, dtype=np.int64)).to(device)
    dec = torch.zeros((batch_size, 1), dtype=torch.long, device=device)
    is_cuda = str(device).startswith("cuda")

    def one_pass():
        outs = []
        with torch.no_grad():
            for b0 in range(0, ids_t.shape[0], batch_size):
                rows = slice(b0, min(b0 + batch_size, ids_t.shape[0]))
                logits = m(input_ids=ids_t[rows], attention_mask=mask_t[rows], decoder_input_ids=dec[:rows.stop - rows.start]).logits
                two = torch.cat((logits[:, :, yes_id], logits[:, :, no_id]), dim=1)
                outs.append(two.float().cpu())          # the D2H copy of the batch's result (synchronises the batch)
        return torch.cat(outs, 0).numpy()

    for _ in range(warmup):
        one_pass()
    best, logits = None, None
    for _ in range(max(1, repeats)):
        if is_cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        logits = one_pass()
        if is_cuda:
            torch.cuda.synchronize()
        dtm = time.perf_counter() - t0
        best = dtm if best is None else min(best, dtm)
    return ids.shape[0] / best, best, logits
