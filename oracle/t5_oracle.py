"""CPU oracle for the hot path: a numpy fp32 restatement of the Flan-T5 forward the reference rankers call.

TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference leg may import it, and only as the checker / CPU baseline.

The arithmetic of the reference's hot path lives in an un-vendored third-party dependency:
`transformers` (reference setup.py:18-20 requires >=4.31.0; this image has 5.5.0). Each function below cites
the file:line of `$TF = transformers/models/t5/modeling_t5.py` it restates, and the reference call site it serves.

Parity pinning: the reference has no tests or golden vectors. This oracle is pinned against
outputs of the reference itself run in the build container — the reference's own PointwiseLlmRanker /
SetwiseLlmRanker / PairwiseLlmRanker.rerank() driving a live transformers T5ForConditionalGeneration (fp32) —
committed as fixtures under tests/golden/ together with the generating script tests/golden/make_golden.py.
tests/test_oracle_golden.py checks logits, scores, generated ids and final orderings against those fixtures.

Like the reference's CPU path (llmrankers/pointwise.py:22-23) everything is fp32, and batches are padded +
masked exactly as HF does (additive finfo.min mask), not packed.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np

F32_MIN = np.finfo(np.float32).min


def round_bf16(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 -> fp32, round-to-nearest-even (what the engine's bf16 stores do)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16).astype(np.uint32).view(np.float32).reshape(np.shape(x))


# ------------------------------------------------------------------------------------------------ pieces
def relative_position_bucket(relative_position: np.ndarray, bidirectional: bool, num_buckets: int = 32,
                             max_distance: int = 128) -> np.ndarray:
    """$TF:189-234 `_relative_position_bucket` (float32 log like torch)."""
    rp = np.asarray(relative_position, dtype=np.int64)
    ret = np.zeros_like(rp)
    if bidirectional:
        num_buckets //= 2
        ret = ret + (rp > 0).astype(np.int64) * num_buckets
        rp = np.abs(rp)
    else:
        rp = -np.minimum(rp, 0)
    max_exact = num_buckets // 2
    is_small = rp < max_exact
    with np.errstate(divide="ignore"):
        val = (np.log(rp.astype(np.float32) / np.float32(max_exact)) / np.float32(math.log(max_distance / max_exact))
               * np.float32(num_buckets - max_exact))
    large = max_exact + np.where(np.isfinite(val), val, 0).astype(np.int64)
    large = np.minimum(large, num_buckets - 1)
    return ret + np.where(is_small, rp, large)


def compute_bias(table: np.ndarray, q_len: int, k_len: int, bidirectional: bool, num_buckets: int,
                 max_distance: int) -> np.ndarray:
    """$TF:236-251 `compute_bias` -> [1, H, q_len, k_len]; table is relative_attention_bias.weight [buckets, H]."""
    ctx = np.arange(q_len)[:, None]
    mem = np.arange(k_len)[None, :]
    buckets = relative_position_bucket(mem - ctx, bidirectional, num_buckets, max_distance)
    return table[buckets].transpose(2, 0, 1)[None].astype(np.float32)


def rms_norm(x: np.ndarray, w: np.ndarray, eps: float) -> np.ndarray:
    """$TF:55-68 T5LayerNorm: no mean subtraction, no bias, fp32 variance."""
    var = np.mean(x.astype(np.float32) ** 2, axis=-1, keepdims=True)
    return (w * (x * (1.0 / np.sqrt(var + np.float32(eps))))).astype(np.float32)


def gelu_new(x: np.ndarray) -> np.ndarray:
    """transformers/activations.py:59-66 NewGELUActivation (tanh approximation)."""
    x = x.astype(np.float32)
    return (0.5 * x * (1.0 + np.tanh(np.float32(math.sqrt(2.0 / math.pi)) * (x + np.float32(0.044715) * x ** 3)))).astype(np.float32)


def softmax(x: np.ndarray, axis: int = -1) -> np.ndarray:
    x = x.astype(np.float32)
    m = np.max(x, axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=axis, keepdims=True)


def log_softmax(x: np.ndarray, axis: int = -1) -> np.ndarray:
    x = x.astype(np.float32)
    m = np.max(x, axis=axis, keepdims=True)
    s = x - m
    return s - np.log(np.sum(np.exp(s), axis=axis, keepdims=True))


def shift_right(labels: np.ndarray, decoder_start_token_id: int = 0, pad_token_id: int = 0) -> np.ndarray:
    """$TF:595-614 `_shift_right`."""
    out = np.zeros_like(labels)
    out[..., 1:] = labels[..., :-1]
    out[..., 0] = decoder_start_token_id
    out[out == -100] = pad_token_id
    return out


# ------------------------------------------------------------------------------------------------ model
class T5Oracle:
    """cfg keys: vocab_size d_model d_kv num_heads d_ff num_layers num_decoder_layers rel_buckets rel_max_distance
    layer_norm_eps scale_decoder_outputs pad_id eos_id. weights: HF state_dict names -> fp32 arrays."""

    def __init__(self, cfg: Dict, weights: Dict[str, np.ndarray], emulate_bf16: bool = False):
        """emulate_bf16=True rounds to bf16 at exactly the points where the CUDA engine stores bf16 (GEMM weights, normed
        activations, q/k/v, attention outputs, gated activations, encoder output, cross K/V, final hidden) while keeping
        fp32 accumulation, fp32 residual stream, fp32 softmax statistics and the fp32 embedding table. It separates
        'inherent bf16 error' from 'kernel bug' in the GPU tests; the fp32 mode is the reference-pinned oracle."""
        self.cfg = dict(cfg)
        self.bf16 = bool(emulate_bf16)
        self.w = {k: np.asarray(v, dtype=np.float32) for k, v in weights.items()}
        if self.bf16:
            for k in list(self.w):
                if k.endswith((".q.weight", ".k.weight", ".v.weight", ".o.weight", ".wi_0.weight", ".wi_1.weight", ".wi.weight", ".wo.weight")) or k == "lm_head.weight":
                    self.w[k] = round_bf16(self.w[k])
        if "lm_head.weight" not in self.w:
            self.w["lm_head.weight"] = round_bf16(self.w["shared.weight"]) if self.bf16 else self.w["shared.weight"]
        self.H = cfg["num_heads"]
        self.dk = cfg.get("d_kv", 64)
        self.eps = cfg.get("layer_norm_eps", 1e-6)
        self.nb = cfg.get("rel_buckets", 32)
        self.md = cfg.get("rel_max_distance", 128)

    def _r(self, x: np.ndarray) -> np.ndarray:
        return round_bf16(x) if self.bf16 else x

    def _norm(self, x: np.ndarray, name: str) -> np.ndarray:
        return self._r(rms_norm(x, self.w[name], self.eps))

    # $TF:153-344 T5Attention.forward (eager path): no 1/sqrt(d) scaling (:308), bias+mask added to the scores,
    # softmax in fp32 (:331)
    def _attention(self, prefix: str, x_q: np.ndarray, x_kv: np.ndarray, bias_plus_mask: np.ndarray) -> np.ndarray:
        B, Tq, _ = x_q.shape
        Tk = x_kv.shape[1]
        H, dk = self.H, self.dk
        q = self._r(x_q @ self.w[prefix + ".q.weight"].T).reshape(B, Tq, H, dk).transpose(0, 2, 1, 3)
        k = self._r(x_kv @ self.w[prefix + ".k.weight"].T).reshape(B, Tk, H, dk).transpose(0, 2, 1, 3)
        v = self._r(x_kv @ self.w[prefix + ".v.weight"].T).reshape(B, Tk, H, dk).transpose(0, 2, 1, 3)
        scores = q @ k.transpose(0, 1, 3, 2)
        scores = scores + bias_plus_mask
        p = softmax(scores, axis=-1)
        o = self._r((p @ v).transpose(0, 2, 1, 3).reshape(B, Tq, H * dk))
        return o @ self.w[prefix + ".o.weight"].T

    # $TF:115-132 T5DenseGatedActDense (T5 v1.1 / Flan-T5: feed_forward_proj "gated-gelu");
    # $TF:88-103 T5DenseActDense (T5 v1.0 / monoT5 / duoT5: feed_forward_proj "relu": wo(relu(wi x)))
    def _ff(self, prefix: str, x: np.ndarray) -> np.ndarray:
        if prefix + ".wi.weight" in self.w:
            return self._r(np.maximum(x @ self.w[prefix + ".wi.weight"].T, 0.0)) @ self.w[prefix + ".wo.weight"].T
        g = gelu_new(x @ self.w[prefix + ".wi_0.weight"].T)
        l = x @ self.w[prefix + ".wi_1.weight"].T
        return self._r(g * l) @ self.w[prefix + ".wo.weight"].T

    # $TF:637-792 T5Stack.forward (encoder), :411-498 T5Block
    def encode(self, input_ids: np.ndarray, attention_mask: Optional[np.ndarray] = None) -> np.ndarray:
        ids = np.asarray(input_ids, dtype=np.int64)
        B, S = ids.shape
        mask = np.ones((B, S), np.float32) if attention_mask is None else np.asarray(attention_mask, np.float32)
        x = self.w["shared.weight"][ids]  # :682, no scaling
        ext = ((1.0 - mask) * F32_MIN)[:, None, None, :].astype(np.float32)  # :703-726 additive key-padding mask
        bias = compute_bias(self.w["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"], S, S, True,
                            self.nb, self.md) + ext  # computed in block 0, reused by all blocks (:625, :758)
        for l in range(self.cfg["num_layers"]):
            p = f"encoder.block.{l}"
            h = self._norm(x, p + ".layer.0.layer_norm.weight")
            x = x + self._attention(p + ".layer.0.SelfAttention", h, h, bias)
            x = x + self._ff(p + ".layer.1.DenseReluDense", self._norm(x, p + ".layer.1.layer_norm.weight"))
        return self._norm(x, "encoder.final_layer_norm.weight")  # :767

    # $TF:637-792 T5Stack.forward (decoder): self-attn (causal, unidirectional bias) -> cross-attn (no bias) -> FF
    def decode(self, decoder_input_ids: np.ndarray, enc_out: np.ndarray, attention_mask: Optional[np.ndarray] = None) -> np.ndarray:
        ids = np.asarray(decoder_input_ids, dtype=np.int64)
        B, T = ids.shape
        S = enc_out.shape[1]
        mask = np.ones((B, S), np.float32) if attention_mask is None else np.asarray(attention_mask, np.float32)
        x = self.w["shared.weight"][ids]
        causal = np.triu(np.full((T, T), F32_MIN, np.float32), k=1)[None, None]
        self_bias = compute_bias(self.w["decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"], T, T, False,
                                 self.nb, self.md) + causal
        cross_bias = np.zeros((1, self.H, T, S), np.float32) + ((1.0 - mask) * F32_MIN)[:, None, None, :]  # :313-315 + :720-726
        for l in range(self.cfg["num_decoder_layers"]):
            p = f"decoder.block.{l}"
            h = self._norm(x, p + ".layer.0.layer_norm.weight")
            x = x + self._attention(p + ".layer.0.SelfAttention", h, h, self_bias)
            h = self._norm(x, p + ".layer.1.layer_norm.weight")
            x = x + self._attention(p + ".layer.1.EncDecAttention", h, enc_out, cross_bias)
            x = x + self._ff(p + ".layer.2.DenseReluDense", self._norm(x, p + ".layer.2.layer_norm.weight"))
        return self._norm(x, "decoder.final_layer_norm.weight")

    # $TF:1064-1110 T5ForConditionalGeneration.forward: encoder -> decoder -> (scale if tied) -> lm_head
    def logits(self, input_ids, attention_mask, decoder_input_ids, cols: Optional[Sequence[int]] = None) -> np.ndarray:
        enc = self.encode(input_ids, attention_mask)
        h = self.decode(decoder_input_ids, enc, attention_mask)
        if self.cfg.get("scale_decoder_outputs", False):
            h = h * np.float32(self.cfg["d_model"] ** -0.5)  # :1105-1108
        W = self.w["lm_head.weight"] if cols is None else self.w["lm_head.weight"][np.asarray(cols)]
        return h @ W.T

    # ------------------------------------------------------------------ the reference's four uses of the forward
    def score_yes_no(self, input_ids, attention_mask, yes_id: int, no_id: int, pad_id: int = 0):
        """llmrankers/pointwise.py:102,117-124: decoder_input_ids=[[pad]], softmax over (yes, no) -> P(yes)."""
        B = np.asarray(input_ids).shape[0]
        lg = self.logits(input_ids, attention_mask, np.full((B, 1), pad_id, np.int64), cols=[yes_id, no_id])[:, 0, :]
        return lg, softmax(lg, axis=1)[:, 0]

    def score_qlm(self, input_ids, attention_mask, labels: Sequence[int], pad_id: int = 0) -> np.ndarray:
        """llmrankers/pointwise.py:58-60,73-79: labels repeated per row, decoder inputs = shift_right(labels),
        score = -sum_t CE(logits_t, label_t)."""
        B = np.asarray(input_ids).shape[0]
        lab = np.tile(np.asarray(labels, np.int64)[None], (B, 1))
        lg = self.logits(input_ids, attention_mask, shift_right(lab, pad_id, pad_id))
        lp = log_softmax(lg, axis=-1)
        return np.take_along_axis(lp, lab[..., None], axis=-1)[..., 0].sum(axis=1)

    def logits_at(self, input_ids, attention_mask, dec_prefix: Sequence[int], cols: Sequence[int], normalize: bool) -> np.ndarray:
        """llmrankers/setwise.py:184-186 (normalize: full-vocab softmax then gather) / pointwise.py:173-178 (raw)."""
        B = np.asarray(input_ids).shape[0]
        dec = np.tile(np.asarray(dec_prefix, np.int64)[None], (B, 1))
        if not normalize:
            return self.logits(input_ids, attention_mask, dec, cols=cols)[:, -1, :]
        lg = self.logits(input_ids, attention_mask, dec)[:, -1, :]
        return softmax(lg, axis=-1)[:, np.asarray(cols)]

    def greedy(self, input_ids, attention_mask, dec_prefix: Sequence[int], max_new: int, eos_id: int = 1, pad_id: int = 0) -> np.ndarray:
        """llmrankers/setwise.py:93-95, pairwise.py:196-200 -> transformers/generation/utils.py:2762-2804 greedy search:
        argmax of the last position, finished rows emit pad (:2797). Returns the NEW ids [B, max_new] (pad after eos).
        The encoder runs once; the decoder prefix is re-run per step (same arithmetic as the KV-cached loop)."""
        enc = self.encode(input_ids, attention_mask)
        B = enc.shape[0]
        dec = np.tile(np.asarray(dec_prefix, np.int64)[None], (B, 1))
        finished = np.zeros(B, bool)
        new = np.full((B, max_new), pad_id, np.int64)
        for s in range(max_new):
            h = self.decode(dec, enc, attention_mask)[:, -1, :]
            if self.cfg.get("scale_decoder_outputs", False):
                h = h * np.float32(self.cfg["d_model"] ** -0.5)
            tok = np.argmax(h @ self.w["lm_head.weight"].T, axis=-1)
            tok = np.where(finished, pad_id, tok)
            finished |= tok == eos_id
            new[:, s] = tok
            dec = np.concatenate([dec, tok[:, None]], axis=1)
        return new


# ------------------------------------------------------------------------------------------------ batching
def pad_batch(rows: List[Sequence[int]], pad_id: int = 0):
    """DataCollatorWithPadding(padding='longest') — llmrankers/pointwise.py:45-56: right-pad with pad_id, mask 0."""
    L = max(len(r) for r in rows)
    ids = np.full((len(rows), L), pad_id, np.int64)
    mask = np.zeros((len(rows), L), np.int64)
    for i, r in enumerate(rows):
        ids[i, : len(r)] = r
        mask[i, : len(r)] = 1
    return ids, mask
