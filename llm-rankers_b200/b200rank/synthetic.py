"""Synthetic model + tokenizer for an offline box (no Flan-T5 checkpoints, no spiece.model, no network).

`synthetic_weights` produces seeded random weights of a given Flan-T5 architecture in HF state_dict naming,
reproducible across machines (numpy PCG64, not the torch RNG). `synthetic_tokenizer` builds an in-memory
`transformers.T5Tokenizer` (Unigram, pad=0, eos=1, unk=2) whose vocabulary covers the reference's prompt words,
`w0..wN` filler words and the labels A..W, so that the reference's prompts tokenise sensibly
(SURVEY.md Appendix A). `synthetic_prompt_ids` builds token-id rows of BASELINE's q_len/p_len shape directly.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

PROMPT_WORDS = ["Passage", "Query", "Does", "the", "passage", "answer", "query", "Answer", "Yes", "No", "or", "Please",
                "write", "a", "question", "based", "on", "this", "Given", "which", "of", "following", "passages", "is",
                "most", "relevant", "one", "to", "Output", "only", "label", "two", "more", "Relevant", "Document"]
# NB: "Relevant:", "Document:", "Document0:", "Document1:" of the monoT5 / duoT5 prompts tokenise into these pieces + characters
LABELS = [chr(ord("A") + i) for i in range(23)]  # setwise.py:22-23 CHARACTERS
N_FILLER_WORDS = 2000


def synthetic_vocab(n_words: int = N_FILLER_WORDS):
    vocab = [("<pad>", 0.0), ("</s>", 0.0), ("<unk>", 0.0), ("▁", -2.0)]
    vocab += [("▁" + w, -3.0) for w in PROMPT_WORDS]
    vocab += [("▁" + c, -3.5) for c in LABELS]
    vocab += [(f"▁w{i}", -5.0) for i in range(n_words)]
    vocab += [(c, -6.0) for c in "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789:'\"?,.!-()"]
    vocab += [(f"<extra_id_{i}>", 0.0) for i in range(99, -1, -1)]
    seen, out = set(), []
    for p, s in vocab:
        if p not in seen:
            seen.add(p)
            out.append((p, s))
    return out


def synthetic_tokenizer(n_words: int = N_FILLER_WORDS):
    from transformers import T5Tokenizer
    return T5Tokenizer(vocab=synthetic_vocab(n_words))


def synthetic_weights(cfg: Dict, seed: int, lm_head_std: float = 0.05) -> Dict[str, np.ndarray]:
    """Seeded random weights in HF state_dict naming. Scales follow T5's fan-in init
    (transformers/models/t5/modeling_t5.py:520-575) except lm_head, which is scaled down so logits are O(1) and
    softmaxes are not saturated (SURVEY.md §7 'random-init degeneracy'); q gets a x2 so attention is moderately peaked without making the random network chaotic under bf16 rounding."""
    rng = np.random.default_rng(seed)
    d, H, dk, F, V = cfg["d_model"], cfg["num_heads"], cfg.get("d_kv", 64), cfg["d_ff"], cfg["vocab_size"]
    I = H * dk
    nb = cfg.get("rel_buckets", 32)

    def n(shape, std):
        return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)

    w: Dict[str, np.ndarray] = {}
    gated = cfg.get("feed_forward_proj", "gated-gelu") == "gated-gelu"
    w["shared.weight"] = n((V, d), 1.0)
    if not cfg.get("tie_word_embeddings", False):   # tied (T5 v1.0): lm_head IS shared, logits scaled by d_model^-0.5
        w["lm_head.weight"] = n((V, d), lm_head_std)
    for stack, L in (("encoder", cfg["num_layers"]), ("decoder", cfg["num_decoder_layers"])):
        w[f"{stack}.final_layer_norm.weight"] = 1.0 + n((d,), 0.1)
        w[f"{stack}.block.0.layer.0.SelfAttention.relative_attention_bias.weight"] = n((nb, H), 0.5)
        for l in range(L):
            p = f"{stack}.block.{l}"
            atts = ["layer.0.SelfAttention"] + (["layer.1.EncDecAttention"] if stack == "decoder" else [])
            for att in atts:
                w[f"{p}.{att}.q.weight"] = n((I, d), (d * dk) ** -0.5 * 2.0)
                w[f"{p}.{att}.k.weight"] = n((I, d), d ** -0.5)
                w[f"{p}.{att}.v.weight"] = n((I, d), d ** -0.5)
                w[f"{p}.{att}.o.weight"] = n((d, I), I ** -0.5)
            ff = "layer.2" if stack == "decoder" else "layer.1"
            if gated:
                w[f"{p}.{ff}.DenseReluDense.wi_0.weight"] = n((F, d), d ** -0.5)
                w[f"{p}.{ff}.DenseReluDense.wi_1.weight"] = n((F, d), d ** -0.5)
            else:
                w[f"{p}.{ff}.DenseReluDense.wi.weight"] = n((F, d), (d / 2) ** -0.5)  # relu halves the variance
            w[f"{p}.{ff}.DenseReluDense.wo.weight"] = n((d, F), F ** -0.5)
            for j in range(3 if stack == "decoder" else 2):
                w[f"{p}.layer.{j}.layer_norm.weight"] = 1.0 + n((d,), 0.1)
    return w


def model_cfg(name: str, vocab_size: int = 32128) -> Dict:
    """Plain-dict architecture description shared by the engine config and the test oracle."""
    from . import MODEL_SHAPES
    shapes = dict(MODEL_SHAPES)
    shapes["t5-tiny"] = dict(d_model=128, num_heads=2, d_ff=256, num_layers=2, num_decoder_layers=2)
    shapes["t5v10-tiny"] = dict(d_model=128, num_heads=2, d_ff=256, num_layers=2, num_decoder_layers=2, v10=True)
    # wide heads (d_kv 128, the T5-3B head shape) at toy size: the generic-width attention path against the oracle
    shapes["t5-tiny-wide"] = dict(d_model=128, num_heads=2, d_kv=128, d_ff=256, num_layers=2, num_decoder_layers=2)
    shapes["t5v10-tiny-wide"] = dict(d_model=128, num_heads=2, d_kv=128, d_ff=256, num_layers=2, num_decoder_layers=2, v10=True)
    if name not in shapes:
        raise KeyError(f"unknown synthetic model {name}; known: {sorted(shapes)}")
    cfg = dict(shapes[name])
    v10 = cfg.pop("v10", False)   # T5 v1.0 (monoT5 / duoT5 checkpoints): relu feed-forward, tied embeddings => scaled logits
    cfg.setdefault("d_kv", 64)
    cfg.update(vocab_size=vocab_size, rel_buckets=32, rel_max_distance=128, layer_norm_eps=1e-6,
               scale_decoder_outputs=bool(v10), pad_id=0, eos_id=1)
    cfg.update(feed_forward_proj="relu" if v10 else "gated-gelu", gated_gelu=not v10, tie_word_embeddings=bool(v10))
    return cfg


# BASELINE.json / SURVEY.md §8d: a yes_no prompt row = tpl_a(3) | passage(p_len) | tpl_b(3) | query(q_len) | tpl_c(17) | </s>
TPL_A = [1782, 10, 3]
TPL_B = [3, 27569, 10]
TPL_C = [3, 4135, 8, 5454, 1525, 8, 11417, 58, 11801, 3, 31, 10070, 31, 42, 3, 31, 4168]
YES_ID, NO_ID = 2163, 465


def synthetic_prompt_ids(n_docs: int, q_len: int = 32, p_len: int = 128, seed: int = 929, ragged: bool = False,
                         vocab_hi: int = 32000):
    """Token-id rows of the headline workload (S = p_len + q_len + 24 = 184 for q32/p128); content ids uniform in
    [3, vocab_hi). ragged=True draws passage lengths in [p_len/2, p_len] to exercise padding."""
    rng = np.random.default_rng(seed)
    query = rng.integers(3, vocab_hi, size=q_len).tolist()
    rows: List[List[int]] = []
    for _ in range(n_docs):
        pl = int(rng.integers(p_len // 2, p_len + 1)) if ragged else p_len
        passage = rng.integers(3, vocab_hi, size=pl).tolist()
        rows.append(TPL_A + passage + TPL_B + query + TPL_C + [1])
    L = max(len(r) for r in rows)
    ids = np.zeros((n_docs, L), np.int32)
    lengths = np.zeros((n_docs,), np.int32)
    for i, r in enumerate(rows):
        ids[i, : len(r)] = r
        lengths[i] = len(r)
    return ids, lengths


def headline_query():
    """The HEADLINE query of bench.py and of the full-size parity test: 100 documents of the BASELINE configs[1] shape (S = 184)
    selected from a pool of random passages so that the fp32 reference's top-11 margins are further apart than bf16 arithmetic can
    move them (tests/golden/make_headline_query.py — the selection reads the fp32 reference only). Returns
    (ids [100,184] int32, lengths [100] int32, ref_logits [100,2] fp32 — the reference's own (yes, no) logits, meta dict).
    Falls back to (synthetic_prompt_ids(seed 929), None, None) when the committed fixture is absent."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    p = os.path.join(root, "tests", "golden", "headline_query.npz")
    if not os.path.exists(p):
        ids, lengths = synthetic_prompt_ids(100, 32, 128, seed=929)
        return ids, lengths, None, None
    z = np.load(p)
    with open(os.path.join(root, "tests", "golden", "headline_query_meta.json")) as f:
        meta = json.load(f)
    return z["ids"].astype(np.int32), z["lengths"].astype(np.int32), z["ref_logits"].astype(np.float32), meta
