"""Shared test helpers: golden fixtures, model reconstruction, tolerance utilities."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

_cache = {}


def golden_meta():
    if "meta" not in _cache:
        with open(os.path.join(GOLDEN, "golden_meta.json")) as f:
            _cache["meta"] = json.load(f)
    return _cache["meta"]


def golden_npz(name):
    if name not in _cache:
        _cache[name] = dict(np.load(os.path.join(GOLDEN, name)))
    return _cache[name]


def golden_v10_meta():
    """Fixtures of tests/golden/make_golden_v10.py: the reference's MonoT5LlmRanker / DuoT5LlmRanker on T5 v1.0 models."""
    if "meta_v10" not in _cache:
        with open(os.path.join(GOLDEN, "golden_v10_meta.json")) as f:
            _cache["meta_v10"] = json.load(f)
    return _cache["meta_v10"]


def v10_model_and_weights(which):
    """which in {'tiny', 'small'}: (cfg, weights) of the T5 v1.0 (relu, tied embeddings) golden models."""
    from b200rank.synthetic import model_cfg, synthetic_weights
    key = ("v10", which)
    if key not in _cache:
        m = golden_v10_meta()[which]
        cfg = model_cfg(m["model"], m["vocab_size"])
        _cache[key] = (cfg, synthetic_weights(cfg, m["seed"]))
    return _cache[key]


def v10_oracle_for(which):
    from oracle.t5_oracle import T5Oracle
    key = ("v10_oracle", which)
    if key not in _cache:
        _cache[key] = T5Oracle(*v10_model_and_weights(which))
    return _cache[key]


def model_and_weights(which, label_favouring=False):
    """Rebuild the (cfg, weights) a golden case was generated with: which in {'tiny', 'small'}."""
    from b200rank.synthetic import model_cfg, synthetic_weights
    key = (which, label_favouring)
    if key not in _cache:
        m = golden_meta()[which]
        cfg = model_cfg(m["model"], m["vocab_size"])
        w = synthetic_weights(cfg, m["seed"])
        if label_favouring:
            w = dict(w)
            w["lm_head.weight"] = w["lm_head.weight"].copy()
            w["lm_head.weight"][m["label_ids"]] *= m["label_boost"]
        _cache[key] = (cfg, w)
    return _cache[key]


def oracle_for(which, label_favouring=False):
    from oracle.t5_oracle import T5Oracle
    key = ("oracle", which, label_favouring)
    if key not in _cache:
        cfg, w = model_and_weights(which, label_favouring)
        _cache[key] = T5Oracle(cfg, w)
    return _cache[key]


def calls(npz, prefix):
    """Group 'prefix/callN/key' arrays into a list of dicts ordered by N."""
    out = {}
    for k, v in npz.items():
        if k.startswith(prefix + "/call"):
            _, c, key = k.split("/", 2)
            out.setdefault(int(c[4:]), {})[key] = v
    return [out[i] for i in sorted(out)]


def rows_from_padded(ids, mask):
    ids = np.asarray(ids)
    lengths = np.asarray(mask).sum(axis=1).astype(np.int32)
    return ids.astype(np.int32), lengths


def lengths_from_ids(ids, pad_id=0):
    """generate() without a mask: HF infers attention_mask = ids != pad (generation/utils.py:731-763)."""
    ids = np.asarray(ids)
    return (ids != pad_id).astype(np.int64)


def c_example_model():
    """The model and prompts examples/score_yes_no.c builds (same xorshift64* stream, same load order): (cfg, weights, ids, lengths)."""
    from b200rank.synthetic import model_cfg
    D, H, F, L, V, STRIDE = 128, 2, 256, 2, 2304, 24
    M64 = (1 << 64) - 1
    state = [0x9E3779B97F4A7C15]

    def tensor(rows, cols, scale, offset):
        n = rows * cols
        raw = np.empty(n, np.float64)
        s = state[0]
        for i in range(n):
            s ^= s >> 12
            s ^= (s << 25) & M64
            s ^= s >> 27
            raw[i] = ((s * 0x2545F4914F6CDD1D) & M64) >> 40
        state[0] = s
        u = (raw / 16777216.0 * 2.0 - 1.0).astype(np.float32)
        return (np.float32(offset) + u * np.float32(scale)).reshape(rows, cols) if rows > 1 else (np.float32(offset) + u * np.float32(scale))

    w = {}
    w["shared.weight"] = tensor(V, D, 1.0, 0.0)
    w["lm_head.weight"] = tensor(V, D, 0.05, 0.0)
    w["encoder.final_layer_norm.weight"] = tensor(1, D, 0.1, 1.0)
    w["decoder.final_layer_norm.weight"] = tensor(1, D, 0.1, 1.0)
    w["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"] = tensor(32, H, 0.5, 0.0)
    w["decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"] = tensor(32, H, 0.5, 0.0)
    for l in range(L):
        for m in "qkvo":
            w[f"encoder.block.{l}.layer.0.SelfAttention.{m}.weight"] = tensor(D, D, 0.06, 0.0)
            w[f"decoder.block.{l}.layer.0.SelfAttention.{m}.weight"] = tensor(D, D, 0.06, 0.0)
            w[f"decoder.block.{l}.layer.1.EncDecAttention.{m}.weight"] = tensor(D, D, 0.06, 0.0)
        for side, ffn in (("encoder", 1), ("decoder", 2)):
            for k in range(ffn + 1):
                w[f"{side}.block.{l}.layer.{k}.layer_norm.weight"] = tensor(1, D, 0.1, 1.0)
            w[f"{side}.block.{l}.layer.{ffn}.DenseReluDense.wi_0.weight"] = tensor(F, D, 0.06, 0.0)
            w[f"{side}.block.{l}.layer.{ffn}.DenseReluDense.wi_1.weight"] = tensor(F, D, 0.06, 0.0)
            w[f"{side}.block.{l}.layer.{ffn}.DenseReluDense.wo.weight"] = tensor(D, F, 0.04, 0.0)
    cfg = model_cfg("t5-tiny", V)
    lengths = np.array([24, 9, 17], np.int32)
    ids = np.zeros((3, STRIDE), np.int32)
    for l in range(3):
        for j in range(STRIDE):
            ids[l, j] = 3 + (l * 131 + j * 17) % (V - 3) if j < lengths[l] - 1 else (1 if j == lengths[l] - 1 else 0)
    return cfg, w, ids, lengths
