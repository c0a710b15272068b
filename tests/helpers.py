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
