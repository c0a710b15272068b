"""CPU tests of the real-checkpoint path: `llmrankers._backend._load_checkpoint` reads what `transformers` itself writes
with `save_pretrained` — single-file and sharded (index json) safetensors, sharded `pytorch_model.bin`, bf16 storage, untied (Flan-T5) and
tied (T5 v1.0) heads — and yields the config + the HF-named fp32 tensors the engine's load_state_dict takes."""
import json
import os

import numpy as np
import pytest

from helpers import model_and_weights, v10_model_and_weights


def _hf_model(cfg, w):
    from oracle import hf_cpu
    return hf_cpu.build_model(cfg, w, threads=2)


def _check(path, cfg, w, atol=0.0):
    from llmrankers._backend import _load_checkpoint
    got_cfg, tensors = _load_checkpoint(path)
    got = dict(tensors)
    for k in ("vocab_size", "d_model", "num_heads", "d_ff", "num_layers", "num_decoder_layers"):
        assert got_cfg[k] == cfg[k], k
    assert got_cfg["gated_gelu"] == bool(cfg.get("gated_gelu", True))
    tied = "lm_head.weight" not in w
    assert got_cfg["scale_decoder_outputs"] == tied
    for name, ref in w.items():
        assert name in got, name
        assert got[name].dtype == np.float32
        np.testing.assert_allclose(got[name], ref, rtol=0, atol=atol, err_msg=name)
    extra = set(got) - set(w) - {"encoder.embed_tokens.weight", "decoder.embed_tokens.weight", "lm_head.weight"}
    assert not extra, extra
    return got_cfg, got


@pytest.mark.parametrize("fmt", ["safetensors", "safetensors_sharded", "bin_sharded"])
def test_flan_checkpoint_formats(fmt, tmp_path):
    cfg, w = model_and_weights("tiny")
    model = _hf_model(cfg, w)
    kw = dict(safe_serialization=fmt.startswith("safetensors"))
    if fmt.endswith("sharded"):
        kw["max_shard_size"] = "300KB"
    model.save_pretrained(str(tmp_path), **kw)
    files = sorted(os.listdir(tmp_path))
    if fmt.endswith("sharded"):
        assert any(f.endswith(".index.json") for f in files) and sum(f.endswith((".safetensors", ".bin")) for f in files) > 1, files
    got_cfg, got = _check(str(tmp_path), cfg, w)
    # the loaded tensors drive the oracle to the same answers as the model that was saved
    from oracle import hf_cpu
    from oracle.t5_oracle import T5Oracle
    rng = np.random.default_rng(0)
    ids = rng.integers(3, cfg["vocab_size"] - 128, size=(3, 17))
    mask = np.ones_like(ids)
    a, _ = T5Oracle(cfg, {k: v for k, v in got.items() if "embed_tokens" not in k}).score_yes_no(ids, mask, 12, 13)
    b, _ = hf_cpu.score_yes_no(model, ids, mask, 12, 13)
    np.testing.assert_allclose(a, b, atol=3e-4)


def test_bf16_checkpoint_loads_as_fp32(tmp_path):
    import torch
    cfg, w = model_and_weights("tiny")
    model = _hf_model(cfg, w).to(torch.bfloat16)
    model.save_pretrained(str(tmp_path), safe_serialization=True)
    from oracle.t5_oracle import round_bf16
    _check(str(tmp_path), cfg, {k: round_bf16(v) for k, v in w.items()})


def test_tied_v10_checkpoint_has_no_lm_head_and_scales_logits(tmp_path):
    cfg, w = v10_model_and_weights("tiny")
    assert "lm_head.weight" not in w
    model = _hf_model(cfg, w)
    model.save_pretrained(str(tmp_path), safe_serialization=True)
    got_cfg, got = _check(str(tmp_path), cfg, w)
    assert got_cfg["gated_gelu"] is False and got_cfg["scale_decoder_outputs"] is True


def test_missing_shard_and_empty_directory_fail_loudly(tmp_path):
    from llmrankers._backend import _load_checkpoint
    cfg, w = model_and_weights("tiny")
    model = _hf_model(cfg, w)
    model.save_pretrained(str(tmp_path), safe_serialization=True, max_shard_size="300KB")
    shard = sorted(f for f in os.listdir(tmp_path) if f.endswith(".safetensors"))[-1]
    os.remove(tmp_path / shard)
    with pytest.raises(FileNotFoundError, match="shards"):
        _load_checkpoint(str(tmp_path))
    for f in os.listdir(tmp_path):
        if f != "config.json":
            os.remove(tmp_path / f)
    with pytest.raises(FileNotFoundError, match="no model.safetensors"):
        _load_checkpoint(str(tmp_path))
    with open(tmp_path / "config.json") as f:
        hf = json.load(f)
    hf["d_kv"] = 96
    with open(tmp_path / "config.json", "w") as f:
        json.dump(hf, f)
    with pytest.raises(NotImplementedError, match="d_kv"):
        _load_checkpoint(str(tmp_path))


def test_legacy_v10_bin_with_cross_attention_bias_key_loads(tmp_path):
    """Legacy T5 v1.0 `.bin` checkpoints (t5-*, monoT5 / duoT5) carry `decoder.block.0.layer.1.EncDecAttention.relative_attention_bias.weight`;
    transformers ignores it (`_keys_to_ignore_on_load_unexpected`). Engine.load_state_dict must skip it instead of refusing the checkpoint."""
    import torch
    import b200rank as br
    cfg, w = v10_model_and_weights("tiny")
    model = _hf_model(cfg, w)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    legacy = "decoder.block.0.layer.1.EncDecAttention.relative_attention_bias.weight"
    sd[legacy] = torch.zeros(32, cfg["num_heads"])
    model.config.save_pretrained(str(tmp_path))
    torch.save(sd, str(tmp_path / "pytorch_model.bin"))
    from llmrankers._backend import _load_checkpoint
    _, tensors = _load_checkpoint(str(tmp_path))
    names = [n for n, _ in tensors]
    assert legacy in names                       # the file reader passes everything through ...
    assert br.is_ignored_tensor(legacy)          # ... and the engine loader drops what HF drops
    assert br.is_ignored_tensor("encoder.embed_tokens.weight") and not br.is_ignored_tensor("decoder.block.0.layer.1.EncDecAttention.k.weight")

    class _Recorder(br.Engine):                  # load_state_dict without a device: record what would be uploaded
        def __init__(self):
            self.seen = []
        def load_tensor(self, name, arr):
            self.seen.append(name)
        def missing_tensors(self):
            return []
        def close(self):
            pass
        def __del__(self):
            pass
    rec = _Recorder()
    rec.load_state_dict((n, np.zeros(1, np.float32)) for n in names)
    assert legacy not in rec.seen and "shared.weight" in rec.seen
    assert rec.seen.count("lm_head.weight") == 1   # tied checkpoint: lm_head falls back to shared
