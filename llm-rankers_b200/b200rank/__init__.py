"""ctypes binding of libb200rank.so (include/b200rank.h) — the thin Python host layer of the engine.

The library is the product; this module only marshals numpy buffers across the C-ABI. There is no
CPU fallback: if the shared library is missing or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("B200RANK_LIB", os.path.join(os.path.dirname(_HERE), "libb200rank.so"))

DTYPE_F32 = 0
DTYPE_BF16 = 1
EPI_BF16, EPI_RESID_F32, EPI_GATED_BF16, EPI_F32 = 0, 1, 2, 3
ERR_ARG, ERR_CUDA, ERR_STATE, ERR_CAPACITY = -1, -2, -3, -4   # include/b200rank.h: B200RANK_ERR_*
ATTN_REL_CLAMP = 128
ATTN_BIAS_LEN = 2 * ATTN_REL_CLAMP + 1


class B200RankError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200rank error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    """Mirror of `b200rank_config` (include/b200rank.h)."""
    _fields_ = [
        ("vocab_size", C.c_int32), ("d_model", C.c_int32), ("d_kv", C.c_int32), ("num_heads", C.c_int32),
        ("d_ff", C.c_int32), ("num_layers", C.c_int32), ("num_decoder_layers", C.c_int32),
        ("rel_buckets", C.c_int32), ("rel_max_distance", C.c_int32), ("layer_norm_eps", C.c_float),
        ("gated_gelu", C.c_int32), ("scale_decoder_outputs", C.c_int32), ("pad_id", C.c_int32), ("eos_id", C.c_int32),
        ("max_tokens", C.c_int32), ("max_docs", C.c_int32), ("max_dec_len", C.c_int32), ("max_logit_rows", C.c_int32),
    ]


# Public Flan-T5 shapes; all: d_kv 64, vocab 32128, gated-gelu, 32 buckets / max distance 128.
MODEL_SHAPES: Dict[str, Dict[str, int]] = {
    "flan-t5-small": dict(d_model=512, num_heads=6, d_ff=1024, num_layers=8, num_decoder_layers=8),
    "flan-t5-base": dict(d_model=768, num_heads=12, d_ff=2048, num_layers=12, num_decoder_layers=12),
    "flan-t5-large": dict(d_model=1024, num_heads=16, d_ff=2816, num_layers=24, num_decoder_layers=24),
    "flan-t5-xl": dict(d_model=2048, num_heads=32, d_ff=5120, num_layers=24, num_decoder_layers=24),
    "flan-t5-xxl": dict(d_model=4096, num_heads=64, d_ff=10240, num_layers=24, num_decoder_layers=24),
    # T5 v1.0 shapes of the castorini monoT5 / duoT5 checkpoints (relu feed-forward, tied embeddings). The 3B ones use d_kv 128:
    # they run on the generic-width attention of csrc/attention_wide.cuh.
    "monot5-small": dict(d_model=512, num_heads=8, d_ff=2048, num_layers=6, num_decoder_layers=6, v10=True),
    "monot5-base": dict(d_model=768, num_heads=12, d_ff=3072, num_layers=12, num_decoder_layers=12, v10=True),
    "monot5-large": dict(d_model=1024, num_heads=16, d_ff=4096, num_layers=24, num_decoder_layers=24, v10=True),
    "duot5-base": dict(d_model=768, num_heads=12, d_ff=3072, num_layers=12, num_decoder_layers=12, v10=True),
    "monot5-3b": dict(d_model=1024, num_heads=32, d_kv=128, d_ff=16384, num_layers=24, num_decoder_layers=24, v10=True),
    "duot5-3b": dict(d_model=1024, num_heads=32, d_kv=128, d_ff=16384, num_layers=24, num_decoder_layers=24, v10=True),
}


def make_config(d_model: int, num_heads: int, d_ff: int, num_layers: int, num_decoder_layers: int,
                vocab_size: int = 32128, d_kv: int = 64, rel_buckets: int = 32, rel_max_distance: int = 128,
                layer_norm_eps: float = 1e-6, gated_gelu: bool = True, scale_decoder_outputs: bool = False,
                pad_id: int = 0, eos_id: int = 1, max_tokens: int = 0, max_docs: int = 0, max_dec_len: int = 0,
                max_logit_rows: int = 0) -> Config:
    return Config(vocab_size, d_model, d_kv, num_heads, d_ff, num_layers, num_decoder_layers, rel_buckets,
                  rel_max_distance, layer_norm_eps, int(gated_gelu), int(scale_decoder_outputs), pad_id, eos_id,
                  max_tokens, max_docs, max_dec_len, max_logit_rows)


_lib = None


def load_library() -> C.CDLL:
    """dlopen libb200rank.so and declare the prototypes of every symbol in include/b200rank.h."""
    global _lib
    if _lib is not None:
        return _lib
    # B200RANK_LIB=<path>: load another build of the same ABI (same-box A/B of kernel changes: tests/gpu_call_ab.sh); default in-tree
    path = os.environ.get("B200RANK_LIB") or _LIB_PATH
    if not os.path.exists(path):
        raise B200RankError(-2, f"{path} not found — run `python __graft_entry__.py build` (no CPU fallback)")
    lib = C.CDLL(path)
    i32p, f32p, vp = C.POINTER(C.c_int32), C.POINTER(C.c_float), C.c_void_p
    lib.b200rank_version.restype = C.c_char_p
    lib.b200rank_last_error.restype = C.c_char_p
    sig = {
        "b200rank_create": [C.POINTER(Config), C.c_int, C.POINTER(vp)],
        "b200rank_load_tensor": [vp, C.c_char_p, vp, C.c_int, C.c_int64, C.c_int64],
        "b200rank_missing_tensors": [vp, C.c_char_p, C.c_int],
        "b200rank_weights_blob": [vp, C.POINTER(vp), C.POINTER(C.c_size_t)],
        "b200rank_mark_weights_loaded": [vp],
        "b200rank_score_yes_no": [vp, i32p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p, f32p],
        "b200rank_score_qlm": [vp, i32p, i32p, C.c_int, C.c_int, i32p, C.c_int, f32p],
        "b200rank_logits_at": [vp, i32p, i32p, C.c_int, C.c_int, i32p, C.c_int, i32p, C.c_int, C.c_int, f32p],
        "b200rank_greedy": [vp, i32p, i32p, C.c_int, C.c_int, i32p, C.c_int, C.c_int, i32p],
        "b200rank_stage": [vp, i32p, i32p, C.c_int, C.c_int],
        "b200rank_run_yes_no_staged": [vp, C.c_int, C.c_int],
        "b200rank_fetch_yes_no": [vp, f32p, f32p],
        "b200rank_sync": [vp],
        "b200rank_submit_yes_no": [vp, i32p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)],
        "b200rank_wait_yes_no": [vp, C.c_uint64, f32p, f32p],
        "b200rank_event_record": [vp, C.c_int],
        "b200rank_event_elapsed_ms": [vp, f32p],
        "b200rank_launch_count": [vp, C.POINTER(C.c_uint64)],
        "b200rank_flush_l2": [vp],
        "b200rank_profile": [vp, C.c_int],
        "b200rank_profile_report": [vp, C.c_char_p, C.c_int],
        "b200rank_device_info": [vp, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)],
        "b200rank_test_gemm": [C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, f32p],
        "b200rank_test_enc_attention": [C.c_int, vp, i32p, C.c_int, C.c_int, f32p, vp, C.c_int],
        "b200rank_rel_bucket": [C.c_int, C.c_int, C.c_int, C.c_int],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.b200rank_destroy.argtypes = [vp]
    lib.b200rank_destroy.restype = None
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "b200rank_version", "b200rank_last_error", "b200rank_create", "b200rank_destroy", "b200rank_load_tensor",
    "b200rank_missing_tensors", "b200rank_weights_blob", "b200rank_mark_weights_loaded", "b200rank_score_yes_no",
    "b200rank_score_qlm", "b200rank_logits_at", "b200rank_greedy", "b200rank_stage", "b200rank_run_yes_no_staged",
    "b200rank_fetch_yes_no", "b200rank_sync", "b200rank_submit_yes_no", "b200rank_wait_yes_no", "b200rank_event_record", "b200rank_event_elapsed_ms",
    "b200rank_launch_count", "b200rank_flush_l2", "b200rank_profile", "b200rank_profile_report", "b200rank_device_info", "b200rank_test_gemm",
    "b200rank_test_enc_attention", "b200rank_rel_bucket",
]


def _check(rc: int) -> None:
    if rc != 0:
        raise B200RankError(rc, load_library().b200rank_last_error().decode("utf-8", "replace"))


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def f32_to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16, returned as uint16 bit patterns (same rounding as torch / __float2bfloat16_rn)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    rounded = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return rounded.astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def rel_bucket(relative_position: int, bidirectional: bool, num_buckets: int = 32, max_distance: int = 128) -> int:
    return load_library().b200rank_rel_bucket(int(relative_position), int(bidirectional), num_buckets, max_distance)


def is_ignored_tensor(name: str) -> bool:
    """Checkpoint keys that carry no weight of their own: the two aliases of `shared.weight`, and the cross-attention relative bias
    that legacy T5 v1.0 checkpoints (t5-*, the monoT5 / duoT5 `.bin` lineage) still hold although no forward reads it — transformers
    drops it silently (`_keys_to_ignore_on_load_unexpected`, modeling_t5.py), so a local checkpoint directory must load here too."""
    return (name in ("encoder.embed_tokens.weight", "decoder.embed_tokens.weight")
            or name.endswith("EncDecAttention.relative_attention_bias.weight"))


class Engine:
    """One Flan-T5 model resident on one GPU. Mirrors what `self.llm` is to the reference rankers."""

    def __init__(self, cfg: Config, device: int = 0):
        self.lib = load_library()
        self.cfg = cfg
        self.device = device
        self._h = C.c_void_p()
        self._in_flight = {}      # ticket -> documents of the pipelined batches not yet waited for (see drain())
        self._staged_n = 0
        _check(self.lib.b200rank_create(C.byref(cfg), device, C.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.b200rank_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights
    def load_tensor(self, name: str, array: np.ndarray) -> None:
        a = np.ascontiguousarray(array, dtype=np.float32)
        rows, cols = (1, a.shape[0]) if a.ndim == 1 else a.shape
        _check(self.lib.b200rank_load_tensor(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), DTYPE_F32, rows, cols))

    def load_state_dict(self, tensors: Iterable[Tuple[str, np.ndarray]]) -> None:
        """Load (name, fp32 array) pairs named like a HF T5 state_dict. A missing `lm_head.weight` (tied
        checkpoints) falls back to `shared.weight`, which is what transformers' tie_weights does."""
        seen_lm_head, shared = False, None
        for name, arr in tensors:
            if is_ignored_tensor(name):
                continue
            if name == "lm_head.weight":
                seen_lm_head = True
            if name == "shared.weight":
                shared = arr
            self.load_tensor(name, arr)
        if not seen_lm_head and shared is not None:
            self.load_tensor("lm_head.weight", shared)
        missing = self.missing_tensors()
        if missing:
            raise B200RankError(-3, f"{len(missing)} tensors missing after load: {missing[:4]}...")

    def missing_tensors(self):
        buf = C.create_string_buffer(1 << 16)
        n = self.lib.b200rank_missing_tensors(self._h, buf, len(buf))
        return [s for s in buf.value.decode().split(",") if s] if n else []

    def weights_blob(self) -> Tuple[int, int]:
        ptr, n = C.c_void_p(), C.c_size_t()
        _check(self.lib.b200rank_weights_blob(self._h, C.byref(ptr), C.byref(n)))
        return int(ptr.value), int(n.value)

    def mark_weights_loaded(self) -> None:
        _check(self.lib.b200rank_mark_weights_loaded(self._h))

    # ---- scoring (host buffers in, host buffers out)
    @staticmethod
    def _ids_lengths(ids, lengths):
        ids = _i32(ids)
        assert ids.ndim == 2
        lengths = _i32(lengths)
        assert lengths.shape == (ids.shape[0],)
        return ids, lengths

    def score_yes_no(self, ids, lengths, yes_id: int, no_id: int) -> Tuple[np.ndarray, np.ndarray]:
        ids, lengths = self._ids_lengths(ids, lengths)
        n = ids.shape[0]
        logits2 = np.empty((n, 2), np.float32)
        scores = np.empty((n,), np.float32)
        _check(self.lib.b200rank_score_yes_no(self._h, _p(ids, C.c_int32), _p(lengths, C.c_int32), n, ids.shape[1],
                                              yes_id, no_id, _p(logits2, C.c_float), _p(scores, C.c_float)))
        return logits2, scores

    def score_qlm(self, ids, lengths, labels) -> np.ndarray:
        ids, lengths = self._ids_lengths(ids, lengths)
        labels = _i32(labels).reshape(-1)
        scores = np.empty((ids.shape[0],), np.float32)
        _check(self.lib.b200rank_score_qlm(self._h, _p(ids, C.c_int32), _p(lengths, C.c_int32), ids.shape[0], ids.shape[1],
                                           _p(labels, C.c_int32), labels.shape[0], _p(scores, C.c_float)))
        return scores

    def logits_at(self, ids, lengths, dec_prefix, cols, normalize: bool) -> np.ndarray:
        ids, lengths = self._ids_lengths(ids, lengths)
        dec_prefix = _i32(dec_prefix).reshape(-1)
        cols = _i32(cols).reshape(-1)
        out = np.empty((ids.shape[0], cols.shape[0]), np.float32)
        _check(self.lib.b200rank_logits_at(self._h, _p(ids, C.c_int32), _p(lengths, C.c_int32), ids.shape[0], ids.shape[1],
                                           _p(dec_prefix, C.c_int32), dec_prefix.shape[0], _p(cols, C.c_int32),
                                           cols.shape[0], int(normalize), _p(out, C.c_float)))
        return out

    def greedy(self, ids, lengths, dec_prefix, max_new: int) -> np.ndarray:
        ids, lengths = self._ids_lengths(ids, lengths)
        dec_prefix = _i32(dec_prefix).reshape(-1)
        out = np.empty((ids.shape[0], max_new), np.int32)
        _check(self.lib.b200rank_greedy(self._h, _p(ids, C.c_int32), _p(lengths, C.c_int32), ids.shape[0], ids.shape[1],
                                        _p(dec_prefix, C.c_int32), dec_prefix.shape[0], max_new, _p(out, C.c_int32)))
        return out

    # ---- split form (bench / overlap)
    def stage(self, ids, lengths) -> None:
        ids, lengths = self._ids_lengths(ids, lengths)
        self._staged_n = ids.shape[0]
        _check(self.lib.b200rank_stage(self._h, _p(ids, C.c_int32), _p(lengths, C.c_int32), ids.shape[0], ids.shape[1]))

    def run_yes_no_staged(self, yes_id: int, no_id: int) -> None:
        _check(self.lib.b200rank_run_yes_no_staged(self._h, yes_id, no_id))

    def fetch_yes_no(self) -> Tuple[np.ndarray, np.ndarray]:
        n = self._staged_n
        logits2 = np.empty((n, 2), np.float32)
        scores = np.empty((n,), np.float32)
        _check(self.lib.b200rank_fetch_yes_no(self._h, _p(logits2, C.c_float), _p(scores, C.c_float)))
        return logits2, scores

    def submit_yes_no(self, ids, lengths, yes_id: int, no_id: int) -> Tuple[int, int]:
        """Asynchronous scoring (at most two batches in flight). Returns (ticket, n_docs) for wait_yes_no."""
        ids, lengths = self._ids_lengths(ids, lengths)
        t = C.c_uint64()
        _check(self.lib.b200rank_submit_yes_no(self._h, _p(ids, C.c_int32), _p(lengths, C.c_int32), ids.shape[0], ids.shape[1],
                                               yes_id, no_id, C.byref(t)))
        self._in_flight[int(t.value)] = ids.shape[0]
        return int(t.value), ids.shape[0]

    def submit_yes_no_staged(self, yes_id: int, no_id: int) -> Tuple[int, int]:
        """Asynchronous scoring of the batch `stage()` left in device memory (no host->device copy)."""
        t = C.c_uint64()
        _check(self.lib.b200rank_submit_yes_no(self._h, None, None, 0, 0, yes_id, no_id, C.byref(t)))
        self._in_flight[int(t.value)] = self._staged_n
        return int(t.value), self._staged_n

    def wait_yes_no(self, ticket: Tuple[int, int]) -> Tuple[np.ndarray, np.ndarray]:
        t, n = ticket
        logits2 = np.empty((n, 2), np.float32)
        scores = np.empty((n,), np.float32)
        _check(self.lib.b200rank_wait_yes_no(self._h, t, _p(logits2, C.c_float), _p(scores, C.c_float)))
        self._in_flight.pop(t, None)
        return logits2, scores

    def drain(self) -> int:
        """Wait for every pipelined batch still in flight and discard its scores (after an error between submit and wait: the
        synchronous entry points refuse to run while tickets are outstanding). Returns how many were drained."""
        n = 0
        for t, docs in sorted(self._in_flight.items()):
            self.wait_yes_no((t, docs))
            n += 1
        return n

    def sync(self) -> None:
        _check(self.lib.b200rank_sync(self._h))

    def event_record(self, which: int) -> None:
        _check(self.lib.b200rank_event_record(self._h, which))

    def event_elapsed_ms(self) -> float:
        ms = C.c_float()
        _check(self.lib.b200rank_event_elapsed_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        n = C.c_uint64()
        _check(self.lib.b200rank_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def profile(self, enable: bool) -> None:
        _check(self.lib.b200rank_profile(self._h, int(enable)))

    def profile_report(self) -> Dict[str, Dict[str, float]]:
        import json
        buf = C.create_string_buffer(1 << 16)
        _check(self.lib.b200rank_profile_report(self._h, buf, len(buf)))
        return json.loads(buf.value.decode())

    def flush_l2(self) -> None:
        _check(self.lib.b200rank_flush_l2(self._h))

    def device_info(self) -> Dict[str, int]:
        sm, wb, ws = C.c_int(), C.c_size_t(), C.c_size_t()
        _check(self.lib.b200rank_device_info(self._h, C.byref(sm), C.byref(wb), C.byref(ws)))
        return {"sm_count": sm.value, "weight_bytes": wb.value, "workspace_bytes": ws.value}


# ---- kernel-level test hooks ---------------------------------------------------------------
def test_gemm(a_f32: np.ndarray, w_f32: np.ndarray, epi: int = EPI_BF16, block_n: int = 0, use_simt: bool = False,
              resid: Optional[np.ndarray] = None, device: int = 0) -> Tuple[np.ndarray, float]:
    """Runs one GEMM through the C-ABI test hook. Inputs are rounded to bf16; returns (out fp32 view, ms)."""
    lib = load_library()
    M, K = a_f32.shape
    N, K2 = w_f32.shape
    assert K == K2
    a = f32_to_bf16_bits(a_f32)
    w = f32_to_bf16_bits(w_f32)
    n_out = N // 2 if epi == EPI_GATED_BF16 else N
    if epi in (EPI_BF16, EPI_GATED_BF16):
        out = np.zeros((M, n_out), np.uint16)
    else:
        out = np.zeros((M, n_out), np.float32) if resid is None else np.ascontiguousarray(resid, np.float32).copy()
    ms = C.c_float()
    _check(lib.b200rank_test_gemm(device, a.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), M, N, K, epi, block_n,
                                  int(use_simt), out.ctypes.data_as(C.c_void_p), C.byref(ms)))
    return (bf16_bits_to_f32(out) if out.dtype == np.uint16 else out), float(ms.value)


def test_enc_attention(qkv_f32: np.ndarray, cu_seqlens, num_heads: int, bias: np.ndarray, device: int = 0, mode: int = 0) -> np.ndarray:
    lib = load_library()
    cu = _i32(cu_seqlens)
    tokens = int(cu[-1])
    inner = num_heads * 64
    assert qkv_f32.shape == (tokens, 3 * inner)
    q = f32_to_bf16_bits(qkv_f32)
    b = np.ascontiguousarray(bias, np.float32)
    assert b.shape == (num_heads, ATTN_BIAS_LEN)
    out = np.zeros((tokens, inner), np.uint16)
    _check(lib.b200rank_test_enc_attention(device, q.ctypes.data_as(C.c_void_p), _p(cu, C.c_int32), cu.shape[0] - 1, num_heads,
                                           _p(b, C.c_float), out.ctypes.data_as(C.c_void_p), mode))
    return bf16_bits_to_f32(out)
