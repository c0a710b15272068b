"""What `self.llm` + `self.tokenizer` are to the reference rankers, re-based on the B200 engine.

A T5Backend owns (a) a tokenizer (transformers.T5Tokenizer — tokenisation is not on the device path) and (b) a
b200rank.Engine holding the model on one GPU. Model sources:
  * "synthetic:<shape>[:seed=N]"  seeded random weights of a Flan-T5 shape + the in-memory synthetic tokenizer
                                   (benchmarks/tests on a box with no checkpoints),
  * a local directory with config.json + model.safetensors | pytorch_model.bin (real Flan-T5 checkpoints),
  * anything else is handed to transformers' from_pretrained (hub cache) to obtain the state dict on the CPU.
There is no CPU execution path: device must be 'cuda' / 'cuda:N' (the reference's device='cpu' raises here).
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

_ENGINE_CACHE: Dict[Tuple[str, int], "T5Backend"] = {}


def _device_index(device) -> int:
    d = str(device)
    if d == "cuda":
        return int(os.environ.get("LOCAL_RANK", "0"))
    if d.startswith("cuda:"):
        return int(d.split(":", 1)[1])
    raise RuntimeError(f"device={device!r}: the B200 engine has no CPU path (use device='cuda')")


def generate_mask_mode() -> str:
    """How generate() calls WITHOUT attention_mask treat pad tokens inside a padded batch (pairwise.py:97-99,196-200):
    'ones'  — pads are ordinary tokens and are attended: transformers >= 5 for encoder-decoder models
              (generation/utils.py:2429: mask inference is skipped when config.is_encoder_decoder); the default,
              because it is what the reference does on this image and what the golden fixtures pin;
    'infer' — attention_mask = ids != pad: transformers 4.31.0, the reference's README-tested pin."""
    m = os.environ.get("B200RANK_GENERATE_MASK", "ones")
    if m not in ("ones", "infer"):
        raise ValueError("B200RANK_GENERATE_MASK must be 'ones' or 'infer'")
    return m


class T5Backend:
    # a document's logits are bit-identical whatever else shares its device pass (tests/test_engine_gpu.py): callers may
    # merge the documents of several queries into one pass (PointwiseLlmRanker.rerank_many)
    batch_invariant = True

    def __init__(self, engine, tokenizer, cfg: Dict):
        self.engine = engine
        self.tokenizer = tokenizer
        self.cfg = cfg
        self.pad_id = cfg.get("pad_id", 0)
        self.eos_id = cfg.get("eos_id", 1)

    # ---------------------------------------------------------------- construction
    @classmethod
    def load(cls, model_name_or_path: str, tokenizer_name_or_path: Optional[str], device, cache_dir=None,
             max_tokens: int = 0, max_docs: int = 0) -> "T5Backend":
        import b200rank as br
        dev = _device_index(device)
        # tokens of one device pass: two 100-hit queries of ~184-token prompts share a pass in rerank_many (36.8 k tokens), so the public
        # constructors size the workspaces for 64 k (8 GB for flan-t5-large, 31 GB for flan-t5-xxl, of 180); B200RANK_MAX_TOKENS overrides
        if max_tokens <= 0:
            max_tokens = int(os.environ.get("B200RANK_MAX_TOKENS", "65536"))
        key = (f"{model_name_or_path}|{tokenizer_name_or_path}|{max_tokens}|{max_docs}", dev)
        if key in _ENGINE_CACHE:
            return _ENGINE_CACHE[key]
        if model_name_or_path.startswith("synthetic:"):
            from b200rank.synthetic import model_cfg, synthetic_tokenizer, synthetic_weights
            parts = model_name_or_path.split(":")
            shape = parts[1]
            opts = dict(p.split("=", 1) for p in parts[2:])
            cfg = model_cfg(shape, int(opts.get("vocab", 32128)))
            tensors: Iterable = synthetic_weights(cfg, int(opts.get("seed", 929)), float(opts.get("lm_head_std", 0.05))).items()
            tokenizer = synthetic_tokenizer()
        else:
            cfg, tensors = _load_checkpoint(model_name_or_path, cache_dir)
            from transformers import T5Tokenizer
            tokenizer = T5Tokenizer.from_pretrained(tokenizer_name_or_path or model_name_or_path, cache_dir=cache_dir)
        c = br.make_config(cfg["d_model"], cfg["num_heads"], cfg["d_ff"], cfg["num_layers"], cfg["num_decoder_layers"],
                           vocab_size=cfg["vocab_size"], d_kv=cfg.get("d_kv", 64), rel_buckets=cfg.get("rel_buckets", 32),
                           rel_max_distance=cfg.get("rel_max_distance", 128), layer_norm_eps=cfg.get("layer_norm_eps", 1e-6),
                           gated_gelu=cfg.get("gated_gelu", True), scale_decoder_outputs=cfg.get("scale_decoder_outputs", False),
                           pad_id=cfg.get("pad_id", 0), eos_id=cfg.get("eos_id", 1), max_tokens=max_tokens, max_docs=max_docs)
        engine = br.Engine(c, dev)
        engine.load_state_dict(tensors)
        be = cls(engine, tokenizer, cfg)
        _ENGINE_CACHE[key] = be
        return be

    # ---------------------------------------------------------------- host-side token plumbing
    def assembler(self, template: str):
        """PromptAssembler for `template` sharing this backend's token cache (_prompts.py): a passage tokenised for one prompt is
        reused by every other prompt / ranker on this tokenizer. B200RANK_PROMPT_ASSEMBLY=0 -> None (callers tokenise whole strings)."""
        if os.environ.get("B200RANK_PROMPT_ASSEMBLY", "1") == "0":
            return None
        from ._prompts import PromptAssembler, TokenCache
        if not hasattr(self, "_token_cache"):
            self._token_cache, self._assemblers = TokenCache(), {}
        a = self._assemblers.get(template)
        if a is None:
            a = self._assemblers[template] = PromptAssembler(self.tokenizer, template, cache=self._token_cache)
        return a

    def prompt_rows(self, template: str, fields: Sequence[Dict[str, str]]) -> List[List[int]]:
        """Token-id rows of `template.format(**f)` for every f: what `tokenizer(prompts)` returns for the rendered strings."""
        a = self.assembler(template)
        if a is None:
            return self.tokenize_prompts([template.format(**f) for f in fields])
        return a.rows(fields)

    def tokenize_prompts(self, prompts: Sequence[str]) -> List[List[int]]:
        """`tokenizer(data)` of Text2TextGenerationDataset (pairwise.py:17-26): appends </s>, no padding, no truncation."""
        prompts = list(prompts)
        return self.tokenizer(prompts)["input_ids"] if prompts else []

    @staticmethod
    def pad_rows(rows: Sequence[Sequence[int]], pad_id: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        n = len(rows)
        lengths = np.fromiter((len(r) for r in rows), dtype=np.int32, count=n)
        ids = np.full((n, int(lengths.max()) if n else 1), pad_id, dtype=np.int32)
        for i, r in enumerate(rows):
            ids[i, : len(r)] = r
        return ids, lengths

    # ---------------------------------------------------------------- the four uses of the forward
    def score_yes_no(self, rows, yes_id: int, no_id: int):
        ids, lengths = self.pad_rows(rows, self.pad_id)
        return self.engine.score_yes_no(ids, lengths, yes_id, no_id)

    def submit_yes_no(self, rows, yes_id: int, no_id: int):
        """Asynchronous score_yes_no (two batches in flight, see b200rank_submit_yes_no); returns a ticket for wait_yes_no, or None
        when the batch does not qualify for the pipelined pass (a document of more than 240 tokens, more documents or tokens than one
        device pass holds, a d_kv = 128 model): b200rank_submit_yes_no validates before it enqueues anything and reports those as
        B200RANK_ERR_ARG / B200RANK_ERR_CAPACITY, and the caller then scores the batch through the synchronous entry point (which
        splits it into device passes and raises for arguments that are wrong in themselves)."""
        import b200rank as br
        ids, lengths = self.pad_rows(rows, self.pad_id)
        try:
            return self.engine.submit_yes_no(ids, lengths, yes_id, no_id)
        except br.B200RankError as err:
            if err.code in (br.ERR_ARG, br.ERR_CAPACITY):
                return None
            raise

    def wait_yes_no(self, ticket):
        return self.engine.wait_yes_no(ticket)

    def score_qlm(self, rows, labels: Sequence[int]) -> np.ndarray:
        ids, lengths = self.pad_rows(rows, self.pad_id)
        return self.engine.score_qlm(ids, lengths, labels)

    def label_probs(self, rows, dec_prefix: Sequence[int], cols: Sequence[int]) -> np.ndarray:
        ids, lengths = self.pad_rows(rows, self.pad_id)
        return self.engine.logits_at(ids, lengths, dec_prefix, cols, normalize=True)

    def generate(self, padded_ids: np.ndarray, dec_prefix: Sequence[int], max_new: int) -> np.ndarray:
        """`self.llm.generate(input_ids, decoder_input_ids=prefix, max_new_tokens=max_new)` for one padded batch; returns the
        HF-shaped output [B, len(prefix) + steps] where steps stops early once every row has emitted eos."""
        ids = np.ascontiguousarray(padded_ids, dtype=np.int32)
        if generate_mask_mode() == "infer" and (ids == self.pad_id).any():
            lengths = (ids != self.pad_id).sum(axis=1).astype(np.int32)
            # right padding only (T5Tokenizer pads right): the non-pad prefix is the real row
        else:
            lengths = np.full((ids.shape[0],), ids.shape[1], np.int32)
        new = self.engine.greedy(ids, lengths, dec_prefix, max_new)
        finished = np.zeros(ids.shape[0], bool)
        steps = max_new
        for s in range(max_new):
            finished |= new[:, s] == self.eos_id
            if finished.all():
                steps = s + 1
                break
        prefix = np.tile(np.asarray(dec_prefix, np.int64)[None], (ids.shape[0], 1))
        return np.concatenate([prefix, new[:, :steps].astype(np.int64)], axis=1)


    def generate_batches(self, batches: Sequence[np.ndarray], dec_prefix: Sequence[int], max_new: int) -> List[np.ndarray]:
        """`[self.generate(b, dec_prefix, max_new) for b in batches]` as ONE engine call. Each element is one padded batch of the
        reference's DataLoader (pairwise.py:175-200); what a row sees of its padding is decided per reference batch exactly as in
        generate() (B200RANK_GENERATE_MASK: 'ones' = the batch's pads are attended like tokens, 'infer' = they are masked), so every
        row's tokens — and every batch's early-stopping length — equal those of the one-batch-at-a-time loop; only the engine's
        device passes get larger (a reference batch of 2 x 320 tokens fills 4 % of a B200)."""
        if not batches:
            return []
        infer = generate_mask_mode() == "infer"
        width = max(int(b.shape[1]) for b in batches)
        total = sum(int(b.shape[0]) for b in batches)
        ids = np.full((total, width), self.pad_id, np.int32)
        lengths = np.empty((total,), np.int32)
        r = 0
        for b in batches:
            b = np.asarray(b, np.int32)
            n, s = b.shape
            ids[r:r + n, :s] = b
            lengths[r:r + n] = (b != self.pad_id).sum(axis=1) if infer and (b == self.pad_id).any() else s
            r += n
        new = self.engine.greedy(ids, lengths, dec_prefix, max_new)
        prefix = np.asarray(dec_prefix, np.int64)[None]
        outs, r = [], 0
        for b in batches:
            n = int(b.shape[0])
            blk = new[r:r + n]
            finished = np.zeros(n, bool)
            steps = max_new
            for st in range(max_new):
                finished |= blk[:, st] == self.eos_id
                if finished.all():
                    steps = st + 1
                    break
            outs.append(np.concatenate([np.tile(prefix, (n, 1)), blk[:, :steps].astype(np.int64)], axis=1))
            r += n
        return outs

    def generate_rows(self, rows: Sequence[Sequence[int]], dec_prefix: Sequence[int], max_new: int) -> List[np.ndarray]:
        """B independent `generate(input_ids=[row], ...)` calls (batch of 1 each: no padding, so no mask question) as ONE engine
        call: the engine packs the real tokens of every row, and a row's result does not depend on what else is in the batch.
        Returns one HF-shaped 1-D output per row: prefix + new tokens up to and including eos (or max_new)."""
        if not rows:
            return []
        ids, lengths = self.pad_rows(rows, self.pad_id)
        new = self.engine.greedy(ids, lengths, dec_prefix, max_new)
        return self._trim_rows(new, dec_prefix, max_new)

    def _trim_rows(self, new: np.ndarray, dec_prefix: Sequence[int], max_new: int) -> List[np.ndarray]:
        prefix = np.asarray(dec_prefix, np.int64)
        outs = []
        for r in range(new.shape[0]):
            hit = np.nonzero(new[r, :max_new] == self.eos_id)[0]
            steps = int(hit[0]) + 1 if hit.size else max_new
            outs.append(np.concatenate([prefix, new[r, :steps].astype(np.int64)]))
        return outs

def _checkpoint_files(path: str) -> List[str]:
    """Weight files of a local HF checkpoint directory, in load order: a single `model.safetensors` / `pytorch_model.bin`, or the shards
    an index json lists (`model.safetensors.index.json` / `pytorch_model.bin.index.json` — how the hub stores flan-t5-xl / -xxl).
    safetensors is preferred when both formats are present, like `from_pretrained`."""
    for single, index in (("model.safetensors", "model.safetensors.index.json"), ("pytorch_model.bin", "pytorch_model.bin.index.json")):
        if os.path.exists(os.path.join(path, single)):
            return [os.path.join(path, single)]
        if os.path.exists(os.path.join(path, index)):
            with open(os.path.join(path, index)) as f:
                shards = sorted(set(json.load(f)["weight_map"].values()))
            missing = [sh for sh in shards if not os.path.exists(os.path.join(path, sh))]
            if missing:
                raise FileNotFoundError(f"{path}: {index} lists shards that are not there: {missing}")
            return [os.path.join(path, sh) for sh in shards]
    return []


def _iter_checkpoint_tensors(files: Sequence[str]):
    """(name, fp32 ndarray) for every tensor of every file, one tensor in host memory at a time. Tensors are read through torch so
    that bf16 / fp16 checkpoints load too (numpy has no bfloat16); the engine converts to its own storage types on upload."""
    import torch
    for fn in files:
        if fn.endswith(".safetensors"):
            from safetensors import safe_open
            with safe_open(fn, framework="pt", device="cpu") as f:
                for name in f.keys():
                    yield name, f.get_tensor(name).to(torch.float32).numpy()
        else:
            sd = torch.load(fn, map_location="cpu", weights_only=True)
            for name in list(sd.keys()):
                yield name, sd.pop(name).to(torch.float32).numpy()


def _load_checkpoint(path: str, cache_dir=None):
    """(cfg dict, iterable of (name, fp32 ndarray)) from a local HF checkpoint dir, else via transformers on the CPU."""
    if os.path.isdir(path) and os.path.exists(os.path.join(path, "config.json")):
        with open(os.path.join(path, "config.json")) as f:
            hf = json.load(f)
        cfg = _cfg_from_hf(hf)
        files = _checkpoint_files(path)
        if not files:
            raise FileNotFoundError(f"{path}: no model.safetensors / pytorch_model.bin (single file or sharded with an index json)")
        return cfg, _iter_checkpoint_tensors(files)
    from transformers import AutoConfig, T5ForConditionalGeneration
    hf_cfg = AutoConfig.from_pretrained(path, cache_dir=cache_dir)
    if hf_cfg.model_type != "t5":
        raise NotImplementedError(f"Model type {hf_cfg.model_type} is not supported by the B200 engine (Flan-T5 only)")
    model = T5ForConditionalGeneration.from_pretrained(path, cache_dir=cache_dir)
    cfg = _cfg_from_hf(hf_cfg.to_dict())
    return cfg, ((k, v.float().numpy()) for k, v in model.state_dict().items())


def _cfg_from_hf(hf: Dict) -> Dict:
    if hf.get("model_type", "t5") != "t5":
        raise NotImplementedError(f"Model type {hf.get('model_type')} is not supported by the B200 engine (Flan-T5 only)")
    proj = hf.get("feed_forward_proj", "relu")
    if proj not in ("gated-gelu", "relu"):
        raise NotImplementedError(f"feed_forward_proj={proj!r}: gated-gelu (Flan-T5 / T5 v1.1) and relu (T5 v1.0: monoT5, duoT5) are implemented")
    if hf["d_kv"] not in (64, 128):
        raise NotImplementedError(f"d_kv={hf['d_kv']}: head widths 64 (specialised tcgen05 kernels) and 128 (the 3B T5 v1.0 checkpoints, generic-width "
                                  "attention) are implemented")
    return dict(vocab_size=hf["vocab_size"], d_model=hf["d_model"], d_kv=hf["d_kv"], num_heads=hf["num_heads"], d_ff=hf["d_ff"],
                num_layers=hf["num_layers"], num_decoder_layers=hf.get("num_decoder_layers") or hf["num_layers"],
                rel_buckets=hf.get("relative_attention_num_buckets", 32), rel_max_distance=hf.get("relative_attention_max_distance", 128),
                layer_norm_eps=hf.get("layer_norm_epsilon", 1e-6), gated_gelu=proj == "gated-gelu",
                scale_decoder_outputs=bool(hf.get("tie_word_embeddings", True)), pad_id=hf.get("pad_token_id", 0),
                eos_id=hf.get("eos_token_id", 1))


class ShardedBackend:
    """Document-level sharding across the GPUs of one box (north_star: "candidate (query, passage) prompts shard
    embarrassingly across the 8 GPUs ... no collective inside the scoring loop").

    Wraps the T5Backend of THIS rank (one process per GPU, torch.distributed already initialised) and presents the same interface to
    the rankers. Every rank runs the same host program on the same inputs, so every scoring call sees the same flattened list of
    prompt rows on all ranks; the wrapper scores the contiguous slice [lo, hi) of that list on its own GPU and concatenates the
    ranks' results in rank order with ONE host-side all-gather of the (tiny) result arrays per call — the engine's device passes
    contain no collective. A row's result does not depend on what it is batched with (tests/test_engine_gpu.py: batch-composition
    invariance, bit-exact), so the gathered result equals the single-GPU result bit for bit, and every rank holds the complete
    answer: sort drivers, counters and output assembly proceed identically everywhere; rank 0 writes the run file.

    Where it pays: one query with many prompts — pairwise allpair (`pairwise.py:169-219`: n(n-1) = 9900 prompts for 100 hits),
    pointwise over 1000 hits (`pointwise.py:41-82`) — is strong-scaled over the GPUs. Sort-based methods issue short dependent
    batches (one compare = one prompt); for those run.py shards QUERIES instead (`B200RANK_SHARD=queries`)."""

    def __init__(self, inner: "T5Backend", group=None):
        import torch.distributed as dist
        self.inner, self.group = inner, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def __getattr__(self, name):           # tokenizer, cfg, engine, pad_id, eos_id, assembler, prompt_rows, pad_rows, ...
        return getattr(self.inner, name)

    # -- plumbing
    def _bounds(self, n: int) -> Tuple[int, int]:
        from b200rank.dist import shard_bounds
        return shard_bounds(n, self.rank, self.world)

    def _gather(self, local):
        """Per-rank Python objects in rank order (small numpy arrays: pickled through torch.distributed, gloo or nccl)."""
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, local, group=self.group)
        return parts

    def _cat(self, local: np.ndarray, tail_shape=()) -> np.ndarray:
        parts = [p for p in self._gather(np.ascontiguousarray(local)) if p.shape[0]]
        return np.concatenate(parts, axis=0) if parts else np.zeros((0,) + tuple(tail_shape), local.dtype)

    # -- for callers that shard their HOST work too (PairwiseLlmRanker allpair): which reference batches are this rank's, and a gather
    def shard_batches(self, n_batches: int) -> Tuple[int, int]:
        return self._bounds(n_batches)

    def gather_objects(self, local):
        return self._gather(local)

    # -- the scoring calls, sharded
    def score_yes_no(self, rows, yes_id: int, no_id: int):
        lo, hi = self._bounds(len(rows))
        if hi > lo:
            lg, sc = self.inner.score_yes_no(rows[lo:hi], yes_id, no_id)
        else:
            lg, sc = np.zeros((0, 2), np.float32), np.zeros((0,), np.float32)
        both = self._cat(np.concatenate([np.asarray(lg, np.float32).reshape(-1, 2), np.asarray(sc, np.float32).reshape(-1, 1)], axis=1), (3,))
        return both[:, :2].copy(), both[:, 2].copy()

    def submit_yes_no(self, rows, yes_id: int, no_id: int):
        """Pipelined form: this rank's slice goes through the engine's submit (or, if the slice does not qualify for the pipelined pass,
        is scored synchronously right away); the all-gather happens in wait_yes_no. The decision to pipeline at all is taken on the
        whole row list, which every rank sees, so all ranks take the same path."""
        lo, hi = self._bounds(len(rows))
        ticket, ready = None, None
        if hi > lo:
            ticket = self.inner.submit_yes_no(rows[lo:hi], yes_id, no_id)
            if ticket is None:
                ready = self.inner.score_yes_no(rows[lo:hi], yes_id, no_id)
        else:
            ready = (np.zeros((0, 2), np.float32), np.zeros((0,), np.float32))
        return ("sharded", ticket, ready)

    def wait_yes_no(self, ticket):
        _, t, ready = ticket
        lg, sc = ready if t is None else self.inner.wait_yes_no(t)
        both = self._cat(np.concatenate([np.asarray(lg, np.float32).reshape(-1, 2), np.asarray(sc, np.float32).reshape(-1, 1)], axis=1), (3,))
        return both[:, :2].copy(), both[:, 2].copy()

    def score_qlm(self, rows, labels: Sequence[int]) -> np.ndarray:
        lo, hi = self._bounds(len(rows))
        local = np.asarray(self.inner.score_qlm(rows[lo:hi], labels), np.float32) if hi > lo else np.zeros((0,), np.float32)
        return self._cat(local)

    def label_probs(self, rows, dec_prefix: Sequence[int], cols: Sequence[int]) -> np.ndarray:
        lo, hi = self._bounds(len(rows))
        local = (np.asarray(self.inner.label_probs(rows[lo:hi], dec_prefix, cols), np.float32) if hi > lo
                 else np.zeros((0, len(cols)), np.float32))
        return self._cat(local, (len(cols),))

    def generate_rows(self, rows, dec_prefix: Sequence[int], max_new: int) -> List[np.ndarray]:
        lo, hi = self._bounds(len(rows))
        local = self.inner.generate_rows(rows[lo:hi], dec_prefix, max_new) if hi > lo else []
        return [x for part in self._gather(list(local)) for x in part]

    def generate_batches(self, batches, dec_prefix: Sequence[int], max_new: int) -> List[np.ndarray]:
        """Whole reference batches are the unit (a batch's padding / early-stopping semantics are per batch): batches [lo, hi) here."""
        lo, hi = self._bounds(len(batches))
        local = self.inner.generate_batches(batches[lo:hi], dec_prefix, max_new) if hi > lo else []
        return [x for part in self._gather(list(local)) for x in part]

    def generate(self, padded_ids: np.ndarray, dec_prefix: Sequence[int], max_new: int) -> np.ndarray:
        """One padded batch is not split (its early-stopping length depends on all of its rows, and compare batches are one or two
        prompts): every rank computes it — identical bits — and no communication is needed."""
        return self.inner.generate(padded_ids, dec_prefix, max_new)


def shard_mode(ranker_kind: str, method: str) -> str:
    """'docs' or 'queries': how run.py / bench.py split work over ranks (B200RANK_SHARD overrides). Document-level sharding for the
    rankers whose per-query work is one big independent prompt list (pointwise yes_no / qlm, MonoT5, pairwise allpair); query-level
    sharding for the sort drivers (heapsort / bubblesort / sliding windows), whose compares are short dependent batches."""
    forced = os.environ.get("B200RANK_SHARD", "").lower()
    if forced in ("docs", "queries"):
        return forced
    if ranker_kind == "pointwise" or (ranker_kind == "pairwise" and method == "allpair"):
        return "docs"
    return "queries"
