"""Pairwise ranker on the B200 engine — drop-in for the T5 branch of the reference's llmrankers/pairwise.py.

Keeps the prompt, the `<pad> Passage` decoder prefix, both-orders comparison, exact-string win/conflict scoring,
counters and output assembly (pairwise.py:30-63, 84-103, 164-290). allpair sends all 2*C(n,2) prompts through the
engine in the reference's batch order (batch composition matters here: generate() is called without an attention
mask, so how pads are treated follows B200RANK_GENERATE_MASK, see _backend.generate_mask_mode).
"""
import copy
import os
from collections import defaultdict
from itertools import combinations
from typing import List, Optional

from ._backend import T5Backend
from ._sorting import (binary_heap_top_k, binary_heap_top_k_batched, binary_heap_top_k_rounds, pairwise_bubble_rounds,
                       pairwise_bubble_top_k)
from .rankers import LlmRanker, SearchResult
from .setwise import _assemble

PAIRWISE_PROMPT = """Given a query "{query}", which of the following two passages is more relevant to the query?

Passage A: "{doc1}"

Passage B: "{doc2}"

Output Passage A or Passage B:"""


class Text2TextGenerationDataset:
    """pairwise.py:17-26 — tokenises all prompts up front (appends </s>, no padding, no truncation)."""

    def __init__(self, data: List[str], tokenizer):
        self.data = tokenizer(list(data))

    def __len__(self):
        return len(self.data['input_ids'])

    def __getitem__(self, item):
        return {'input_ids': self.data['input_ids'][item], 'attention_mask': self.data['attention_mask'][item]}


class PairwiseLlmRanker(LlmRanker):
    def __init__(self, model_name_or_path, tokenizer_name_or_path, device, method="allpair", batch_size=2, k=10, cache_dir=None,
                 *, backend: Optional[T5Backend] = None):
        self.device = device
        self.method = method
        self.batch_size = batch_size
        self.k = k
        self.prompt = PAIRWISE_PROMPT
        self.backend = backend or T5Backend.load(model_name_or_path, tokenizer_name_or_path, device, cache_dir)
        self.tokenizer = self.backend.tokenizer
        self.llm = self.backend.engine
        self.config = self.backend.cfg
        self.decoder_input_ids = self.tokenizer.encode("<pad> Passage", add_special_tokens=False)
        self.total_compare = 0
        self.total_completion_tokens = 0
        self.total_prompt_tokens = 0

    def compare(self, query: str, docs: List):
        """Both presentation orders of (doc1, doc2) in one padded batch of 2 -> two decoded strings (pairwise.py:84-103)."""
        self.total_compare += 1
        doc1, doc2 = docs[0], docs[1]
        rows = self.backend.prompt_rows(self.prompt, [dict(query=query, doc1=doc1, doc2=doc2), dict(query=query, doc1=doc2, doc2=doc1)])
        ids, _ = self.backend.pad_rows(rows, self.backend.pad_id)
        self.total_prompt_tokens += ids.shape[0] * ids.shape[1]
        out = self.backend.generate(ids, self.decoder_input_ids, 2)
        self.total_completion_tokens += out.shape[0] * out.shape[1]
        return self.tokenizer.batch_decode(out.tolist(), skip_special_tokens=True)

    def _first_wins(self, query: str, a, b) -> bool:
        out = self.compare(query, [a.text, b.text])
        return out[0] == "Passage A" and out[1] == "Passage B"

    def _compare_items(self, items: List) -> List:
        """Verdicts for several independent (query, a, b) compares in one engine call: every pair stays its own padded batch of
        two (T5Backend.generate_batches), so each verdict equals compare()'s. Returns [(first_wins, prompt_tokens,
        completion_tokens)]; counters are left to the caller (the pairs may belong to different queries)."""
        fields = []
        for query, a, b in items:
            fields.append(dict(query=query, doc1=a.text, doc2=b.text))
            fields.append(dict(query=query, doc1=b.text, doc2=a.text))
        rows = self.backend.prompt_rows(self.prompt, fields)
        batches = [self.backend.pad_rows(rows[i:i + 2], self.backend.pad_id)[0] for i in range(0, len(rows), 2)]
        res = []
        for ids, out in zip(batches, self.backend.generate_batches(batches, self.decoder_input_ids, 2)):
            txt = self.tokenizer.batch_decode(out.tolist(), skip_special_tokens=True)
            res.append((txt[0] == "Passage A" and txt[1] == "Passage B", ids.shape[0] * ids.shape[1], out.shape[0] * out.shape[1]))
        return res

    def _has_batched_compares(self) -> bool:
        """True when `_compare_items` answers what this class's compare() answers (a subclass that overrides the single compare
        without its batched twin falls back to the sequential drivers)."""
        cls = type(self)
        owner = lambda name: next(c for c in cls.__mro__ if name in c.__dict__)   # noqa: E731
        return owner("_compare_items") is owner("_first_wins")

    def _first_wins_many(self, query: str, pairs: List) -> List[bool]:
        """`_first_wins` for several independent pairs of one query; counters as len(pairs) compare() calls."""
        res = self._compare_items([(query, a, b) for a, b in pairs])
        self.total_compare += len(res)
        self.total_prompt_tokens += sum(r[1] for r in res)
        self.total_completion_tokens += sum(r[2] for r in res)
        return [r[0] for r in res]

    def rerank_many(self, requests, window: int = 8):
        """Extension (not in the reference): rerank an iterable of (query, ranking) pairs (heapsort or bubblesort) with up to `window`
        queries' sorts advancing in lockstep, every round one engine batch of all their pending pair compares (see
        SetwiseLlmRanker.rerank_many). Per-query compares, order, scores and counters are exactly rerank()'s. allpair (already one
        large batch per query) and subclasses with their own single compare but no batched twin fall back to rerank()."""
        if self.method not in ("heapsort", "bubblesort") or not self._has_batched_compares():
            for query, ranking in requests:
                yield self.rerank(query, ranking)
            return
        heap = self.method == "heapsort"
        it = iter(requests)
        active, done, next_out, seq = [], {}, 0, 0

        def finish(st):
            ranking = st["arr"]
            if heap:
                ranking = [SearchResult(docid=doc.docid, score=-i, text=None) for i, doc in enumerate(reversed(st["arr"]))]
            done[st["seq"]] = (_assemble(ranking, st["original"], self.k), st["counters"])

        def admit():
            nonlocal seq
            while len(active) < max(1, window):
                try:
                    query, ranking = next(it)
                except StopIteration:
                    return
                st = dict(seq=seq, query=query, original=copy.deepcopy(ranking), arr=list(ranking), counters=[0, 0, 0])
                st["gen"] = binary_heap_top_k_rounds(st["arr"], self.k) if heap else pairwise_bubble_rounds(st["arr"], self.k)
                seq += 1
                try:
                    st["round"] = next(st["gen"])
                    active.append(st)
                except StopIteration:
                    finish(st)

        admit()
        while active or next_out in done:
            while next_out in done:
                result, c = done.pop(next_out)
                self.total_compare, self.total_prompt_tokens, self.total_completion_tokens = c
                next_out += 1
                yield result
            if not active:
                break
            res = self._compare_items([(st["query"], a, b) for st in active for a, b in st["round"]])
            pos, still = 0, []
            for st in active:
                n = len(st["round"])
                mine = res[pos:pos + n]
                pos += n
                st["counters"][0] += n
                st["counters"][1] += sum(r[1] for r in mine)
                st["counters"][2] += sum(r[2] for r in mine)
                try:
                    st["round"] = st["gen"].send([r[0] for r in mine])
                    still.append(st)
                except StopIteration:
                    finish(st)
            active[:] = still
            admit()

    def rerank(self, query: str, ranking: List[SearchResult]) -> List[SearchResult]:
        original_ranking = copy.deepcopy(ranking)
        self.total_compare = 0
        self.total_completion_tokens = 0
        self.total_prompt_tokens = 0
        if self.method == "allpair":
            doc_pairs = list(combinations(ranking, 2))

            def field(r):      # prompt row r: pair r // 2 in the order (d1, d2) for even r, (d2, d1) for odd r (pairwise.py:169-174)
                d1, d2 = doc_pairs[r >> 1]
                return dict(query=query, doc1=d1.text, doc2=d2.text) if r % 2 == 0 else dict(query=query, doc1=d2.text, doc2=d1.text)
            n_rows = 2 * len(doc_pairs)
            n_batches = (n_rows + self.batch_size - 1) // self.batch_size
            # Under document-level sharding (ShardedBackend) this rank assembles, pads and generates only ITS contiguous range of the
            # reference's DataLoader batches — host work shards with the GPU work — and the per-batch outputs and counters are gathered
            # once per query. A batch's result does not depend on which rank or engine call it lands in.
            sharded = hasattr(self.backend, "shard_batches")
            b_lo, b_hi = self.backend.shard_batches(n_batches) if sharded else (0, n_batches)
            be = self.backend.inner if sharded else self.backend
            r_lo, r_hi = b_lo * self.batch_size, min(b_hi * self.batch_size, n_rows)
            rows = be.prompt_rows(self.prompt, [field(r) for r in range(r_lo, r_hi)]) if r_hi > r_lo else []
            outputs = []
            # the reference's DataLoader batches (batch_size rows, padded to the batch's longest), many of them per engine call
            batches = [be.pad_rows(rows[i:i + self.batch_size], be.pad_id)[0] for i in range(0, len(rows), self.batch_size)]
            per_call = max(1, 4096 // max(1, self.batch_size)) if os.environ.get("B200RANK_BATCHED_SORT", "1") != "0" else 1
            for c0 in range(0, len(batches), per_call):
                chunk = batches[c0:c0 + per_call]
                for ids, out in zip(chunk, be.generate_batches(chunk, self.decoder_input_ids, 2)):
                    self.total_compare += 1
                    self.total_prompt_tokens += ids.shape[0] * ids.shape[1]
                    self.total_completion_tokens += out.shape[0] * out.shape[1]
                    outputs.extend(out.tolist())
            if sharded:
                parts = self.backend.gather_objects((outputs, self.total_compare, self.total_prompt_tokens, self.total_completion_tokens))
                outputs = [o for part in parts for o in part[0]]
                self.total_compare = sum(part[1] for part in parts)
                self.total_prompt_tokens = sum(part[2] for part in parts)
                self.total_completion_tokens = sum(part[3] for part in parts)
            # fewer than two documents: no pairs, nothing to decode (the reference raises IndexError in tokenizer([]) here; the
            # drop-in returns the trivial ranking)
            outputs = self.tokenizer.batch_decode(outputs, skip_special_tokens=True) if outputs else []
            scores = defaultdict(float)
            for i in range(0, len(outputs), 2):
                d1, d2 = doc_pairs[i // 2]
                if outputs[i] == "Passage A" and outputs[i + 1] == "Passage B":
                    scores[d1.docid] += 1
                elif outputs[i] == "Passage B" and outputs[i + 1] == "Passage A":
                    scores[d2.docid] += 1
                else:  # conflict
                    scores[d1.docid] += 0.5
                    scores[d2.docid] += 0.5
            ranking = sorted([SearchResult(docid=docid, score=score, text=None) for docid, score in scores.items()],
                             key=lambda x: x.score, reverse=True)
        elif self.method == "heapsort":
            arr = list(ranking)
            if self._has_batched_compares() and os.environ.get("B200RANK_BATCHED_SORT", "1") != "0":
                binary_heap_top_k_batched(arr, self.k, lambda pairs: self._first_wins_many(query, pairs))   # level-parallel build
            else:
                binary_heap_top_k(arr, self.k, lambda a, b: self._first_wins(query, a, b))
            ranking = [SearchResult(docid=doc.docid, score=-i, text=None) for i, doc in enumerate(reversed(arr))]
        elif self.method == "bubblesort":
            pairwise_bubble_top_k(ranking, self.k, lambda a, b: self._first_wins(query, a, b))
        else:
            raise NotImplementedError(f'Method {self.method} is not implemented.')
        return _assemble(ranking, original_ranking, self.k)

    def truncate(self, text, length):
        return self.tokenizer.convert_tokens_to_string(self.tokenizer.tokenize(text)[:length])


DUOT5_PROMPT = 'Query: {query} Document0: {doc1} Document1: {doc2} Relevant:'


class DuoT5LlmRanker(PairwiseLlmRanker):
    """pairwise.py:296-352 — duoT5 (T5 v1.0: relu feed-forward, tied embeddings). compare() scores both presentation orders
    in one batch of two, P(true) = softmax(logits[:, 0, [6136, 1176]])[:, 1], and returns P(order 1) > P(order 2);
    rerank() supports heapsort only, like the reference."""

    def compare(self, query: str, docs: List[str]) -> bool:
        self.total_compare += 1
        self.prompt = DUOT5_PROMPT
        inputs = [self.prompt.format(query=query, doc1=docs[0], doc2=docs[1]),
                  self.prompt.format(query=query, doc1=docs[1], doc2=docs[0])]
        rows = self.tokenizer(inputs, truncation=True)["input_ids"]   # padding=True happens in pad_rows
        self.total_prompt_tokens += 2 * max(len(r) for r in rows)
        _, probs = self.backend.score_yes_no(rows, 1176, 6136)
        return bool(probs[0] > probs[1])

    def _first_wins(self, query: str, a, b) -> bool:
        return self.compare(query, [a.text, b.text])

    def _compare_items(self, items: List) -> List:
        """compare() for several independent (query, a, b) pairs in ONE engine call: the reference pads each pair to its longer
        prompt and masks the pads (padding=True with the attention mask, pairwise.py:303-311), the engine computes on real tokens
        only, and a row's result does not depend on its neighbours — so every verdict equals the sequential compare()'s. Returns
        [(first_wins, prompt_tokens, completion_tokens)]; counters are left to the caller."""
        inputs = []
        for query, a, b in items:
            inputs.append(DUOT5_PROMPT.format(query=query, doc1=a.text, doc2=b.text))
            inputs.append(DUOT5_PROMPT.format(query=query, doc1=b.text, doc2=a.text))
        rows = self.tokenizer(inputs, truncation=True)["input_ids"] if inputs else []
        probs = self.backend.score_yes_no(rows, 1176, 6136)[1] if rows else []
        return [(bool(probs[i] > probs[i + 1]), 2 * max(len(rows[i]), len(rows[i + 1])), 0) for i in range(0, len(rows), 2)]

    def rerank(self, query: str, ranking: List[SearchResult]) -> List[SearchResult]:
        if self.method != "heapsort":
            raise NotImplementedError(f'Method {self.method} is not implemented.')
        return super().rerank(query, ranking)


class OpenAiPairwiseLlmRanker(PairwiseLlmRanker):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("OpenAI-backed rankers are a remote API, outside the B200 engine's scope (SURVEY.md §2.1)")
