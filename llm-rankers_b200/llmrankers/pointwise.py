"""Pointwise rankers on the B200 engine — drop-in for the reference's llmrankers/pointwise.py.

Kept from the reference (pointwise.py:13-133): constructor keywords, the two prompts, yes/no token ids, the
`<pad> {query}` label construction, per-batch counters, in-place score assignment in input order and the stable
descending sort. Replaced: tokenise -> DataLoader(4 worker forks) -> padded HF forward -> per-document .item()
becomes one tokenizer call, one C-ABI call for the whole candidate list (the engine packs real tokens and splits
device passes itself; per-document results do not depend on batch composition) and one D2H copy.
"""
from typing import List, Optional

from ._backend import T5Backend
from .rankers import LlmRanker, SearchResult

YES_NO_PROMPT = "Passage: {text}\nQuery: {query}\nDoes the passage answer the query? Answer 'Yes' or 'No'"
QLM_PROMPT = "Passage: {text}\nPlease write a question based on this passage."


class PointwiseLlmRanker(LlmRanker):
    def __init__(self, model_name_or_path, tokenizer_name_or_path, device, method="qlm", batch_size=1, cache_dir=None,
                 *, backend: Optional[T5Backend] = None):
        # `backend` (keyword-only, not in the reference) lets callers share one loaded engine between rankers
        self.backend = backend or T5Backend.load(model_name_or_path, tokenizer_name_or_path, device, cache_dir)
        self.tokenizer = self.backend.tokenizer
        self.llm = self.backend.engine
        self.config = self.backend.cfg
        self.device = device
        self.method = method
        self.batch_size = batch_size
        self.total_compare = 0
        self.total_completion_tokens = 0
        self.total_prompt_tokens = 0

    def _count_batches(self, rows: List[List[int]], dec_len: int) -> None:
        """Counters exactly as the reference accumulates them per DataLoader batch (pointwise.py:64-70, 106-115):
        one 'compare' per batch; prompt tokens = B x (longest row of the batch) + B x decoder length."""
        bs = self.batch_size
        for i in range(0, len(rows), bs):
            chunk = rows[i:i + bs]
            self.total_compare += 1
            self.total_prompt_tokens += len(chunk) * max(len(r) for r in chunk) + len(chunk) * dec_len

    def rerank(self, query: str, ranking: List[SearchResult]) -> List[SearchResult]:
        self.total_compare = 0
        self.total_completion_tokens = 0
        self.total_prompt_tokens = 0
        if self.method == "qlm":
            rows = self.backend.tokenize_prompts([QLM_PROMPT.format(text=doc.text) for doc in ranking])
            labels = self.tokenizer.encode(f"<pad> {query}", add_special_tokens=False)  # pointwise.py:58-60
            self._count_batches(rows, len(labels))
            scores = self.backend.score_qlm(rows, labels) if rows else []
            for doc, s in zip(ranking, scores):
                doc.score = float(s)
        elif self.method == "yes_no":
            yes_id = self.tokenizer.encode("Yes", add_special_tokens=False)[0]
            no_id = self.tokenizer.encode("No", add_special_tokens=False)[0]
            rows = self.backend.tokenize_prompts([YES_NO_PROMPT.format(text=doc.text, query=query) for doc in ranking])
            self._count_batches(rows, 1)
            if rows:
                _, scores = self.backend.score_yes_no(rows, yes_id, no_id)
                for doc, s in zip(ranking, scores):
                    doc.score = float(s)
        # any other method: like the reference, nothing is scored and the input order is sorted by its existing scores
        return sorted(ranking, key=lambda x: x.score, reverse=True)

    def rerank_many(self, requests):
        """Extension (not in the reference): rerank an iterable of (query, ranking) pairs with two queries in flight on the GPU
        — query i+1 is tokenised, copied and its encoder pass started while query i's decoder pass finishes. Yields the same
        list `rerank(query, ranking)` would return for each pair, in order; counters hold the totals of the last query.
        Only the yes_no method is pipelined (the headline path); other methods fall back to rerank()."""
        if self.method != "yes_no":
            for query, ranking in requests:
                yield self.rerank(query, ranking)
            return
        yes_id = self.tokenizer.encode("Yes", add_special_tokens=False)[0]
        no_id = self.tokenizer.encode("No", add_special_tokens=False)[0]
        pending = None  # (ticket, ranking, rows)

        def finish(item):
            ticket, ranking, rows = item
            self.total_compare = 0
            self.total_completion_tokens = 0
            self.total_prompt_tokens = 0
            self._count_batches(rows, 1)
            if ticket is not None:
                _, scores = self.backend.wait_yes_no(ticket)
                for doc, s in zip(ranking, scores):
                    doc.score = float(s)
            return sorted(ranking, key=lambda x: x.score, reverse=True)

        for query, ranking in requests:
            rows = self.backend.tokenize_prompts([YES_NO_PROMPT.format(text=doc.text, query=query) for doc in ranking])
            ticket = self.backend.submit_yes_no(rows, yes_id, no_id) if rows else None
            if pending is not None:
                yield finish(pending)
            pending = (ticket, ranking, rows)
        if pending is not None:
            yield finish(pending)

    def truncate(self, text, length):
        return self.tokenizer.convert_tokens_to_string(self.tokenizer.tokenize(text)[:length])


class MonoT5LlmRanker(PointwiseLlmRanker):
    """pointwise.py:136-186 — monoT5 checkpoints are T5 v1.0 (ungated ReLU feed-forward, tied embeddings); the engine
    implements the gated-GELU Flan-T5 family only, so construction fails loudly at weight load."""

    def rerank(self, query: str, ranking: List[SearchResult]) -> List[SearchResult]:
        raise NotImplementedError("MonoT5 (T5 v1.0 relu feed-forward) is not implemented by the B200 engine yet")
