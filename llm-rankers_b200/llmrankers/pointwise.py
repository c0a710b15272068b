"""Pointwise rankers on the B200 engine — drop-in for the reference's llmrankers/pointwise.py.

Kept from the reference (pointwise.py:13-133): constructor keywords, the two prompts, yes/no token ids, the
`<pad> {query}` label construction, per-batch counters, in-place score assignment in input order and the stable
descending sort. Replaced: tokenise -> DataLoader(4 worker forks) -> padded HF forward -> per-document .item()
becomes one tokenizer call, one C-ABI call for the whole candidate list (the engine packs real tokens and splits
device passes itself; per-document results do not depend on batch composition) and one D2H copy.
"""
import os
from collections import deque
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

    def _rows(self, template: str, fields: List[dict]) -> List[List[int]]:
        """Token-id rows of `template.format(**f)` for every f — what `Text2TextGenerationDataset` (pairwise.py:17-26) produces,
        through token-level assembly + a per-document token cache (_prompts.py; B200RANK_PROMPT_ASSEMBLY=0 tokenises whole strings)."""
        return self.backend.prompt_rows(template, fields)

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
            rows = self._rows(QLM_PROMPT, [dict(text=doc.text) for doc in ranking])
            labels = self.tokenizer.encode(f"<pad> {query}", add_special_tokens=False)  # pointwise.py:58-60
            self._count_batches(rows, len(labels))
            scores = self.backend.score_qlm(rows, labels) if rows else []
            for doc, s in zip(ranking, scores):
                doc.score = float(s)
        elif self.method == "yes_no":
            yes_id = self.tokenizer.encode("Yes", add_special_tokens=False)[0]
            no_id = self.tokenizer.encode("No", add_special_tokens=False)[0]
            rows = self._rows(YES_NO_PROMPT, [dict(text=doc.text, query=query) for doc in ranking])
            self._count_batches(rows, 1)
            if rows:
                _, scores = self.backend.score_yes_no(rows, yes_id, no_id)
                for doc, s in zip(ranking, scores):
                    doc.score = float(s)
        # any other method: like the reference, nothing is scored and the input order is sorted by its existing scores
        return sorted(ranking, key=lambda x: x.score, reverse=True)

    def _pipeline_spec(self):
        """(prompt template, field builder, yes id, no id) of the single-decoder-position scoring this ranker does, or None if
        rerank_many has to fall back to rerank() (qlm: 33 decoder positions, not pipelined)."""
        if self.method != "yes_no":
            return None
        return (YES_NO_PROMPT, lambda query, doc: dict(text=doc.text, query=query),
                self.tokenizer.encode("Yes", add_special_tokens=False)[0], self.tokenizer.encode("No", add_special_tokens=False)[0])

    def rerank_many(self, requests, tokenizer_threads: int = 4, lookahead: int = 8, queries_per_pass: Optional[int] = None):
        """Extension (not in the reference): rerank an iterable of (query, ranking) pairs as a pipeline. Upcoming queries are
        tokenised on `tokenizer_threads` worker threads (the Rust tokenizer releases the GIL), up to `lookahead` queries ahead, and
        two device passes are in flight on the GPU — pass i+1's encoder runs while pass i's decoder chain finishes — each holding the
        documents of up to `queries_per_pass` consecutive queries (default B200RANK_QUERIES_PER_PASS = 2). Yields the
        same list `rerank(query, ranking)` would return for each pair, in order; counters hold the totals of the last query.
        Only the yes_no method is pipelined (the headline path); other methods fall back to rerank()."""
        spec = self._pipeline_spec()
        from concurrent.futures import ThreadPoolExecutor
        if spec is None and self.method == "qlm" and type(self).rerank is PointwiseLlmRanker.rerank:
            # qlm (33 decoder positions) is not pipelined on the GPU, but the host work is: upcoming queries are tokenised on worker
            # threads while the engine scores the current one (ctypes releases the GIL during the call) — configs[4] of BASELINE.json
            # tokenises 1000 passages per query
            def tokenise_qlm(query, ranking):
                return (self._rows(QLM_PROMPT, [dict(text=doc.text) for doc in ranking]),
                        self.tokenizer.encode(f"<pad> {query}", add_special_tokens=False))
            it = iter(requests)
            window = deque()
            with ThreadPoolExecutor(max(1, tokenizer_threads)) as pool:
                def refill_qlm():
                    while len(window) < max(1, lookahead):
                        try:
                            query, ranking = next(it)
                        except StopIteration:
                            return
                        window.append((ranking, pool.submit(tokenise_qlm, query, ranking)))
                refill_qlm()
                while window:
                    ranking, fut = window.popleft()
                    rows, labels = fut.result()
                    refill_qlm()
                    self.total_compare = 0
                    self.total_completion_tokens = 0
                    self.total_prompt_tokens = 0
                    self._count_batches(rows, len(labels))
                    scores = self.backend.score_qlm(rows, labels) if rows else []
                    for doc, sc in zip(ranking, scores):
                        doc.score = float(sc)
                    yield sorted(ranking, key=lambda x: x.score, reverse=True)
            return
        if spec is None:
            for query, ranking in requests:
                yield self.rerank(query, ranking)
            return
        template, fields_of, yes_id, no_id = spec
        # Several queries per device pass: the decoder chain of a pass (~250 small launches next to the following pass's encoder GEMMs)
        # costs about the same for 100 documents as for 400, so merging consecutive queries amortises it (B200: +3 % docs/s at two
        # queries per pass, profiles/r02_bench_queries_per_step_ab.txt). It rests on the engine's batch-composition invariance — a
        # document's logits are bit-identical whatever shares its pass — so only backends that state it (`batch_invariant`) merge.
        if queries_per_pass is None:
            queries_per_pass = int(os.environ.get("B200RANK_QUERIES_PER_PASS", "2"))
        if not getattr(self.backend, "batch_invariant", False):
            queries_per_pass = 1
        group_size = [max(1, queries_per_pass)]

        def finish(item):
            ticket, members, scores = item
            if ticket is not None:
                _, scores = self.backend.wait_yes_no(ticket)
            out, off = [], 0
            for ranking, rows in members:
                self.total_compare = 0
                self.total_completion_tokens = 0
                self.total_prompt_tokens = 0
                self._count_batches(rows, 1)
                if scores is not None:
                    for doc, s in zip(ranking, scores[off:off + len(rows)]):
                        doc.score = float(s)
                off += len(rows)
                out.append(sorted(ranking, key=lambda x: x.score, reverse=True))
            return out

        def tokenise(query, ranking):
            return self._rows(template, [fields_of(query, doc) for doc in ranking])

        class _Ready:   # a tokenised query handed back to the window
            def __init__(self, rows):
                self.rows = rows

            def result(self):
                return self.rows

        it = iter(requests)
        window = deque()   # (ranking, future of rows), in request order
        pending = None     # (ticket, [(ranking, rows)], scores) of the pass whose decoder chain is still running
        with ThreadPoolExecutor(max(1, tokenizer_threads)) as pool:
            def refill():
                while len(window) < max(1, lookahead):
                    try:
                        query, ranking = next(it)
                    except StopIteration:
                        return
                    window.append((ranking, pool.submit(tokenise, query, ranking)))
            try:
                refill()
                while window:
                    members = []
                    while window and len(members) < group_size[0]:
                        ranking, fut = window.popleft()
                        members.append((ranking, fut.result()))
                        refill()
                    rows = [r for _, rs in members for r in rs]
                    ticket = self.backend.submit_yes_no(rows, yes_id, no_id) if rows else None
                    if rows and ticket is None and len(members) > 1:
                        # the merged pass was declined (capacity of one device pass, or a long document in one of the queries): hand the
                        # other queries back and go on one query per pass
                        for ranking, rs in reversed(members[1:]):
                            window.appendleft((ranking, _Ready(rs)))
                        members, group_size[0] = members[:1], 1
                        rows = members[0][1]
                        ticket = self.backend.submit_yes_no(rows, yes_id, no_id) if rows else None
                    scores = None
                    if rows and ticket is None:
                        # not a pipelined batch (long documents / more than one device pass): drain what is in flight — the
                        # synchronous entry points refuse to run next to pipelined batches — then score this query as rerank() does
                        if pending is not None:
                            item, pending = pending, None
                            yield from finish(item)
                        _, scores = self.backend.score_yes_no(rows, yes_id, no_id)
                    if pending is not None:
                        item, pending = pending, (ticket, members, scores)
                        yield from finish(item)
                    else:
                        pending = (ticket, members, scores)
                if pending is not None:
                    item, pending = pending, None
                    yield from finish(item)
            finally:
                # the consumer stopped early or something failed between submit and wait: do not leave a ticket in flight (the
                # engine's synchronous entry points refuse to run until every pipelined batch has been waited for)
                if pending is not None and pending[0] is not None:
                    try:
                        self.backend.wait_yes_no(pending[0])
                    except Exception:   # noqa: BLE001 - already unwinding
                        pass

    def truncate(self, text, length):
        return self.tokenizer.convert_tokens_to_string(self.tokenizer.tokenize(text)[:length])


MONOT5_PROMPT = "Query: {query} Document: {document} Relevant:"
MONOT5_FALSE_ID, MONOT5_TRUE_ID = 6136, 1176  # "the indexes of the tokens false and true in T5" (pointwise.py:176)


class MonoT5LlmRanker(PointwiseLlmRanker):
    """pointwise.py:136-186 — monoT5 checkpoints are T5 v1.0 (relu feed-forward, tied embeddings => logits scaled by
    d_model^-0.5); the score is softmax(logits[:, 0, [false, true]])[:, 1], i.e. the yes_no entry point of the engine with
    (yes, no) = (true, false). Counters as the reference: one compare per batch, B x longest row + B decoder tokens."""

    def _pipeline_spec(self):
        return (MONOT5_PROMPT, lambda query, doc: dict(query=query, document=doc.text), MONOT5_TRUE_ID, MONOT5_FALSE_ID)

    def rerank(self, query: str, ranking: List[SearchResult]) -> List[SearchResult]:
        self.total_compare = 0
        self.total_completion_tokens = 0
        self.total_prompt_tokens = 0
        rows = self._rows(MONOT5_PROMPT, [dict(query=query, document=doc.text) for doc in ranking])
        self._count_batches(rows, 1)
        if rows:
            _, scores = self.backend.score_yes_no(rows, MONOT5_TRUE_ID, MONOT5_FALSE_ID)
            for doc, s in zip(ranking, scores):
                doc.score = float(s)
        return sorted(ranking, key=lambda x: x.score, reverse=True)
