"""Public base types of the drop-in package.

`SearchResult` and `LlmRanker` keep the names, fields and method signatures user code and run.py rely on (reference
llmrankers/rankers.py:5-17): a result is (docid, score, text); a ranker exposes rerank(query, ranking) and truncate(text, length)
and, after every rerank, the counters total_compare / total_prompt_tokens / total_completion_tokens.

On top of that interface every ranker of this package has `rerank_many(requests)`: the same results as calling rerank() per
(query, ranking) pair, in order, with that query's counters readable after each yield. The base implementation below is the
plain loop; the pointwise / setwise / pairwise rankers override it to keep the GPU busy across queries (two queries in flight
with tokenisation look-ahead; sort-based rerankers advancing several queries' compares as one batch).
"""
from dataclasses import dataclass
from typing import Iterable, Iterator, List, Optional, Tuple


@dataclass
class SearchResult:
    """One candidate: `score` is the first-stage score on input and the ranker's score on output; the sort-based rankers return
    new objects with `text=None` and `score=-rank`, the pointwise rankers return the input objects re-ordered."""
    docid: str
    score: float
    text: Optional[str]


class LlmRanker:
    total_compare: int = 0             # LLM calls of the last rerank(): batches (pointwise, allpair) or compares (sorts)
    total_prompt_tokens: int = 0       # padded encoder tokens + decoder input tokens, as the reference counts them
    total_completion_tokens: int = 0   # generated tokens (generation scoring only)

    def rerank(self, query: str, ranking: List[SearchResult]) -> List[SearchResult]:
        """Re-ordered candidates for `query`, best first."""
        raise NotImplementedError

    def truncate(self, text: str, length: int) -> str:
        """`text` cut to its first `length` tokenizer pieces, detokenised."""
        raise NotImplementedError

    def rerank_many(self, requests: Iterable[Tuple[str, List[SearchResult]]]) -> Iterator[List[SearchResult]]:
        for query, ranking in requests:
            yield self.rerank(query, ranking)
