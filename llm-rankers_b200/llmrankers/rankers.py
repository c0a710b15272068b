"""Public base types — interface-identical to the reference's llmrankers/rankers.py:5-17."""
from dataclasses import dataclass
from typing import List


@dataclass
class SearchResult:
    docid: str
    score: float
    text: str


class LlmRanker:
    """rerank(query, ranking) -> re-ordered list of SearchResult; truncate(text, length) -> str."""

    def rerank(self, query: str, ranking: List[SearchResult]) -> List[SearchResult]:
        raise NotImplementedError

    def truncate(self, text, length):
        raise NotImplementedError
