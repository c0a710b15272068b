"""Listwise rankers are outside the hot path this engine replaces (free-form permutation
generation / OpenAI API). The names exist so that `run.py`'s imports resolve; constructing one fails loudly."""
from .rankers import LlmRanker


class OpenAiListwiseLlmRanker(LlmRanker):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("listwise ranking is outside the B200 engine's scope (SURVEY.md §2.1)")


class ListwiseLlmRanker(LlmRanker):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("listwise ranking is outside the B200 engine's scope (SURVEY.md §2.1)")
