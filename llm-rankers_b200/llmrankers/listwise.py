"""Listwise ranker on the B200 engine — drop-in for the T5 branch of the reference's llmrankers/listwise.py.

`ListwiseLlmRanker` keeps the reference's constructor keywords, sliding-window `rerank` (listwise.py:177-195: windows of
`window_size` documents moved up by `step_size`, `num_repeat` passes, final score = -position), `compare` and counters:
  * scoring='likelihood' (listwise.py:273-286): the setwise prompt, decoder prefix `<pad> Passage`, full-vocabulary softmax at the
    last prefix position gathered at the label tokens A..W, labels sorted by probability -> "[i]>[j]>..." — one `logits_at` call
    on the engine, the same forward as SetwiseLlmRanker's likelihood mode;
  * scoring='generation' (listwise.py:246-256): the RankGPT completion prompt, truncated like `tokenizer(..., truncation=True)`,
    then free-form greedy decoding `self.llm.generate(input_ids)` — the engine's `greedy` entry point with its self-attention K/V
    cache (one call covers the default budget; longer budgets continue in chunks of GREEDY_CHUNK new tokens whose decoder prefix
    is what has been generated) until </s> or the generation budget.
The response parser (`clean_response` / `remove_duplicate` / `receive_permutation`, listwise.py:110-144) is host logic and is pinned
to the reference by fixtures (tests/golden/make_golden_listwise.py).

`rerank_many` (extension): the window chain of one query is sequential, so several queries advance in lockstep and every round is one
engine batch of their pending windows; each query's compares, order and counters are those of `rerank()`.

The llama branch and `OpenAiListwiseLlmRanker` (remote API) are other model families / services and stay out of scope.
"""
import copy
import os
from typing import List, Optional

from ._backend import T5Backend
from .rankers import LlmRanker, SearchResult
from .setwise import SetwiseLlmRanker

# `generate()` without arguments uses the model's default generation budget: 20 NEW tokens in transformers >= 5 (what the reference does
# on this image and what the fixtures pin: 21 ids per compare including the decoder start token); transformers 4.31 counted the start
# token inside max_length=20, i.e. 19 new tokens — B200RANK_LISTWISE_MAX_NEW=19 reproduces that pin.
DEFAULT_MAX_NEW_TOKENS = 20


def create_permutation_instruction_complete(query: str, docs: List[SearchResult]) -> str:
    """The completion-style RankGPT prompt of listwise.py:87-107: header, "[rank] passage" blocks (passage = first 300 whitespace
    words, the 'Title: Content: ' marker dropped), then the instruction tail. (The reference glues the query and 'I will rank'
    together without a separator; kept, because it changes the tokens.)"""
    n = len(docs)
    parts = ["This is RankGPT, an intelligent assistant that can rank passages based on their relevancy to the query.\n\n"
             f"The following are {n} passages, each indicated by number identifier []. "
             f"I can rank them based on their relevance to query: {query}\n\n"]
    for rank, doc in enumerate(docs, start=1):
        words = doc.text.replace('Title: Content: ', '').strip().split()[:300]
        parts.append(f"[{rank}] {' '.join(words)}\n\n")
    parts.append(f"The search query is: {query}"
                 f"I will rank the {n} passages above based on their relevance to the search query. The passages "
                 "will be listed in descending order using identifiers, and the most relevant passages should be listed "
                 "first, and the output format should be [] > [] > etc, e.g., [1] > [2] > etc.\n\n"
                 f"The ranking results of the {n} passages (only identifiers) is:")
    return "".join(parts)


def clean_response(response: str) -> str:
    """Every non-digit becomes a blank (listwise.py:110-118)."""
    return "".join(c if c.isdigit() else " " for c in response).strip()


def remove_duplicate(response: List[int]) -> List[int]:
    """First occurrences, in order (listwise.py:121-126)."""
    return list(dict.fromkeys(response))


def receive_permutation(ranking: List[SearchResult], permutation: str, rank_start: int = 0, rank_end: int = 100) -> List[SearchResult]:
    """Apply a "[2] > [1] > ..." response to ranking[rank_start:rank_end] in place (listwise.py:129-144): 1-based identifiers, duplicates
    and out-of-window numbers dropped, unmentioned documents appended in their current order."""
    window = copy.deepcopy(ranking[rank_start:rank_end])
    picked = [i for i in remove_duplicate([int(tok) - 1 for tok in clean_response(permutation).split()]) if 0 <= i < len(window)]
    seen = set(picked)
    picked.extend(i for i in range(len(window)) if i not in seen)
    for offset, src in enumerate(picked):
        ranking[rank_start + offset] = window[src]
    return ranking


def _window_positions(n: int, window_size: int, step_size: int):
    """(start, end) of every window of one pass, bottom of the list first (listwise.py:183-190)."""
    end = n
    start = end - window_size
    while start >= 0:
        yield start, end
        end -= step_size
        start -= step_size


class ListwiseLlmRanker(LlmRanker):
    GREEDY_CHUNK = 32   # new tokens per b200rank_greedy call (the engine takes up to 64 decoder positions)
    CHARACTERS = ["A", "B", "C", "D", "E", "F", "G", "H", "I", "J", "K", "L",
                  "M", "N", "O", "P", "Q", "R", "S", "T", "U", "V", "W"]

    def __init__(self, model_name_or_path, tokenizer_name_or_path, device, window_size, step_size,
                 scoring='generation', num_repeat=1, cache_dir=None, *, backend: Optional[T5Backend] = None):
        self.scoring = scoring
        self.device = device
        self.window_size = window_size
        self.step_size = step_size
        self.num_repeat = num_repeat
        self.backend = backend or T5Backend.load(model_name_or_path, tokenizer_name_or_path, device, cache_dir)
        self.tokenizer = self.backend.tokenizer
        self.llm = self.backend.engine
        self.config = self.backend.cfg
        self.decoder_input_ids = self.tokenizer.encode("<pad> Passage", add_special_tokens=False)
        self.target_token_ids = [self.tokenizer.encode(f"<pad> Passage {c}", add_special_tokens=False)[-1] for c in self.CHARACTERS]
        self.total_compare = 0
        self.total_prompt_tokens = 0
        self.total_completion_tokens = 0

    # ------------------------------------------------------------------ one window
    def _likelihood_rows(self, query: str, doc_sets: List[List]) -> List[List[int]]:
        rows = []
        for docs in doc_sets:
            if len(docs) > len(self.CHARACTERS):
                raise IndexError("list index out of range")   # listwise.py:274 runs out of labels the same way
            fields = {"query": query}
            fields.update({f"d{j}": d.text for j, d in enumerate(docs)})
            rows.extend(self.backend.prompt_rows(SetwiseLlmRanker._template(len(docs), self.CHARACTERS[:len(docs)]), [fields]))
        return rows

    def _generation_row(self, query: str, docs: List) -> List[int]:
        text = create_permutation_instruction_complete(query, docs)
        return list(self.tokenizer(text, truncation=True)["input_ids"])

    @staticmethod
    def _permutation_from_probs(probs, n_docs: int) -> str:
        ranked = sorted(zip([f"[{i + 1}]" for i in range(n_docs)], probs[:n_docs]), key=lambda x: x[1], reverse=True)
        return '>'.join(r[0] for r in ranked)

    def _generate_free(self, row: List[int]):
        """Greedy `generate(input_ids)` with the default budget: decoder starts from the pad token, stops at </s>. Returns the HF-shaped
        id vector (start token + new tokens)."""
        budget = int(os.environ.get("B200RANK_LISTWISE_MAX_NEW", DEFAULT_MAX_NEW_TOKENS))
        out = [self.backend.pad_id]
        while len(out) - 1 < budget:
            # one engine call per GREEDY_CHUNK new tokens: b200rank_greedy keeps a self-attention K/V cache inside a call (one decoder
            # position per step); a further chunk re-encodes the prompt and re-runs what has been generated as its prefix
            chunk = min(self.GREEDY_CHUNK, budget - (len(out) - 1))
            got = self.backend.generate_rows([row], out, chunk)[0].tolist()
            new = got[len(out):]
            out = got
            if not new or new[-1] == self.backend.eos_id:
                break
        return out

    def compare(self, query: str, docs: List) -> str:
        self.total_compare += 1
        if self.scoring == 'generation':
            row = self._generation_row(query, docs)
            self.total_prompt_tokens += len(row)
            out = self._generate_free(row)
            self.total_completion_tokens += len(out)
            return self.tokenizer.decode(out, skip_special_tokens=True).strip()
        if self.scoring == 'likelihood':
            row = self._likelihood_rows(query, [docs])[0]
            self.total_prompt_tokens += len(row)
            probs = self.backend.label_probs([row], self.decoder_input_ids, self.target_token_ids[:len(docs)])[0]
            return self._permutation_from_probs(probs, len(docs))
        raise NotImplementedError(f"scoring={self.scoring!r}")

    # ------------------------------------------------------------------ the sliding window
    def _window_chain(self, ranking: List[SearchResult]):
        """Generator over the windows of all passes: yields (start, end) with `state[0]` holding the current list, expects the
        response string via send(); returns nothing — the final list is state[0]."""
        state = [ranking]
        for _ in range(self.num_repeat):
            state[0] = copy.deepcopy(state[0])
            for start, end in _window_positions(len(state[0]), self.window_size, self.step_size):
                response = yield state, start, end
                state[0] = receive_permutation(state[0], response, start, end)
        return state

    def rerank(self, query: str, ranking: List[SearchResult]) -> List[SearchResult]:
        self.total_compare = 0
        self.total_prompt_tokens = 0
        self.total_completion_tokens = 0
        if self.num_repeat < 1:          # listwise.py:181 never rebinds the list: the caller's objects get the scores
            final = ranking
        else:
            chain = self._window_chain(ranking)
            final = None
            try:
                state, start, end = next(chain)
                while True:
                    state, start, end = chain.send(self.compare(query, state[0][start:end]))
            except StopIteration as stop:
                final = stop.value[0] if stop.value is not None else None
            if final is None:            # no window fits (fewer documents than window_size): the deep copy, untouched
                final = copy.deepcopy(ranking)
        for i, doc in enumerate(final):
            doc.score = -i
        return final

    def rerank_many(self, requests, window: int = 8):
        """Extension (not in the reference): an iterable of (query, ranking) pairs with up to `window` queries advancing in lockstep;
        every round scores the pending window of each active query in ONE engine call (likelihood mode). Results, per-query counters
        (readable after each yield) and order of the yields are those of successive rerank() calls. Generation mode decodes free-form
        text per window and simply loops over rerank()."""
        if self.scoring != 'likelihood' or self.num_repeat < 1:
            for query, ranking in requests:
                yield self.rerank(query, ranking)
            return
        it = iter(requests)
        active, done, next_out, seq = [], {}, 0, 0

        def finish(st, final):
            for i, doc in enumerate(final):
                doc.score = -i
            done[st["seq"]] = (final, st["counters"])

        def admit():
            nonlocal seq
            while len(active) < max(1, window):
                try:
                    query, ranking = next(it)
                except StopIteration:
                    return
                st = dict(seq=seq, query=query, counters=[0, 0, 0], chain=self._window_chain(ranking))
                seq += 1
                try:
                    st["pending"] = next(st["chain"])
                    active.append(st)
                except StopIteration as stop:
                    finish(st, stop.value[0] if stop.value is not None else copy.deepcopy(ranking))

        admit()
        while active or next_out in done:
            while next_out in done:
                result, c = done.pop(next_out)
                self.total_compare, self.total_prompt_tokens, self.total_completion_tokens = c
                next_out += 1
                yield result
            if not active:
                break
            sets = [st["pending"][0][0][st["pending"][1]:st["pending"][2]] for st in active]
            rows = []
            for st, docs in zip(active, sets):
                rows.extend(self._likelihood_rows(st["query"], [docs]))
            width = max(len(docs) for docs in sets)
            probs = self.backend.label_probs(rows, self.decoder_input_ids, self.target_token_ids[:width])
            still = []
            for st, docs, row, p in zip(active, sets, rows, probs):
                st["counters"][0] += 1
                st["counters"][1] += len(row)
                try:
                    st["pending"] = st["chain"].send(self._permutation_from_probs(p, len(docs)))
                    still.append(st)
                except StopIteration as stop:
                    finish(st, stop.value[0])
            active[:] = still
            admit()

    def truncate(self, text, length):
        return self.tokenizer.convert_tokens_to_string(self.tokenizer.tokenize(text)[:length])


class OpenAiListwiseLlmRanker(LlmRanker):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("OpenAI-backed rankers are a remote API, outside the B200 engine's scope (SURVEY.md §2.1)")
