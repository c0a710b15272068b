"""Token-level prompt assembly with a per-field token cache.

The reference renders one string per (query, document) and tokenises all of them (`Text2TextGenerationDataset`,
pairwise.py:17-26): the query and the template are re-tokenised once per document, and a document that appears in several
prompts (or that `run.py` already tokenised for `truncate`) is tokenised again. At >= 7 k documents/s/GPU that host work is
the bottleneck (32 ms of tokenizer time per 100-document query against 13 ms on the GPU).

T5's tokenizer (SentencePiece Unigram, `split_by_whitespace`; the `tokenizers` port pre-tokenises on whitespace / Metaspace)
segments every whitespace-delimited word independently, so a prompt whose `{fields}` are bounded by whitespace tokenises to
the concatenation of the tokenisations of its pieces: literal | field | literal | ... | </s>. `PromptAssembler` splits a
template once, tokenises the literals once, caches field values (LRU by string) and concatenates ids. It does NOT trust that
argument blindly: the first `verify` prompts of every assembler are checked against the tokenizer on the whole string, and on
any mismatch the assembler permanently falls back to whole-string tokenisation. Punctuation glued to a field (the quotes of the
setwise / pairwise prompts, `Passage A: "{text}"`) becomes part of the cached unit, so those prompts assemble too — a passage that
appears in 198 pair prompts of an allpair rerank is tokenised once.
"""
import threading
from collections import OrderedDict
from string import Formatter
from typing import Dict, List, Sequence


class TokenCache:
    """LRU text -> token ids, shareable between assemblers (one per tokenizer), safe to use from several threads."""

    def __init__(self, size: int = 200_000):
        self.data: "OrderedDict[str, List[int]]" = OrderedDict()
        self.size = size
        self.lock = threading.Lock()
        self.hits = self.misses = 0


class PromptAssembler:
    def __init__(self, tokenizer, template: str, cache_size: int = 200_000, verify: int = 16, cache: TokenCache = None):
        self.tokenizer = tokenizer
        self.template = template
        self.tc = cache if cache is not None else TokenCache(cache_size)
        self.verify_left = verify
        self.eos = tokenizer.eos_token_id
        # the Rust tokenizer behind a fast tokenizer: encode_batch releases the GIL and does not touch the truncation / padding
        # state that transformers' __call__ mutates (which is what makes concurrent __call__s raise "Already borrowed")
        self.raw = getattr(tokenizer, "backend_tokenizer", None) if getattr(tokenizer, "is_fast", False) else None
        self.eligible = True
        # Split the template into literals and fields. Punctuation glued to a field (the quotes of `Passage A: "{text}"`) moves
        # out of the literals into the field's unit: the cached string is glue + value + glue, which IS bounded by whitespace.
        pieces = [(lit, field) for lit, field, spec, conv in Formatter().parse(template)]
        if any(spec or conv for _, _, spec, conv in Formatter().parse(template)):
            self.eligible = False
        lits = [lit for lit, _ in pieces] + [""]            # literal before every field, plus the tail after the last one
        fields = [f for _, f in pieces]
        if fields and fields[-1] is None:                   # trailing literal came as (lit, None)
            lits[-1] = lits[-2]
            lits.pop(-2)
            fields.pop()
        self.units = []                                     # [(glue_before, field, glue_after)]
        for i, f in enumerate(fields):
            before, after = lits[i], lits[i + 1]
            nb = _trailing_nonspace(before)
            na = _leading_nonspace(after)
            if i + 1 < len(fields) and na == len(after) and after != "":
                self.eligible = False                       # two fields joined by punctuation only: not separable
            gb, ga = before[len(before) - nb:], after[:na]
            lits[i], lits[i + 1] = before[:len(before) - nb], after[na:]
            self.units.append((gb, f, ga))
        self.lit_ids = [self._encode(l) if l.strip() else [] for l in lits]
        self.hits = self.misses = 0

    # ------------------------------------------------------------------ tokenizer access
    def _encode(self, text: str) -> List[int]:
        return self.tokenizer.encode(text, add_special_tokens=False)

    def _encode_batch(self, texts: List[str], specials: bool) -> List[List[int]]:
        if self.raw is not None:
            return [e.ids for e in self.raw.encode_batch(texts, add_special_tokens=specials)]
        return self.tokenizer(texts, add_special_tokens=specials)["input_ids"]

    def _unit_ids(self, texts: Sequence[str]) -> List[List[int]]:
        """Token ids of every unit string, through the shared LRU cache; misses are tokenised in ONE batched tokenizer call."""
        tc = self.tc
        out: List = [None] * len(texts)
        todo: Dict[str, List[int]] = {}
        with tc.lock:
            for i, v in enumerate(texts):
                ids = tc.data.get(v)
                if ids is None:
                    todo.setdefault(v, []).append(i)
                else:
                    tc.data.move_to_end(v)
                    out[i] = ids
                    tc.hits += 1
                    self.hits += 1
        if todo:
            keys = list(todo)
            enc = self._encode_batch(keys, False)
            with tc.lock:
                for t, ids in zip(keys, enc):
                    tc.misses += len(todo[t])
                    self.misses += len(todo[t])
                    tc.data[t] = ids
                    for i in todo[t]:
                        out[i] = ids
                while len(tc.data) > tc.size:
                    tc.data.popitem(last=False)
        return out

    def warm(self, field: str, values: Sequence[str]) -> None:
        """Tokenise the units of `field` for all `values` in one batched call (e.g. every candidate passage at the start of a
        sort-based rerank, whose compares then only concatenate cached ids)."""
        if not self.eligible:
            return
        for gb, f, ga in self.units:
            if f == field:
                self._unit_ids([gb + v + ga for v in values])

    # ------------------------------------------------------------------ rows
    def whole_string(self, rows_fields: Sequence[Dict[str, str]]) -> List[List[int]]:
        prompts = [self.template.format(**f) for f in rows_fields]
        return self._encode_batch(prompts, True) if prompts else []

    def rows(self, rows_fields: Sequence[Dict[str, str]]) -> List[List[int]]:
        """Token-id rows (</s> appended) for one prompt per dict of field values — identical to tokenising the rendered strings."""
        rows_fields = list(rows_fields)
        if not rows_fields:
            return []
        if not self.eligible:
            return self.whole_string(rows_fields)
        n_u = len(self.units)
        flat = self._unit_ids([gb + rf[f] + ga for rf in rows_fields for gb, f, ga in self.units])
        rows = []
        for r in range(len(rows_fields)):
            ids: List[int] = []
            for u in range(n_u):
                ids += self.lit_ids[u]
                ids += flat[r * n_u + u]
            ids += self.lit_ids[n_u]
            ids.append(self.eos)
            rows.append(ids)
        if self.verify_left > 0:
            n = min(self.verify_left, len(rows))
            want = self.whole_string(rows_fields[:n])
            self.verify_left -= n
            if want != rows[:n]:
                self.eligible = False     # this vocabulary / normaliser does not segment per word: never assemble again
                return self.whole_string(rows_fields)
        return rows


def _trailing_nonspace(s: str) -> int:
    n = 0
    while n < len(s) and not s[len(s) - 1 - n].isspace():
        n += 1
    return n


def _leading_nonspace(s: str) -> int:
    n = 0
    while n < len(s) and not s[n].isspace():
        n += 1
    return n
