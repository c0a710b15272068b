"""Token-level prompt assembly with a per-field token cache.

The reference renders one string per (query, document) and tokenises all of them (`Text2TextGenerationDataset`,
pairwise.py:17-26): the query and the template are re-tokenised once per document, and a document that appears in several
prompts (or that `run.py` already tokenised for `truncate`) is tokenised again. At >= 7 k documents/s/GPU that host work is
the bottleneck (32 ms of tokenizer time per 100-document query against 13 ms on the GPU).

T5's tokenizer (SentencePiece Unigram, `split_by_whitespace`; the `tokenizers` port pre-tokenises on whitespace / Metaspace)
segments every whitespace-delimited word independently, so a prompt whose `{fields}` are bounded by whitespace tokenises to
the concatenation of the tokenisations of its pieces: literal | field | literal | ... | </s>. `PromptAssembler` splits a
template once, tokenises the literals once, caches field values (LRU by string) and concatenates ids. It does NOT trust that
argument blindly: the first `verify` prompts of every assembler (and every prompt whose field is not whitespace-bounded) are
checked against the tokenizer on the whole string, and on any mismatch the assembler permanently falls back to whole-string
tokenisation. Templates whose fields touch punctuation (the quoted passages of the setwise / pairwise prompts) are not eligible.
"""
import threading
from collections import OrderedDict
from string import Formatter
from typing import Dict, List, Sequence


class PromptAssembler:
    def __init__(self, tokenizer, template: str, cache_size: int = 200_000, verify: int = 16):
        self.tokenizer = tokenizer
        self.template = template
        self.cache: "OrderedDict[str, List[int]]" = OrderedDict()
        self.cache_size = cache_size
        self.verify_left = verify
        self.hits = self.misses = 0
        self.eos = tokenizer.eos_token_id
        self.lock = threading.Lock()   # rerank_many tokenises upcoming queries on worker threads
        # the Rust tokenizer behind a fast tokenizer: encode_batch releases the GIL and does not touch the truncation / padding
        # state that transformers' __call__ mutates (which is what makes concurrent __call__s raise "Already borrowed")
        self.raw = getattr(tokenizer, "backend_tokenizer", None) if getattr(tokenizer, "is_fast", False) else None
        self.parts = []      # [(literal ids, field name | None)]
        self.eligible = True
        pieces = list(Formatter().parse(template))
        for i, (lit, field, spec, conv) in enumerate(pieces):
            if spec or conv:
                self.eligible = False
            if field is not None:
                before_ok = (lit == "" and i == 0) or (lit != "" and lit[-1].isspace())
                nxt = pieces[i + 1][0] if i + 1 < len(pieces) else ""
                after_ok = (i + 1 == len(pieces)) or (nxt != "" and nxt[0].isspace())
                if not (before_ok and after_ok):
                    self.eligible = False   # a field glued to punctuation changes the segmentation of its first / last word
            self.parts.append((self._encode(lit) if lit else [], field))

    def _encode(self, text: str) -> List[int]:
        return self.tokenizer.encode(text, add_special_tokens=False)

    def _encode_batch(self, texts: List[str], specials: bool) -> List[List[int]]:
        if self.raw is not None:
            return [e.ids for e in self.raw.encode_batch(texts, add_special_tokens=specials)]
        return self.tokenizer(texts, add_special_tokens=specials)["input_ids"]

    def _field_ids(self, values: Sequence[str]) -> List[List[int]]:
        """Token ids of every value, through the LRU cache; misses are tokenised in ONE batched tokenizer call."""
        out: List = [None] * len(values)
        todo: Dict[str, List[int]] = {}
        with self.lock:
            for i, v in enumerate(values):
                ids = self.cache.get(v)
                if ids is None:
                    todo.setdefault(v, []).append(i)
                else:
                    self.cache.move_to_end(v)
                    out[i] = ids
                    self.hits += 1
        if todo:
            texts = list(todo)
            enc = self._encode_batch(texts, False)
            with self.lock:
                for t, ids in zip(texts, enc):
                    self.misses += len(todo[t])
                    self.cache[t] = ids
                    for i in todo[t]:
                        out[i] = ids
                while len(self.cache) > self.cache_size:
                    self.cache.popitem(last=False)
        return out

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
        names = [f for _, f in self.parts if f is not None]
        per_field = {n: self._field_ids([rf[n] for rf in rows_fields]) for n in set(names)}
        rows = []
        for r in range(len(rows_fields)):
            ids: List[int] = []
            for lit, field in self.parts:
                ids += lit
                if field is not None:
                    ids += per_field[field][r]
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
