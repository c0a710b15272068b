"""Setwise ranker on the B200 engine — drop-in for the T5 branch of the reference's llmrankers/setwise.py.

compare() keeps the reference's prompt, decoder prefix `<pad> Passage`, label extraction (last character of the decoded
generation, or arg-max label probability under scoring='likelihood'), fallbacks and counters (setwise.py:79-198); the
sort drivers reproduce the reference's compare sequence (setwise.py:200-313) via llmrankers/_sorting.py.
The llama / OpenAI / Rank-R1(vLLM) variants of the reference are different model families and out of scope.
"""
import copy
import os
import random
from collections import Counter
from typing import List, Optional

import numpy as np

from ._backend import T5Backend
from ._sorting import heap_top_k, heap_top_k_batched, heap_top_k_rounds, setwise_bubble_rounds, setwise_bubble_top_k
from .rankers import LlmRanker, SearchResult

random.seed(929)  # setwise.py:18


class SetwiseLlmRanker(LlmRanker):
    CHARACTERS = ["A", "B", "C", "D", "E", "F", "G", "H", "I", "J", "K", "L",
                  "M", "N", "O", "P", "Q", "R", "S", "T", "U", "V", "W"]

    def __init__(self, model_name_or_path, tokenizer_name_or_path, device, num_child=3, k=10, scoring='generation',
                 method="heapsort", num_permutation=1, cache_dir=None, *, backend: Optional[T5Backend] = None):
        self.device = device
        self.num_child = num_child
        self.num_permutation = num_permutation
        self.k = k
        self.backend = backend or T5Backend.load(model_name_or_path, tokenizer_name_or_path, device, cache_dir)
        self.tokenizer = self.backend.tokenizer
        self.llm = self.backend.engine
        self.config = self.backend.cfg
        # setwise.py:51-59 — the reference's batch_encode_plus call no longer exists in transformers 5; same ids
        self.decoder_input_ids = self.tokenizer.encode("<pad> Passage", add_special_tokens=False)
        self.target_token_ids = [self.tokenizer.encode(f"<pad> Passage {c}", add_special_tokens=False)[-1] for c in self.CHARACTERS]
        self.scoring = scoring
        self.method = method
        self.total_compare = 0
        self.total_completion_tokens = 0
        self.total_prompt_tokens = 0

    @staticmethod
    def _prompt(query: str, texts: List[str], labels: List[str]) -> str:
        passages = "\n\n".join(f'Passage {labels[i]}: "{t}"' for i, t in enumerate(texts))
        return (f'Given a query "{query}", which of the following passages is the most relevant one to the query?\n\n'
                + passages + '\n\nOutput only the passage label of the most relevant passage:')

    @staticmethod
    def _template(n_docs: int, labels: List[str]) -> str:
        """_prompt as a format template with fields {query}, {d0} .. {d<n-1>} for the token-level assembler (braces in the text
        itself never pass through str.format this way)."""
        passages = "\n\n".join('Passage %s: "{d%d}"' % (labels[i], i) for i in range(n_docs))
        return ('Given a query "{query}", which of the following passages is the most relevant one to the query?\n\n'
                + passages + '\n\nOutput only the passage label of the most relevant passage:')

    def _rows(self, query: str, doc_sets: List[List], label_sets: Optional[List[List[str]]] = None) -> List[List[int]]:
        """Token rows of `_prompt(query, texts, labels)` for every document set: quoted passages and the quoted query are cached
        units, so a passage is tokenised once per rerank however many compares it takes part in."""
        rows = []
        for i, docs in enumerate(doc_sets):
            labels = label_sets[i] if label_sets is not None else self.CHARACTERS
            fields = {"query": query}
            fields.update({f"d{j}": d.text for j, d in enumerate(docs)})
            rows.extend(self.backend.prompt_rows(self._template(len(docs), list(labels[:len(docs)])), [fields]))
        return rows

    def compare(self, query: str, docs: List):
        self.total_compare += 1 if self.num_permutation == 1 else self.num_permutation
        if self.scoring == 'generation':
            if self.num_permutation == 1:
                row = self._rows(query, [docs])[0]
                self.total_prompt_tokens += len(row)
                out = self.backend.generate(np.asarray([row], np.int32), self.decoder_input_ids, 2)[0]
                self.total_completion_tokens += int(out.shape[0])
                output = self.tokenizer.decode(out.tolist(), skip_special_tokens=True).strip()
                output = output[-1]
            else:
                output = self._compare_permutations(query, docs)
        elif self.scoring == 'likelihood':
            row = self._rows(query, [docs])[0]
            self.total_prompt_tokens += len(row)
            probs = self.backend.label_probs([row], self.decoder_input_ids, self.target_token_ids[:len(docs)])[0]
            ranked = sorted(zip(self.CHARACTERS[:len(docs)], probs), key=lambda x: x[1], reverse=True)
            output = ranked[0][0]
        else:
            raise NotImplementedError
        if not (len(output) == 1 and output in self.CHARACTERS):
            print(f"Unexpected output: {output}")
        return output

    def _compare_permutations(self, query: str, docs: List) -> str:
        """setwise.py:102-157 — vote over num_permutation shuffles of passages and labels (module RNG, seed 929)."""
        id_passage = list(enumerate(docs))
        labels = [self.CHARACTERS[i] for i in range(len(docs))]
        refs, prompts = [], []
        for _ in range(self.num_permutation):
            perm = random.sample(id_passage, len(id_passage))
            chars = random.sample(labels, len(labels))
            refs.append(([p[0] for p in perm], chars))
            prompts.append(([p[1] for p in perm], chars))
        rows = self._rows(query, [p[0] for p in prompts], [p[1] for p in prompts])
        ids, _ = self.backend.pad_rows(rows, self.backend.pad_id)
        self.total_prompt_tokens += ids.shape[1] * ids.shape[0]
        out = self.backend.generate(ids, self.decoder_input_ids, 2)
        texts = self.tokenizer.batch_decode(out[:, len(self.decoder_input_ids):].tolist(), skip_special_tokens=True)
        candidates = []
        for (docids, chars), result in zip(refs, texts):
            result = result.strip().upper()
            if len(result) != 1 or result not in chars:
                print(f"Unexpected output: {result}")
                continue
            candidates.append(docids[chars.index(result)])
        if not candidates:
            print(f"Unexpected voting: {texts}")
            return "Unexpected voting."
        counts = Counter(candidates)
        top = max(counts.values())
        best = [c for c, n in counts.items() if n == top]
        return self.CHARACTERS[best[0] if len(best) == 1 else random.choice(best)]

    def _compare_items(self, items: List) -> List:
        """Labels for several independent (query, docs) compare sets in ONE engine call (num_permutation == 1). Every prompt is a
        batch of one in the reference (no padding, setwise.py:90), the engine packs real tokens only and a row's result does not
        depend on its neighbours, so each label equals that of a sequential compare(). Returns [(label, prompt_tokens,
        completion_tokens)] and leaves the counters to the caller (the sets may belong to different queries)."""
        rows = []
        for query, docs in items:
            rows.extend(self._rows(query, [docs]))
        out = []
        if self.scoring == 'generation':
            for row, gen in zip(rows, self.backend.generate_rows(rows, self.decoder_input_ids, 2)):
                out.append((self.tokenizer.decode(gen.tolist(), skip_special_tokens=True).strip()[-1], len(row), int(gen.shape[0])))
        elif self.scoring == 'likelihood':
            width = max(len(docs) for _, docs in items)
            probs = self.backend.label_probs(rows, self.decoder_input_ids, self.target_token_ids[:width])
            for row, (_, docs), p in zip(rows, items, probs):
                ranked = sorted(zip(self.CHARACTERS[:len(docs)], p[:len(docs)]), key=lambda x: x[1], reverse=True)
                out.append((ranked[0][0], len(row), 0))
        else:
            raise NotImplementedError
        for label, _, _ in out:
            if not (len(label) == 1 and label in self.CHARACTERS):
                print(f"Unexpected output: {label}")
        return out

    def _compare_many(self, query: str, doc_sets: List[List]) -> List[str]:
        """`compare()` for several independent document sets of one query; counters as len(doc_sets) sequential compare() calls."""
        res = self._compare_items([(query, docs) for docs in doc_sets])
        self.total_compare += len(res)
        self.total_prompt_tokens += sum(r[1] for r in res)
        self.total_completion_tokens += sum(r[2] for r in res)
        return [r[0] for r in res]

    def _pick(self, label: str, inds: List[int]) -> int:
        """Label -> arr index: an unknown label keeps the parent, and so does a label beyond the compared set (setwise.py:206-213)."""
        b = self.CHARACTERS.index(label) if label in self.CHARACTERS else 0
        return inds[b] if b < len(inds) else inds[0]

    def rerank_many(self, requests, window: int = 8):
        """Extension (not in the reference): rerank an iterable of (query, ranking) pairs (heapsort or bubblesort) with up to `window`
        queries advancing in lockstep — every round sends the pending compares of all active queries (several per query while a heap level
        is being built) to the GPU as one batch. A compare is latency-bound by its two ~250-kernel decoder passes, so batching
        across queries multiplies throughput; each query's compares, order, scores and counters are exactly those of rerank().
        Yields rerank()'s result for each pair in request order; after each yield the three counters hold that query's totals.
        Other methods / num_permutation > 1 fall back to rerank()."""
        if self.method not in ("heapsort", "bubblesort") or self.num_permutation != 1:
            for query, ranking in requests:
                yield self.rerank(query, ranking)
            return
        heap = self.method == "heapsort"
        it = iter(requests)
        active, done, next_out, seq = [], {}, 0, 0
        template1 = self._template(1, self.CHARACTERS)

        def admit():
            nonlocal seq
            while len(active) < max(1, window):
                try:
                    query, ranking = next(it)
                except StopIteration:
                    return
                ranking = list(ranking)
                a = self.backend.assembler(template1)
                if a is not None:
                    a.warm("d0", [d.text for d in ranking])
                st = dict(seq=seq, query=query, original=copy.deepcopy(ranking), arr=ranking, counters=[0, 0, 0])
                st["gen"] = (heap_top_k_rounds(st["arr"], self.num_child, self.k) if heap
                             else setwise_bubble_rounds(st["arr"], self.num_child, self.k))
                try:
                    st["round"] = next(st["gen"])
                    active.append(st)
                except StopIteration:       # nothing to compare (fewer than two documents)
                    st["round"] = None
                    finish(st)
                seq += 1

        def finish(st):
            done[st["seq"]] = (_assemble(list(reversed(st["arr"])) if heap else st["arr"], st["original"], self.k), st["counters"])

        admit()
        while active or next_out in done:
            while next_out in done:
                result, c = done.pop(next_out)
                self.total_compare, self.total_prompt_tokens, self.total_completion_tokens = c
                next_out += 1
                yield result
            if not active:
                break
            # a heap round holds (docs, arr indices) requests, a bubble round one bare window
            items = [(st["query"], req[0] if heap else req) for st in active for req in st["round"]]
            res = self._compare_items(items)
            pos, still = 0, []
            for st in active:
                n = len(st["round"])
                mine = res[pos:pos + n]
                pos += n
                st["counters"][0] += n
                st["counters"][1] += sum(r[1] for r in mine)
                st["counters"][2] += sum(r[2] for r in mine)
                if heap:
                    picks = [self._pick(r[0], inds) for r, (_, inds) in zip(mine, st["round"])]
                else:
                    # bubblesort: label -> offset from the window's head. An unknown label keeps the head; a label past the window is
                    # applied as the reference applies it (setwise.py:252-259 indexes the whole ranking: it swaps with the element
                    # that far down the list, and raises IndexError only past the end of the list — the generator does the same)
                    picks = [self.CHARACTERS.index(r[0]) if r[0] in self.CHARACTERS else 0 for r in mine]
                try:
                    st["round"] = st["gen"].send(picks)
                    still.append(st)
                except StopIteration:
                    finish(st)
            active[:] = still
            admit()

    def _best_index(self, query: str, docs: List) -> int:
        """Label -> position in the compared set; an unknown label keeps the head (setwise.py:206-209, 252-255)."""
        try:
            return self.CHARACTERS.index(self.compare(query, docs))
        except ValueError:
            return 0

    def rerank(self, query: str, ranking: List[SearchResult]) -> List[SearchResult]:
        original_ranking = copy.deepcopy(ranking)
        self.total_compare = 0
        self.total_completion_tokens = 0
        self.total_prompt_tokens = 0
        a = self.backend.assembler(self._template(1, self.CHARACTERS))
        if a is not None:
            a.warm("d0", [d.text for d in ranking])   # every candidate passage tokenised once, in one batched call
        if self.method == "heapsort":
            def pick(docs, inds):
                b = self._best_index(query, docs)
                return inds[b] if b < len(inds) else inds[0]  # a label beyond the set keeps the parent (setwise.py:210-213)
            if self.num_permutation == 1 and os.environ.get("B200RANK_BATCHED_SORT", "1") != "0":
                # level-parallel heap construction: the compares of independent subtrees go to the GPU as one batch (_sorting.py)
                def pick_many(requests):
                    labels = self._compare_many(query, [docs for docs, _ in requests])
                    return [self._pick(lab, inds) for (_, inds), lab in zip(requests, labels)]
                heap_top_k_batched(ranking, self.num_child, self.k, pick_many)
            else:
                heap_top_k(ranking, self.num_child, self.k, pick)
            ranking = list(reversed(ranking))
        elif self.method == "bubblesort":
            setwise_bubble_top_k(ranking, self.num_child, self.k, lambda window: self._best_index(query, window))
        else:
            raise NotImplementedError(f'Method {self.method} is not implemented.')
        return _assemble(ranking, original_ranking, self.k)

    def truncate(self, text, length):
        return self.tokenizer.convert_tokens_to_string(self.tokenizer.tokenize(text)[:length])


def _assemble(ranking: List[SearchResult], original: List[SearchResult], k: int) -> List[SearchResult]:
    """setwise.py:300-311 / pairwise.py:279-290: top-k get score -rank; the rest follow in their ORIGINAL order."""
    results, top, rank = [], set(), 1
    for doc in ranking[:k]:
        top.add(doc.docid)
        results.append(SearchResult(docid=doc.docid, score=-rank, text=None))
        rank += 1
    for doc in original:
        if doc.docid not in top:
            results.append(SearchResult(docid=doc.docid, score=-rank, text=None))
            rank += 1
    return results


class OpenAiSetwiseLlmRanker(SetwiseLlmRanker):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("OpenAI-backed rankers are a remote API, outside the B200 engine's scope (SURVEY.md §2.1)")


class RankR1SetwiseLlmRanker(SetwiseLlmRanker):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("Rank-R1 (vLLM decoder-only models) is outside the B200 engine's scope (SURVEY.md §2.1)")
