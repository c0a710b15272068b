"""Comparison-driven top-k selection used by the setwise and pairwise rankers, written once against a callback.

The compare *sequence* is what has to match the reference (each call is one LLM forward, counted in total_compare), so
the index arithmetic follows the reference's algorithms exactly:
  * c-ary max-heap (setwise.py:200-232; pairwise.py:133-162 is the c = 2 case with a boolean comparison),
  * the windowed bubble passes of setwise.py:243-273 and pairwise.py:248-273.
"""
from typing import Callable, List, Sequence


def heap_top_k(arr: List, num_child: int, k: int, pick_best: Callable[[List, List[int]], int]) -> None:
    """In-place: after the call arr[-1], arr[-2], ... hold the k best items (best last), as in setwise.py:219-232.
    pick_best(docs, inds) returns the arr-index of the preferred element among inds (inds[0] is the parent)."""
    n = len(arr)

    def sift(limit: int, i: int) -> None:
        while num_child * i + 1 < limit:
            lo, hi = num_child * i + 1, min(num_child * (i + 1) + 1, limit)
            inds = [i] + list(range(lo, hi))
            largest = pick_best([arr[j] for j in inds], inds)
            if largest == i:
                return
            arr[i], arr[largest] = arr[largest], arr[i]
            i = largest

    for i in range(n // num_child, -1, -1):
        sift(n, i)
    ranked = 0
    for i in range(n - 1, 0, -1):
        arr[i], arr[0] = arr[0], arr[i]
        ranked += 1
        if ranked == k:
            break
        sift(i, 0)


def _rounds(gens):
    """Sub-generator: advance request-yielding generators in lockstep. Yields one ROUND (the list of the pending request of every
    live generator) at a time and is sent the list of answers; with a single generator this is the sequential algorithm."""
    pending = []
    for g in gens:
        try:
            pending.append((g, next(g)))
        except StopIteration:
            pass
    while pending:
        answers = yield [req for _, req in pending]
        nxt = []
        for (g, _), a in zip(pending, answers):
            try:
                nxt.append((g, g.send(a)))
            except StopIteration:
                pass
        pending = nxt


def drive_rounds(gen, answer_many: Callable[[List], List]) -> None:
    """Run a round generator to completion against a batch oracle."""
    try:
        reqs = next(gen)
        while True:
            reqs = gen.send(answer_many(reqs))
    except StopIteration:
        pass


def heap_top_k_rounds(arr: List, num_child: int, k: int):
    """heap_top_k as a generator of compare ROUNDS with level-parallel heap construction. The reference builds
    the heap with `for i in range(n // c, -1, -1): heapify(arr, n, i)` (setwise.py:219-223); heapify(i) only ever touches the
    subtree of i, and nodes of one tree level have disjoint subtrees, so the sift-downs of a level can advance in lockstep: round r
    holds the r-th compare of every still-moving sift of that level. The array after each level — hence the final order — and the
    multiset of compares (total_compare, prompt / completion token counters) are identical to the sequential build; only the
    interleaving of compares across independent subtrees changes. The k extractions are inherently sequential (rounds of one).
    Yields [(docs, inds), ...]; expects [arr-index of the preferred element among inds, ...]. Being a generator, several
    queries' sorts can be advanced together (SetwiseLlmRanker.rerank_many)."""
    n = len(arr)

    def sift(limit: int, i: int):
        while num_child * i + 1 < limit:
            lo, hi = num_child * i + 1, min(num_child * (i + 1) + 1, limit)
            inds = [i] + list(range(lo, hi))
            largest = yield ([arr[j] for j in inds], inds)
            if largest == i:
                return
            arr[i], arr[largest] = arr[largest], arr[i]
            i = largest

    # tree levels: level L holds indices [first(L), first(L+1)), first(L+1) = first(L) * c + 1
    levels, first = [], 0
    while first <= n // num_child:
        nxt = first * num_child + 1
        levels.append(range(first, min(nxt, n // num_child + 1)))
        first = nxt
    for level in reversed(levels):
        yield from _rounds([sift(n, i) for i in reversed(level)])
    ranked = 0
    for i in range(n - 1, 0, -1):
        arr[i], arr[0] = arr[0], arr[i]
        ranked += 1
        if ranked == k:
            break
        yield from _rounds([sift(i, 0)])


def heap_top_k_batched(arr: List, num_child: int, k: int, pick_best_many: Callable[[List], List[int]]) -> None:
    """In place, like heap_top_k, with the compares of every round resolved by ONE pick_best_many call."""
    drive_rounds(heap_top_k_rounds(arr, num_child, k), pick_best_many)


def binary_heap_top_k(arr: List, k: int, greater: Callable[[object, object], bool]) -> None:
    """pairwise.py:133-162: binary max-heap where `greater(a, b)` costs one LLM compare; left child is tested first and
    the right child is compared against the current largest."""
    n = len(arr)

    def sift(limit: int, i: int) -> None:
        while True:
            largest, l, r = i, 2 * i + 1, 2 * i + 2
            if l < limit and greater(arr[l], arr[i]):
                largest = l
            if r < limit and greater(arr[r], arr[largest]):
                largest = r
            if largest == i:
                return
            arr[i], arr[largest] = arr[largest], arr[i]
            i = largest

    for i in range(n // 2, -1, -1):
        sift(n, i)
    ranked = 0
    for i in range(n - 1, 0, -1):
        arr[i], arr[0] = arr[0], arr[i]
        ranked += 1
        if ranked == k:
            break
        sift(i, 0)


def binary_heap_top_k_rounds(arr: List, k: int):
    """binary_heap_top_k as a generator of compare rounds with level-parallel heap construction (see heap_top_k_rounds): yields
    [(a, b), ...] and expects [greater(a, b), ...]. Same final array and the same multiset of compares as pairwise.py:133-162;
    the k extractions stay sequential (rounds of one)."""
    n = len(arr)

    def sift(limit: int, i: int):
        while True:
            largest, l, r = i, 2 * i + 1, 2 * i + 2
            if l < limit and (yield (arr[l], arr[i])):
                largest = l
            if r < limit and (yield (arr[r], arr[largest])):
                largest = r
            if largest == i:
                return
            arr[i], arr[largest] = arr[largest], arr[i]
            i = largest

    levels, first = [], 0
    while first <= n // 2:
        nxt = 2 * first + 1
        levels.append(range(first, min(nxt, n // 2 + 1)))
        first = nxt
    for level in reversed(levels):
        yield from _rounds([sift(n, i) for i in reversed(level)])
    ranked = 0
    for i in range(n - 1, 0, -1):
        arr[i], arr[0] = arr[0], arr[i]
        ranked += 1
        if ranked == k:
            break
        yield from _rounds([sift(i, 0)])


def binary_heap_top_k_batched(arr: List, k: int, greater_many: Callable[[List], List[bool]]) -> None:
    """In place, like binary_heap_top_k, with the compares of every round resolved by ONE greater_many call."""
    drive_rounds(binary_heap_top_k_rounds(arr, k), greater_many)


def setwise_bubble_rounds(ranking: List, num_child: int, k: int):
    """setwise_bubble_top_k as a generator of compare rounds (always one request: the passes are a dependent chain), so that the
    bubble sorts of several queries can advance together. Yields [window]; expects [index of the preferred doc in the window]."""
    width = num_child + 1
    last_start = len(ranking) - width
    for i in range(k):
        start, end = last_start, last_start + width
        changed = False
        while True:
            if start < i:
                start = i
            window = ranking[start:end]
            b = (yield [window])[0]
            if b != 0:
                ranking[start], ranking[start + b] = ranking[start + b], ranking[start]
                if not changed:
                    changed = True
                    if last_start != len(ranking) - width and b == len(window) - 1:
                        last_start += len(window) - 1
            if start == i:
                break
            if not changed:
                last_start -= num_child
            start -= num_child
            end -= num_child


def setwise_bubble_top_k(ranking: List, num_child: int, k: int, best_index: Callable[[Sequence], int]) -> None:
    """setwise.py:243-273: k passes of a (num_child+1)-wide window sliding from the tail to position i, with the
    'skip the unchanged tail' bookkeeping (last_start). best_index(window) -> index of the preferred doc in the window
    (0 keeps the head). An index beyond the window raises IndexError exactly like the reference does (:259)."""
    width = num_child + 1
    last_start = len(ranking) - width
    for i in range(k):
        start, end = last_start, last_start + width
        changed = False
        while True:
            if start < i:
                start = i
            window = ranking[start:end]
            b = best_index(window)
            if b != 0:
                ranking[start], ranking[start + b] = ranking[start + b], ranking[start]
                if not changed:
                    changed = True
                    if last_start != len(ranking) - width and b == len(window) - 1:
                        last_start += len(window) - 1
            if start == i:
                break
            if not changed:
                last_start -= num_child
            start -= num_child
            end -= num_child


def pairwise_bubble_top_k(ranking: List, k: int, first_wins: Callable[[object, object], bool]) -> None:
    """pairwise.py:248-273: adjacent swaps from the tail; first_wins(lower, upper) costs one compare."""
    k = min(k, len(ranking))
    last_end = len(ranking) - 1
    for i in range(k):
        cur = last_end
        changed = False
        while cur > i:
            if first_wins(ranking[cur], ranking[cur - 1]):
                ranking[cur - 1], ranking[cur] = ranking[cur], ranking[cur - 1]
                if not changed:
                    changed = True
                    if last_end != len(ranking) - 1:
                        last_end += 1
            if not changed:
                last_end -= 1
            cur -= 1


def pairwise_bubble_rounds(ranking: List, k: int):
    """pairwise_bubble_top_k as a generator of compare rounds (always one request: the passes are a dependent chain), so that the
    bubble sorts of several queries can advance together. Yields [(lower, upper)]; expects [first_wins(lower, upper)]."""
    k = min(k, len(ranking))
    last_end = len(ranking) - 1
    for i in range(k):
        cur = last_end
        changed = False
        while cur > i:
            if (yield [(ranking[cur], ranking[cur - 1])])[0]:
                ranking[cur - 1], ranking[cur] = ranking[cur], ranking[cur - 1]
                if not changed:
                    changed = True
                    if last_end != len(ranking) - 1:
                        last_end += 1
            if not changed:
                last_end -= 1
            cur -= 1
