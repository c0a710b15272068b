"""CLI with the reference's grammar (run.py:206-258): `python run.py run <run flags> {pointwise|pairwise|setwise} <flags>`,
same flags and defaults, same four summary prints and the same TREC output line (run.py:41-49, 198-201).

Differences, all additive: ir_datasets / pyserini are imported lazily (absent on an offline box), and two extra
`run` flags name plain-text sources so the CLI is usable without them:
    --queries_tsv  qid<TAB>text         --collection_tsv  docid<TAB>text
`--model_name_or_path synthetic:flan-t5-large` selects seeded random weights + the synthetic tokenizer.
"""
import argparse
import collections
import json
import os
import logging
import random
import sys
import time

from llmrankers.pairwise import DuoT5LlmRanker, OpenAiPairwiseLlmRanker, PairwiseLlmRanker
from llmrankers.pointwise import MonoT5LlmRanker, PointwiseLlmRanker
from llmrankers.rankers import SearchResult
from llmrankers.setwise import OpenAiSetwiseLlmRanker, SetwiseLlmRanker
from llmrankers.listwise import ListwiseLlmRanker, OpenAiListwiseLlmRanker

random.seed(929)
logger = logging.getLogger(__name__)


def parse_args(parser, commands, argv=None):
    """Split argv at sub-command names and parse each slice into its own namespace (run.py:20-38)."""
    argv = sys.argv[1:] if argv is None else argv
    groups = [[]]
    for tok in argv:
        if tok in commands.choices:
            groups.append([tok])
        else:
            groups[-1].append(tok)
    args = argparse.Namespace(**{c: None for c in commands.choices})
    parser.parse_args(groups[0], namespace=args)
    for g in groups[1:]:
        ns = argparse.Namespace()
        setattr(args, g[0], ns)
        parser.parse_args(g, namespace=ns)
    return args


def write_run_file(path, results, tag):
    with open(path, 'w') as f:
        for qid, _, ranking in results:
            for rank, doc in enumerate(ranking, start=1):
                f.write(f"{qid}\tQ0\t{doc.docid}\t{rank}\t{doc.score}\t{tag}\n")


def build_ranker(args):
    run = args.run
    common = dict(model_name_or_path=run.model_name_or_path, tokenizer_name_or_path=run.tokenizer_name_or_path,
                  device=run.device, cache_dir=run.cache_dir)
    if args.pointwise:
        cls = MonoT5LlmRanker if 'monot5' in run.model_name_or_path else PointwiseLlmRanker
        return cls(method=args.pointwise.method, batch_size=args.pointwise.batch_size, **common)
    if args.setwise:
        if run.openai_key:
            return OpenAiSetwiseLlmRanker(model_name_or_path=run.model_name_or_path, api_key=run.openai_key,
                                          num_child=args.setwise.num_child, method=args.setwise.method, k=args.setwise.k)
        return SetwiseLlmRanker(num_child=args.setwise.num_child, scoring=run.scoring, method=args.setwise.method,
                                num_permutation=args.setwise.num_permutation, k=args.setwise.k, **common)
    if args.pairwise:
        if args.pairwise.method != 'allpair':
            args.pairwise.batch_size = 2
            logger.info('Setting batch_size to 2.')
        if run.openai_key:
            return OpenAiPairwiseLlmRanker(model_name_or_path=run.model_name_or_path, api_key=run.openai_key,
                                           method=args.pairwise.method, k=args.pairwise.k)
        cls = DuoT5LlmRanker if 'duot5' in run.model_name_or_path else PairwiseLlmRanker
        return cls(method=args.pairwise.method, batch_size=args.pairwise.batch_size, k=args.pairwise.k, **common)
    if args.listwise:
        if run.openai_key:
            return OpenAiListwiseLlmRanker(model_name_or_path=run.model_name_or_path, api_key=run.openai_key,
                                           window_size=args.listwise.window_size, step_size=args.listwise.step_size,
                                           num_repeat=args.listwise.num_repeat)
        return ListwiseLlmRanker(window_size=args.listwise.window_size, step_size=args.listwise.step_size, scoring=run.scoring,
                                 num_repeat=args.listwise.num_repeat, **common)
    raise ValueError('Must specify either --pointwise, --setwise, --pairwise or --listwise.')


def _read_tsv(path):
    out = {}
    with open(path) as f:
        for line in f:
            key, _, text = line.rstrip("\n").partition("\t")
            out[key] = text
    return out


def load_sources(run, ranker):
    """query_map (truncated queries) and a docid -> text getter, from TSV files, ir_datasets or pyserini (run.py:135-149)."""
    query_map = {}
    if run.queries_tsv is not None or run.collection_tsv is not None:
        if run.queries_tsv is None or run.collection_tsv is None:
            raise ValueError('--queries_tsv and --collection_tsv must be given together.')
        for qid, text in _read_tsv(run.queries_tsv).items():
            query_map[qid] = ranker.truncate(text, run.query_length)
        docs = _read_tsv(run.collection_tsv)
        return query_map, docs.__getitem__
    if run.ir_dataset_name is not None:
        import ir_datasets
        dataset = ir_datasets.load(run.ir_dataset_name)
        for q in dataset.queries_iter():
            query_map[q.query_id] = ranker.truncate(q.text, run.query_length)
        store = dataset.docs_store()

        def get(docid):
            d = store.get(docid)
            return f'{d.title} {d.text}' if 'title' in dir(d) else d.text
        return query_map, get
    from pyserini.search._base import get_topics
    from pyserini.search.lucene import LuceneSearcher
    topics = get_topics(run.pyserini_index + '-test')
    for tid in list(topics.keys()):
        query_map[str(tid)] = ranker.truncate(topics[tid]['title'], run.query_length)
    searcher = LuceneSearcher.from_prebuilt_index(run.pyserini_index + '.flat')

    def get(docid):
        data = json.loads(searcher.doc(docid).raw())
        return f'{data["title"]} {data["text"]}' if 'title' in data else data['text']
    return query_map, get


def read_first_stage(run, ranker, query_map, get_text):
    """6-column TREC run -> [(qid, query, [SearchResult] capped at --hits)] in file order (run.py:151-176)."""
    rankings, cur_qid, cur = [], None, []
    with open(run.run_path) as f:
        for line in f:
            qid, _, docid, _, score, _ = line.strip().split()
            if qid != cur_qid:
                if cur_qid is not None:
                    rankings.append((cur_qid, query_map[cur_qid], cur[:run.hits]))
                cur, cur_qid = [], qid
            if len(cur) >= run.hits:
                continue
            cur.append(SearchResult(docid=docid, score=float(score), text=ranker.truncate(get_text(docid), run.passage_length)))
    rankings.append((cur_qid, query_map[cur_qid], cur[:run.hits]))
    return rankings


def _dist_setup(args):
    """Under torchrun (WORLD_SIZE > 1): one process per GPU, queries sharded contiguously over the ranks (every
    query's rerank is independent; sort-based methods shard by query). Returns (rank, world); rank 0 gathers and writes the run."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world == 1:
        return 0, 1
    import torch
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if str(args.run.device).startswith('cuda'):
        args.run.device = f'cuda:{local}'
    if not dist.is_initialized():
        use_nccl = torch.cuda.is_available() and str(args.run.device).startswith('cuda')
        if use_nccl:
            torch.cuda.set_device(local)
        dist.init_process_group('nccl' if use_nccl else 'gloo')
    return dist.get_rank(), world


def flat_document_pieces(rankings, batch_size, rank, world):
    """Document-level sharding of a pointwise run ("shard the concatenated (query, doc) list, not queries"): the
    documents of ALL queries form one list, cut into `world` contiguous, equally long slices; a query that straddles a cut is
    scored in two pieces on two GPUs. Cuts fall on multiples of `batch_size` inside a query, so the reference's DataLoader batches —
    and with them its per-batch counters (pointwise.py:64-70,106-115) — are exactly those of the unsharded run. Returns this rank's
    pieces [(query index, start, stop)] in order; the slices of all ranks tile the list."""
    units, total = [], 0                       # (query index, start, stop) of every reference batch
    for qi, ranking in enumerate(rankings):
        for s0 in range(0, len(ranking), batch_size):
            units.append((qi, s0, min(s0 + batch_size, len(ranking))))
            total += units[-1][2] - s0
    lo_doc, hi_doc = total * rank // world, total * (rank + 1) // world
    pieces, seen = [], 0
    for qi, s0, s1 in units:                   # a batch belongs to the rank whose slice holds its first document
        if lo_doc <= seen < hi_doc:
            if pieces and pieces[-1][0] == qi and pieces[-1][2] == s0:
                pieces[-1] = (qi, pieces[-1][1], s1)
            else:
                pieces.append((qi, s0, s1))
        seen += s1 - s0
    return pieces


def main_pointwise_flat(args, ranker, first_stage, rank, world):
    """--pointwise under torchrun with document-level sharding: every rank scores its slice of the flattened (query, document) list
    through the ranker's own rerank_many (two batches in flight per GPU), the (index, score) pairs and counters are gathered ONCE at
    the end, and rank 0 applies every query's final sort. No collective inside the scoring loop."""
    from b200rank.dist import gather_lists
    items = list(first_stage)
    pieces = flat_document_pieces([r for _, _, r in items], ranker.batch_size, rank, world)
    tic = time.time()
    requests = [(items[qi][1], items[qi][2][s0:s1]) for qi, s0, s1 in pieces]
    local = []
    many = hasattr(ranker, 'rerank_many') and os.environ.get('B200RANK_RERANK_MANY', '1') != '0'
    results = ranker.rerank_many(iter(requests)) if many else (ranker.rerank(q, r) for q, r in requests)
    for (qi, s0, s1), (_, sub), _ in zip(pieces, requests, results):
        local.append((qi, s0, [float(d.score) for d in sub], ranker.total_compare, ranker.total_prompt_tokens, ranker.total_completion_tokens))
    toc = time.time()
    gathered = gather_lists(local)
    import torch.distributed as dist
    dist.barrier()
    if rank != 0:
        return
    n_cmp = n_prompt = n_completion = 0
    for qi, s0, scores, c, pt, ct in gathered:
        for doc, sc in zip(items[qi][2][s0:s0 + len(scores)], scores):
            doc.score = sc
        n_cmp, n_prompt, n_completion = n_cmp + c, n_prompt + pt, n_completion + ct
    # the ranker's own final step (pointwise.py:82,127): a stable descending sort over the query's hits in their input order
    reranked = [(qid, query, sorted(ranking, key=lambda x: x.score, reverse=True)) for qid, query, ranking in items]
    print(f'Avg comparisons: {n_cmp / len(reranked)}')
    print(f'Avg prompt tokens: {n_prompt / len(reranked)}')
    print(f'Avg completion tokens: {n_completion / len(reranked)}')
    print(f'Avg time per query: {(toc - tic) / len(reranked)}')
    write_run_file(args.run.save_path, reranked, 'LLMRankers')


def main(args):
    rank, world = _dist_setup(args)
    ranker = build_ranker(args)
    query_map, get_text = load_sources(args.run, ranker)
    logger.info(f'Loading first stage run from {args.run.run_path}.')
    first_stage = read_first_stage(args.run, ranker, query_map, get_text)

    reranked, n_cmp, n_prompt, n_completion = [], 0, 0, 0
    tic = time.time()

    def prepared():
        """(qid, query, ranking) in file order with --shuffle_ranking applied (run.py:181-189)."""
        for qid, query, ranking in first_stage:
            if args.run.shuffle_ranking is not None:
                if args.run.shuffle_ranking == 'random':
                    random.shuffle(ranking)
                elif args.run.shuffle_ranking == 'inverse':
                    ranking = ranking[::-1]
                else:
                    raise ValueError(f'Invalid shuffle ranking method: {args.run.shuffle_ranking}.')
            yield qid, query, ranking

    # The reference calls ranker.rerank() once per query (run.py:190). Rankers of this package that offer rerank_many (pointwise:
    # two queries in flight + tokenisation look-ahead; setwise heapsort: several queries' sorts in lockstep) yield exactly
    # rerank()'s result and counters per query, in order — B200RANK_RERANK_MANY=0 restores the one-call-per-query loop.
    # Queries are FED LAZILY: the shuffle of query i+1 is drawn when the ranker asks for it, i.e. after rerank(query i) for a
    # ranker that works query by query — the interleaving of run.py:181-190, which matters when the ranker itself draws from the
    # module RNG (setwise num_permutation > 1; its rerank_many falls back to one rerank() per request).
    source = prepared()
    # How work is split over the ranks (llmrankers/_backend.py::shard_mode, B200RANK_SHARD=docs|queries): document level for the rankers
    # whose per-query work is one independent prompt list, query level for the sort drivers.
    mode = 'queries'
    if world > 1:
        from llmrankers._backend import ShardedBackend, shard_mode
        kind = 'pointwise' if args.pointwise else 'pairwise' if args.pairwise else 'setwise' if args.setwise else 'listwise'
        sub = getattr(args, kind)
        mode = shard_mode(kind, getattr(sub, 'method', '') or '')
        if mode == 'docs' and kind == 'pointwise' and hasattr(ranker, 'batch_size'):
            return main_pointwise_flat(args, ranker, source, rank, world)
        if mode == 'docs' and hasattr(ranker, 'backend'):
            # pairwise allpair: every rank walks all queries; each query's n(n-1) prompts are split over the GPUs inside the backend
            ranker.backend = ShardedBackend(ranker.backend)
    if world > 1 and mode == 'queries':
        from b200rank.dist import shard_bounds
        items = list(source)                      # every rank draws the shuffles of all queries (one stream), then keeps its shard
        lo, hi = shard_bounds(len(items), rank, world)
        source = iter(items[lo:hi])
    fed = collections.deque()

    def feed():
        for qid, query, ranking in source:
            fed.append((qid, query))
            yield query, ranking

    if hasattr(ranker, 'rerank_many') and os.environ.get('B200RANK_RERANK_MANY', '1') != '0':
        results = ranker.rerank_many(feed())
    else:
        results = (ranker.rerank(query, ranking) for query, ranking in feed())
    for result in results:
        qid, query = fed.popleft()
        reranked.append((qid, query, result))
        n_cmp += ranker.total_compare
        n_prompt += ranker.total_prompt_tokens
        n_completion += ranker.total_completion_tokens
    toc = time.time()
    if world > 1 and mode == 'docs':
        import torch.distributed as dist
        dist.barrier()
        if rank != 0:       # every rank holds the complete, identical result: rank 0 reports it
            return
    elif world > 1:
        import torch.distributed as dist
        from b200rank.dist import gather_lists
        reranked = gather_lists(reranked)                       # rank order == file order (contiguous shards)
        n_cmp, n_prompt, n_completion = (sum(x) for x in zip(*gather_lists([(n_cmp, n_prompt, n_completion)])))
        dist.barrier()
        if rank != 0:
            return
    print(f'Avg comparisons: {n_cmp / len(reranked)}')
    print(f'Avg prompt tokens: {n_prompt / len(reranked)}')
    print(f'Avg completion tokens: {n_completion / len(reranked)}')
    print(f'Avg time per query: {(toc - tic) / len(reranked)}')
    write_run_file(args.run.save_path, reranked, 'LLMRankers')


# The reference's CLI grammar (run.py:206-258) as data: sub-command -> [(flag, type, default, choices, help)]. Flags, defaults and
# choices are the interface contract; --queries_tsv / --collection_tsv are this package's offline sources.
CLI_GRAMMAR = {
    'run': [('run_path', str, None, None, 'first-stage run file (TREC format) to rerank'),
            ('save_path', str, None, None, 'where to write the reranked run file (TREC format)'),
            ('model_name_or_path', str, None, None, 'checkpoint directory / hub id, or synthetic:<shape>'),
            ('tokenizer_name_or_path', str, None, None, None), ('ir_dataset_name', str, None, None, None),
            ('pyserini_index', str, None, None, None), ('hits', int, 100, None, None), ('query_length', int, 128, None, None),
            ('passage_length', int, 128, None, None), ('device', str, 'cuda', None, None), ('cache_dir', str, None, None, None),
            ('openai_key', str, None, None, None), ('scoring', str, 'generation', ['generation', 'likelihood'], None),
            ('shuffle_ranking', str, None, ['inverse', 'random'], None),
            ('queries_tsv', str, None, None, '(extension) qid<TAB>text'), ('collection_tsv', str, None, None, '(extension) docid<TAB>text')],
    'pointwise': [('method', str, 'yes_no', ['qlm', 'yes_no'], None), ('batch_size', int, 2, None, None)],
    'pairwise': [('method', str, 'allpair', ['allpair', 'heapsort', 'bubblesort'], None), ('batch_size', int, 2, None, None),
                 ('k', int, 10, None, None)],
    'setwise': [('num_child', int, 3, None, None), ('method', str, 'heapsort', ['heapsort', 'bubblesort'], None), ('k', int, 10, None, None),
                ('num_permutation', int, 1, None, None)],
    'listwise': [('window_size', int, 3, None, None), ('step_size', int, 1, None, None), ('num_repeat', int, 1, None, None)],
}


def make_parser():
    parser = argparse.ArgumentParser()
    commands = parser.add_subparsers(title='sub-commands')
    for name, flags in CLI_GRAMMAR.items():
        sub = commands.add_parser(name)
        for flag, typ, default, choices, text in flags:
            sub.add_argument('--' + flag, type=typ, default=default, choices=choices, help=text)
    return parser, commands


def cli(argv=None):
    parser, commands = make_parser()
    args = parse_args(parser, commands, argv)
    if args.run is not None and args.run.ir_dataset_name is not None and args.run.pyserini_index is not None:
        raise ValueError('Must specify either --ir_dataset_name or --pyserini_index, not both.')
    chosen = vars(args)
    if chosen['run'] is None or sum(v is not None for v in chosen.values()) != 2:
        raise ValueError('Need to set --run and can only set one of --pointwise, --pairwise, --setwise, --listwise')
    try:
        main(args)
    finally:
        # leave the rendezvous in an orderly way: a rank that exits while its peers are still inside a collective (or with the
        # process group's threads alive) aborts the interpreter on some backends
        if int(os.environ.get('WORLD_SIZE', '1')) > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                try:
                    dist.barrier()
                finally:
                    dist.destroy_process_group()


if __name__ == '__main__':
    cli()
