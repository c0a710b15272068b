"""Test double for llmrankers._backend.T5Backend that routes the four engine calls to the CPU oracle, so the host logic
of the drop-in rankers (prompts, batching, counters, sort drivers, output assembly, CLI) is testable without a GPU.
Lives under tests/ — the product never imports it."""
import numpy as np

from llmrankers._backend import T5Backend, generate_mask_mode


class OracleBackend(T5Backend):
    batch_invariant = False   # numpy's BLAS results depend on the batch shape in the last bits: no merged passes where tests compare exactly

    def __init__(self, oracle, tokenizer, cfg):
        super().__init__(engine=None, tokenizer=tokenizer, cfg=cfg)
        self.oracle = oracle

    def _padded(self, rows):
        ids, lengths = self.pad_rows(rows, self.pad_id)
        mask = (np.arange(ids.shape[1])[None] < lengths[:, None]).astype(np.int64)
        return ids.astype(np.int64), mask

    def score_yes_no(self, rows, yes_id, no_id):
        ids, mask = self._padded(rows)
        return self.oracle.score_yes_no(ids, mask, yes_id, no_id)

    def submit_yes_no(self, rows, yes_id, no_id):
        return ("ticket", self.score_yes_no(rows, yes_id, no_id))

    def wait_yes_no(self, ticket):
        return ticket[1]

    def score_qlm(self, rows, labels):
        ids, mask = self._padded(rows)
        return self.oracle.score_qlm(ids, mask, labels)

    def label_probs(self, rows, dec_prefix, cols):
        ids, mask = self._padded(rows)
        return self.oracle.logits_at(ids, mask, dec_prefix, cols, normalize=True)

    def generate(self, padded_ids, dec_prefix, max_new):
        ids = np.asarray(padded_ids, np.int64)
        mask = (ids != self.pad_id).astype(np.int64) if generate_mask_mode() == "infer" else np.ones_like(ids)
        new = self.oracle.greedy(ids, mask, dec_prefix, max_new, self.eos_id, self.pad_id)
        finished = np.zeros(ids.shape[0], bool)
        steps = max_new
        for s in range(max_new):
            finished |= new[:, s] == self.eos_id
            if finished.all():
                steps = s + 1
                break
        prefix = np.tile(np.asarray(dec_prefix, np.int64)[None], (ids.shape[0], 1))
        return np.concatenate([prefix, new[:, :steps]], axis=1)

    def generate_rows(self, rows, dec_prefix, max_new):
        if not rows:
            return []
        ids, mask = self._padded(rows)
        new = self.oracle.greedy(ids, mask, dec_prefix, max_new, self.eos_id, self.pad_id)
        return self._trim_rows(new, dec_prefix, max_new)

    def generate_batches(self, batches, dec_prefix, max_new):
        return [self.generate(b, dec_prefix, max_new) for b in batches]   # the definition the product's merged call must reproduce
