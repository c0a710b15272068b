"""Helper for tests/test_engine_gpu.py::test_kernel_variants_agree: scores a fixed synthetic batch with whatever kernel
variant the environment selects (B200RANK_GEMM_CG, B200RANK_FUSE_NORM, B200RANK_ATTN, B200RANK_GEMM_DIRECT_EPI) and saves
the logits. Run as a subprocess because the library reads these switches once per process."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llm-rankers_b200"))


def main(out_path, model="flan-t5-base", n_docs=96):
    import b200rank as br
    from b200rank.synthetic import NO_ID, YES_ID, model_cfg, synthetic_prompt_ids, synthetic_weights
    cfg = model_cfg(model)
    c = br.make_config(cfg["d_model"], cfg["num_heads"], cfg["d_ff"], cfg["num_layers"], cfg["num_decoder_layers"],
                       max_tokens=n_docs * 192, max_docs=128, max_logit_rows=256)
    e = br.Engine(c, 0)
    e.load_state_dict(synthetic_weights(cfg, 5).items())
    ids, lengths = synthetic_prompt_ids(n_docs, 32, 128, seed=11, ragged=True)
    lg, sc = e.score_yes_no(ids, lengths, YES_ID, NO_ID)
    # the other entry points on a few documents: full-vocabulary reductions (qlm log-probs, label softmax, greedy argmax)
    k = min(16, n_docs)
    qlm = e.score_qlm(ids[:k], lengths[:k], [71, 272, 205, 309, 262, 377, 350, 1])
    # three times each: the decoder steps of these entry points run eagerly on a shape's first occurrence, are captured into a CUDA graph
    # on the second and replayed from the third on — what is saved is the replayed result
    for _ in range(3):
        probs = e.logits_at(ids[:k], lengths[:k], [0, 5], [71, 272, 205, 309], normalize=True)
        new = e.greedy(ids[:k], lengths[:k], [0, 5], 3)
    np.savez(out_path, logits=lg, scores=sc, qlm=qlm, probs=probs, greedy=new)


if __name__ == "__main__":
    main(sys.argv[1], *(sys.argv[2:3]), **({"n_docs": int(sys.argv[3])} if len(sys.argv) > 3 else {}))
