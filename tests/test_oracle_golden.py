"""Pins the CPU oracle (oracle/t5_oracle.py) against outputs of the reference itself (tests/golden/*, produced by
tests/golden/make_golden.py from /root/reference + transformers fp32). CPU only."""
import numpy as np
import pytest

from helpers import calls, golden_meta, golden_npz, oracle_for

ATOL = 3e-4  # fp32 numpy vs fp32 torch on O(1..10) logits


def test_relative_position_buckets_match_hf():
    from oracle.t5_oracle import relative_position_bucket
    g = golden_npz("buckets.npz")
    assert np.array_equal(relative_position_bucket(g["rel"], True), g["bidirectional"])
    assert np.array_equal(relative_position_bucket(g["rel"], False), g["unidirectional"])


@pytest.mark.parametrize("which,case", [("tiny", "yes_no"), ("small", "yes_no")])
def test_yes_no_logits_scores_order(which, case):
    meta = golden_meta()
    m = meta[which]
    c = meta["cases"]["yes_no" if which == "tiny" else "small_yes_no"]
    orc = oracle_for(which)
    docs = [d["docid"] for d in m["docs"]]
    scores = []
    for call in calls(golden_npz(f"golden_{which}.npz"), "yes_no"):
        lg, sc = orc.score_yes_no(call["input_ids"], call["attention_mask"], m["yes_id"], m["no_id"])
        gold = call["logits"][:, 0, :]
        gold2 = gold[:, [m["yes_id"], m["no_id"]]] if gold.shape[-1] > 2 else gold
        np.testing.assert_allclose(lg, gold2, atol=ATOL, rtol=1e-4)
        assert np.array_equal(call["decoder_input_ids"], np.zeros((len(lg), 1)))
        scores.extend(sc.tolist())
    gold_scores = [c["scores"][d] for d in docs]
    np.testing.assert_allclose(scores, gold_scores, atol=1e-5)
    order = [d for d, _ in sorted(zip(docs, scores), key=lambda t: t[1], reverse=True)]
    assert order == c["order"]


def test_tiny_full_vocab_logits():
    orc = oracle_for("tiny")
    for call in calls(golden_npz("golden_tiny.npz"), "yes_no"):
        lg = orc.logits(call["input_ids"], call["attention_mask"], call["decoder_input_ids"])
        np.testing.assert_allclose(lg, call["logits"], atol=ATOL, rtol=1e-4)


@pytest.mark.parametrize("which", ["tiny", "small"])
def test_qlm_scores_order(which):
    meta = golden_meta()
    m = meta[which]
    c = meta["cases"]["qlm" if which == "tiny" else "small_qlm"]
    orc = oracle_for(which)
    docs = [d["docid"] for d in m["docs"]]
    scores = []
    for call in calls(golden_npz(f"golden_{which}.npz"), "qlm"):
        assert np.array_equal(call["labels"][0], np.asarray(c["labels"]))
        scores.extend(orc.score_qlm(call["input_ids"], call["attention_mask"], c["labels"]).tolist())
        if "logits" in call:
            from oracle.t5_oracle import shift_right
            lg = orc.logits(call["input_ids"], call["attention_mask"], shift_right(call["labels"]))
            np.testing.assert_allclose(lg, call["logits"], atol=ATOL, rtol=1e-4)
    gold = [c["scores"][d] for d in docs]
    np.testing.assert_allclose(scores, gold, rtol=2e-5, atol=2e-3)
    order = [d for d, _ in sorted(zip(docs, scores), key=lambda t: t[1], reverse=True)]
    assert order == c["order"]


@pytest.mark.parametrize("case", ["setwise_heap_lik", "setwise_bubble_lik"])
def test_setwise_likelihood_logits(case):
    m = golden_meta()["tiny"]
    orc = oracle_for("tiny")
    for call in calls(golden_npz("golden_tiny.npz"), case):
        assert call["input_ids"].shape[0] == 1
        lg = orc.logits(call["input_ids"], None, call["decoder_input_ids"])
        np.testing.assert_allclose(lg, call["logits"], atol=ATOL, rtol=1e-4)
        # setwise.py:184-188: softmax over the vocabulary then gather the label ids
        probs = orc.logits_at(call["input_ids"], None, m["decoder_prefix"], m["target_token_ids"], normalize=True)[0]
        from oracle.t5_oracle import softmax
        np.testing.assert_allclose(probs, softmax(call["logits"][0, -1])[m["target_token_ids"]], rtol=2e-3, atol=1e-7)


@pytest.mark.parametrize("case,lab", [("setwise_heap_gen", True), ("setwise_bubble_gen", True), ("pairwise_allpair", True),
                                      ("pairwise_heap", True), ("pairwise_bubble", True)])
def test_generation_ids(case, lab):
    m = golden_meta()["tiny"]
    orc = oracle_for("tiny", label_favouring=lab)
    n = 0
    for call in calls(golden_npz("golden_tiny.npz"), case):
        ids = call["input_ids"]
        # generate() is called WITHOUT attention_mask (setwise.py:93-95, pairwise.py:196-200). transformers 5.5.0 (the
        # version the fixtures were generated with) does not infer a mask from pad tokens for encoder-decoder models
        # (generation/utils.py:2429-2433: `not self.config.is_encoder_decoder`), so padded rows ATTEND to their pads;
        # transformers 4.31 (the reference's tested pin) would infer ids != pad. Parity follows the run-here behaviour.
        mask = np.ones_like(ids)
        new = orc.greedy(ids, mask, m["decoder_prefix"], 2)
        out = call["output"]  # [B, 2 + steps]; HF stops early when every row is finished
        steps = out.shape[1] - 2
        assert np.array_equal(out[:, :2], np.tile(np.asarray(m["decoder_prefix"])[None], (len(ids), 1)))
        assert np.array_equal(new[:, :steps], out[:, 2:]), (case, n)
        if steps < 2:
            assert np.all((new[:, :steps] == 1).any(axis=1))  # all rows hit eos, so HF stopped
        n += 1
    assert n == golden_meta()["cases"][case]["n_calls"]


# ---------------------------------------------------------------- T5 v1.0 (monoT5 / duoT5): relu feed-forward, tied + scaled lm_head
def test_v10_synthetic_model_is_relu_and_tied():
    from helpers import v10_model_and_weights
    cfg, w = v10_model_and_weights("tiny")
    assert cfg["scale_decoder_outputs"] and not cfg["gated_gelu"] and "lm_head.weight" not in w
    assert "encoder.block.0.layer.1.DenseReluDense.wi.weight" in w and "encoder.block.0.layer.1.DenseReluDense.wi_0.weight" not in w


@pytest.mark.parametrize("which,case", [("tiny", "mono"), ("small", "small_mono")])
def test_monot5_logits_scores_order(which, case):
    from helpers import golden_v10_meta, v10_oracle_for
    meta = golden_v10_meta()
    m, c = meta[which], meta["cases"][case]
    f_id, t_id = meta["false_id"], meta["true_id"]
    orc = v10_oracle_for(which)
    docs = [d["docid"] for d in m["docs"]]
    scores = []
    for call in calls(golden_npz("golden_v10.npz"), case):
        lg, sc = orc.score_yes_no(call["input_ids"], call["attention_mask"], t_id, f_id)   # (yes, no) = (true, false)
        gold = call["logits"][:, 0, :]
        gold2 = gold[:, [t_id, f_id]] if gold.shape[-1] > 2 else gold[:, [1, 0]]            # stored columns are [false, true]
        np.testing.assert_allclose(lg, gold2, atol=ATOL, rtol=1e-4)
        if gold.shape[-1] > 2:   # tiny: the whole scaled-tied-lm_head logits row
            full = orc.logits(call["input_ids"], call["attention_mask"], call["decoder_input_ids"])
            np.testing.assert_allclose(full, call["logits"], atol=ATOL, rtol=1e-4)
        scores.extend(sc.tolist())
    np.testing.assert_allclose(scores, [c["scores"][d] for d in docs], atol=1e-5)
    assert [d for d, _ in sorted(zip(docs, scores), key=lambda t: t[1], reverse=True)] == c["order"]


def test_duot5_compare_probabilities():
    from helpers import golden_v10_meta, v10_oracle_for
    meta = golden_v10_meta()
    c = meta["cases"]["duo_heap"]
    orc = v10_oracle_for("tiny")
    got = []
    for call, v in zip(calls(golden_npz("golden_v10.npz"), "duo_heap"), c["verdicts"]):
        _, p = orc.score_yes_no(call["input_ids"], call["attention_mask"], meta["true_id"], meta["false_id"])
        np.testing.assert_allclose(p, v["p"], atol=1e-5)
        got.append(bool(p[0] > p[1]))
    assert got == [v["first_wins"] for v in c["verdicts"]]


@pytest.mark.parametrize("which", ["tiny", "small"])
def test_hf_cpu_baseline_leg_reproduces_reference_fixtures(which):
    """oracle/hf_cpu.py (bench.py's cpu_baseline / --impl reference leg: transformers fp32 on torch CPU, called like
    pointwise.py:117-124) rebuilt from (model, seed) gives the logits and scores the reference's own rerank() recorded."""
    from helpers import model_and_weights
    from oracle import hf_cpu
    meta = golden_meta()
    m = meta[which]
    c = meta["cases"]["yes_no" if which == "tiny" else "small_yes_no"]
    cfg, w = model_and_weights(which)
    model = hf_cpu.build_model(cfg, w, threads=2)
    docs = [d["docid"] for d in m["docs"]]
    scores = []
    for call in calls(golden_npz(f"golden_{which}.npz"), "yes_no"):
        lg, sc = hf_cpu.score_yes_no(model, call["input_ids"], call["attention_mask"], m["yes_id"], m["no_id"], batch_size=64)
        gold = call["logits"][:, 0, :]
        gold2 = gold[:, [m["yes_id"], m["no_id"]]] if gold.shape[-1] > 2 else gold
        np.testing.assert_allclose(lg, gold2, atol=ATOL, rtol=1e-4)
        scores.extend(sc.tolist())
    np.testing.assert_allclose(scores, [c["scores"][d] for d in docs], atol=1e-5)
    # and the numpy oracle agrees with it on a fresh ragged batch (what bench.py's `parity` object relies on)
    rng = np.random.default_rng(3)
    lengths = rng.integers(9, 41, size=5)
    mask = (np.arange(40)[None] < lengths[:, None]).astype(np.int64)
    ids = rng.integers(3, cfg["vocab_size"] - 128, size=(5, 40)) * mask
    a, _ = hf_cpu.score_yes_no(model, ids, mask, m["yes_id"], m["no_id"], batch_size=2)
    b, _ = oracle_for(which).score_yes_no(ids.astype(np.int64), mask, m["yes_id"], m["no_id"])
    np.testing.assert_allclose(a, b, atol=ATOL, rtol=1e-4)


@pytest.mark.parametrize("shape", ["t5-tiny-wide", "t5v10-tiny-wide"])
def test_oracle_wide_heads_match_live_transformers(shape):
    """d_kv = 128 (the head shape of monot5-3b / duot5-3b, pointwise.py:136-186, pairwise.py:296-352): the numpy oracle against the live
    transformers fp32 model on the same synthetic weights — yes_no logits, a 3-position decoder prefix and greedy ids. This pins the
    checker the (experimental) generic-width attention path of the engine is held to on the GPU."""
    import torch
    from b200rank.synthetic import model_cfg, synthetic_weights
    from oracle.hf_cpu import build_model
    from oracle.t5_oracle import T5Oracle, pad_batch
    cfg = model_cfg(shape, 512)
    assert cfg["d_kv"] == 128
    w = synthetic_weights(cfg, 11, lm_head_std=0.5)
    rng = np.random.default_rng(5)
    rows = [rng.integers(3, 500, size=n).tolist() + [1] for n in (9, 70, 33, 1)]
    ids, mask = pad_batch(rows)
    orc = T5Oracle(cfg, w)
    model = build_model(cfg, w, threads=2)
    with torch.no_grad():
        dec = torch.tensor([[0, 17, 301]] * len(rows))
        hf = model(input_ids=torch.tensor(ids), attention_mask=torch.tensor(mask), decoder_input_ids=dec).logits.numpy()
    got = orc.logits(ids, mask, dec.numpy())
    assert got.shape == hf.shape
    assert np.abs(got - hf).max() <= 3e-4 * max(1.0, np.abs(hf).max())
    with torch.no_grad():
        gen = model.generate(input_ids=torch.tensor(ids), attention_mask=torch.tensor(mask), decoder_input_ids=torch.tensor([[0, 17]] * len(rows)),
                             max_new_tokens=2, do_sample=False).numpy()
    mine = orc.greedy(ids, mask, [0, 17], 2)
    assert np.array_equal(mine, gen[:, 2:4])


def test_headline_query_fixture_reproduces_from_the_reference_arithmetic():
    """tests/golden/headline_query.npz (the query bench.py times and the full-size GPU parity test checks all 100 documents of) holds the
    reference's fp32 logits computed in the build container. Re-derive three of its documents — the reference's rank 1, rank 10 and its
    last — with the transformers fp32 CPU path at full model size, and check the selection property the ordering claim rests on."""
    from b200rank.synthetic import NO_ID, YES_ID, headline_query, model_cfg, synthetic_weights
    from b200rank.tolerance import logit_tolerance
    from oracle import hf_cpu
    ids, lengths, ref, meta = headline_query()
    assert ref is not None and ids.shape == (100, 184) and (lengths == 184).all()
    m = ref[:, 0].astype(np.float64) - ref[:, 1]
    order = np.argsort(-m, kind="stable")
    assert [int(x) for x in order] == meta["order"]
    tol = logit_tolerance(ref, 48).max(1)
    top = order[:11]
    assert all(m[a] - m[b] >= 2 * max(tol[a], tol[b]) - 1e-6 for a, b in zip(top[:-1], top[1:])), "top-11 reference margins must be >= 2 x tolerance apart"
    assert all(m[top[-1]] - m[i] >= 2 * max(tol[top[-1]], tol[i]) - 1e-6 for i in order[11:])
    cfg = model_cfg("flan-t5-large")
    model = hf_cpu.build_model(cfg, synthetic_weights(cfg, meta["weights_seed"]))
    pick = [int(order[0]), int(order[9]), int(order[-1])]
    lg, _ = hf_cpu.score_yes_no(model, ids[pick].astype(np.int64), np.ones((3, 184), np.int64), YES_ID, NO_ID, 32)
    np.testing.assert_allclose(lg, ref[pick], atol=2e-4)
