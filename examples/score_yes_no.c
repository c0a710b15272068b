/*
 * A plain-C consumer of include/b200rank.h: what a host that is not Python (or a maintainer wiring the library into another
 * runtime) writes against the ABI. It builds a tiny random T5 (d_model 128, 2 heads, 2+2 layers), loads every tensor by its HuggingFace name, scores
 * three ragged prompts with b200rank_score_yes_no (the call that replaces llmrankers/pointwise.py:117-124) and prints P(yes).
 *
 *   gcc -std=c99 -pedantic -Wall -Iinclude examples/score_yes_no.c -Lllm-rankers_b200 -lb200rank -Wl,-rpath,$PWD/llm-rankers_b200 -lm -o /tmp/score_yes_no
 *
 * Without a CUDA device b200rank_create fails with B200RANK_ERR_CUDA and the message from b200rank_last_error() — there is no
 * CPU fallback; tests/test_host_logic.py compiles this file and checks exactly that on the build container, and the GPU suite
 * runs it on the B200 and compares its output with the Python binding on the same seeded weights.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200rank.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static float rnd(float scale) { /* xorshift64*, uniform in [-scale, scale): deterministic across hosts */
    rng_state ^= rng_state >> 12;
    rng_state ^= rng_state << 25;
    rng_state ^= rng_state >> 27;
    return (float)(((rng_state * 0x2545F4914F6CDD1Dull) >> 40) / 16777216.0 * 2.0 - 1.0) * scale;
}

static int load(b200rank_engine* e, const char* name, int64_t rows, int64_t cols, float scale, float offset) {
    size_t n = (size_t)rows * (size_t)cols, i;
    float* w = (float*)malloc(n * sizeof(float));
    int rc;
    if (!w) return -100;
    for (i = 0; i < n; ++i) w[i] = offset + rnd(scale);
    rc = b200rank_load_tensor(e, name, w, B200RANK_DTYPE_F32, rows, cols);
    free(w);
    if (rc) fprintf(stderr, "load %s: %s\n", name, b200rank_last_error());
    return rc;
}

int main(void) {
    b200rank_config cfg;
    b200rank_engine* e = NULL;
    char name[160], missing[512];
    const char* attn[4] = {"q", "k", "v", "o"};
    int l, j, rc;
    enum { D = 128, H = 2, F = 256, L = 2, V = 2304, STRIDE = 24, NDOCS = 3 };
    int32_t ids[NDOCS][STRIDE], lengths[NDOCS] = {24, 9, 17};
    float logits[NDOCS][2], scores[NDOCS];

    memset(&cfg, 0, sizeof cfg);
    cfg.vocab_size = V; cfg.d_model = D; cfg.d_kv = 64; cfg.num_heads = H; cfg.d_ff = F;
    cfg.num_layers = L; cfg.num_decoder_layers = L; cfg.rel_buckets = 32; cfg.rel_max_distance = 128;
    cfg.layer_norm_eps = 1e-6f; cfg.gated_gelu = 1; cfg.scale_decoder_outputs = 0; cfg.pad_id = 0; cfg.eos_id = 1;
    cfg.max_tokens = 1024; cfg.max_docs = 16; cfg.max_logit_rows = 128;
    printf("%s\n", b200rank_version());
    rc = b200rank_create(&cfg, 0, &e);
    if (rc != B200RANK_OK) {
        fprintf(stderr, "b200rank_create failed (%d): %s\n", rc, b200rank_last_error());
        return rc == B200RANK_ERR_CUDA ? 3 : 1;
    }
    rc = load(e, "shared.weight", V, D, 1.0f, 0.0f) || load(e, "lm_head.weight", V, D, 0.05f, 0.0f) ||
         load(e, "encoder.final_layer_norm.weight", 1, D, 0.1f, 1.0f) || load(e, "decoder.final_layer_norm.weight", 1, D, 0.1f, 1.0f) ||
         load(e, "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight", 32, H, 0.5f, 0.0f) ||
         load(e, "decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight", 32, H, 0.5f, 0.0f);
    for (l = 0; l < L && !rc; ++l) {
        for (j = 0; j < 4 && !rc; ++j) {
            snprintf(name, sizeof name, "encoder.block.%d.layer.0.SelfAttention.%s.weight", l, attn[j]);
            rc = load(e, name, D, D, 0.06f, 0.0f);
            snprintf(name, sizeof name, "decoder.block.%d.layer.0.SelfAttention.%s.weight", l, attn[j]);
            rc = rc || load(e, name, D, D, 0.06f, 0.0f);
            snprintf(name, sizeof name, "decoder.block.%d.layer.1.EncDecAttention.%s.weight", l, attn[j]);
            rc = rc || load(e, name, D, D, 0.06f, 0.0f);
        }
        for (j = 0; j < 2 && !rc; ++j) {   /* j = 0: encoder (FFN is layer.1), j = 1: decoder (FFN is layer.2) */
            const char* side = j ? "decoder" : "encoder";
            int ffn = j ? 2 : 1, k;
            for (k = 0; k <= ffn && !rc; ++k) {
                snprintf(name, sizeof name, "%s.block.%d.layer.%d.layer_norm.weight", side, l, k);
                rc = load(e, name, 1, D, 0.1f, 1.0f);
            }
            snprintf(name, sizeof name, "%s.block.%d.layer.%d.DenseReluDense.wi_0.weight", side, l, ffn);
            rc = rc || load(e, name, F, D, 0.06f, 0.0f);
            snprintf(name, sizeof name, "%s.block.%d.layer.%d.DenseReluDense.wi_1.weight", side, l, ffn);
            rc = rc || load(e, name, F, D, 0.06f, 0.0f);
            snprintf(name, sizeof name, "%s.block.%d.layer.%d.DenseReluDense.wo.weight", side, l, ffn);
            rc = rc || load(e, name, D, F, 0.04f, 0.0f);
        }
    }
    if (rc) { b200rank_destroy(e); return 1; }
    if (b200rank_missing_tensors(e, missing, (int)sizeof missing) != 0) {
        fprintf(stderr, "missing tensors: %s\n", missing);
        b200rank_destroy(e);
        return 1;
    }
    for (l = 0; l < NDOCS; ++l)
        for (j = 0; j < STRIDE; ++j) ids[l][j] = j < lengths[l] - 1 ? 3 + (int32_t)((l * 131 + j * 17) % (V - 3)) : (j == lengths[l] - 1 ? 1 : 0);
    rc = b200rank_score_yes_no(e, &ids[0][0], lengths, NDOCS, STRIDE, /*yes_id=*/12, /*no_id=*/13, &logits[0][0], scores);
    if (rc != B200RANK_OK) {
        fprintf(stderr, "b200rank_score_yes_no failed (%d): %s\n", rc, b200rank_last_error());
        b200rank_destroy(e);
        return 1;
    }
    for (l = 0; l < NDOCS; ++l) printf("doc %d: yes %.6f no %.6f P(yes) %.6f\n", l, logits[l][0], logits[l][1], scores[l]);
    b200rank_destroy(e);
    return 0;
}
