/*
 * b200rank — C-ABI of the B200-native Flan-T5 reranking engine (libb200rank.so).
 *
 * This is the drop-in boundary for the hot path of ielab/llm-rankers: the six call sites
 * where the reference touches a `transformers` T5 model object. Each entry point below cites
 * the reference interface it replaces (paths relative to the reference repo root; $TF =
 * transformers/models/t5/modeling_t5.py, the un-vendored dependency that holds the arithmetic).
 *
 * Conventions: plain pointers and sizes only (no torch types); caller-owned HOST buffers unless a
 * function name says `_staged`; every function returns 0 on success and a negative code on failure,
 * with a thread-local message available from b200rank_last_error(); no exceptions cross the ABI.
 * One engine per GPU. An engine is not thread-safe; engines on different GPUs may be driven
 * concurrently from different host threads/processes. There is NO CPU fallback: without a CUDA
 * device every compute entry point fails with B200RANK_ERR_CUDA.
 *
 * Token-id inputs are right-padded int32 matrices `ids[n_docs][stride]` plus `lengths[n_docs]`
 * (= attention_mask.sum(1) of DataCollatorWithPadding(padding='longest'), llmrankers/pointwise.py:45-56).
 * The engine packs the real tokens and never computes on padding; results are identical to the
 * reference's masked computation for every real token.
 */
#ifndef B200RANK_H_
#define B200RANK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RANK_OK 0
#define B200RANK_ERR_ARG (-1)      /* bad argument / unsupported configuration */
#define B200RANK_ERR_CUDA (-2)     /* CUDA runtime/driver error (incl. "no device") */
#define B200RANK_ERR_STATE (-3)    /* weights not loaded, nothing staged, ... */
#define B200RANK_ERR_CAPACITY (-4) /* a single document exceeds the engine's token capacity */

#define B200RANK_DTYPE_F32 0
#define B200RANK_DTYPE_BF16 1

/* Model + capacity description. Mirrors the fields of transformers T5Config that the forward reads
 * (llmrankers/pointwise.py:18 AutoConfig.from_pretrained; $TF:637-792). */
typedef struct b200rank_config {
    int32_t vocab_size;            /* 32128 for Flan-T5 */
    int32_t d_model;
    int32_t d_kv;                  /* must be 64 */
    int32_t num_heads;
    int32_t d_ff;
    int32_t num_layers;            /* encoder blocks */
    int32_t num_decoder_layers;
    int32_t rel_buckets;           /* relative_attention_num_buckets (32) */
    int32_t rel_max_distance;      /* relative_attention_max_distance (128) */
    float layer_norm_eps;          /* 1e-6 */
    int32_t gated_gelu;            /* 1: feed_forward_proj "gated-gelu" (Flan-T5, T5 v1.1: wi_0, wi_1); 0: "relu" (T5 v1.0, monoT5/duoT5: wi) */
    int32_t scale_decoder_outputs; /* 1 iff tie_word_embeddings: hidden *= d_model^-0.5 before lm_head ($TF:1105-1108) */
    int32_t pad_id;                /* 0 */
    int32_t eos_id;                /* 1 */
    int32_t max_tokens;            /* encoder-token capacity of one device pass (0 = default 32768) */
    int32_t max_docs;              /* documents per device pass (0 = default 1024) */
    int32_t max_dec_len;           /* longest decoder sequence (0 = default 64; hard limit 64) */
    int32_t max_logit_rows;        /* rows of the full-vocabulary logits scratch (0 = default 16384) */
} b200rank_config;

typedef struct b200rank_engine b200rank_engine;

/* Library / error plumbing (no reference equivalent; Python exceptions play this role there). */
const char* b200rank_version(void);
const char* b200rank_last_error(void);

/* Replaces T5ForConditionalGeneration.from_pretrained(...) — llmrankers/pointwise.py:20-24,
 * setwise.py:47-51, pairwise.py:56-60: allocate the model on one GPU (weights + workspaces). */
int b200rank_create(const b200rank_config* cfg, int device, b200rank_engine** out);
void b200rank_destroy(b200rank_engine* e);

/* Load one tensor by its HuggingFace state_dict name (row-major [rows, cols]; 1-D tensors as
 * [1, cols]); names as in SURVEY.md §8c: shared.weight, lm_head.weight,
 * {encoder,decoder}.final_layer_norm.weight, encoder.block.N.layer.0.SelfAttention.{q,k,v,o}.weight,
 * ...relative_attention_bias.weight, ...layer_norm.weight, ...DenseReluDense.{wi_0,wi_1,wo}.weight,
 * decoder.block.N.layer.1.EncDecAttention.{q,k,v,o}.weight, decoder.block.N.layer.2.DenseReluDense.*.
 * The engine converts to its device layout (bf16 GEMM operands, fused QKV, tile-interleaved wi_0/wi_1,
 * all decoder cross-attention K|V stacked, fp32 embedding / norm scales, expanded bias tables). */
int b200rank_load_tensor(b200rank_engine* e, const char* hf_name, const void* data, int dtype, int64_t rows,
                         int64_t cols);
/* Number of tensors still missing; optionally writes a comma-separated list of names into buf. */
int b200rank_missing_tensors(b200rank_engine* e, char* buf, int buflen);
/* The device weight arena (every loaded tensor lives inside it) for the one-time NCCL broadcast of
 * weights at load (north_star; the reference has no collective). After a broadcast into this arena on
 * a non-root rank call b200rank_mark_weights_loaded(). */
int b200rank_weights_blob(b200rank_engine* e, void** device_ptr, size_t* nbytes);
int b200rank_mark_weights_loaded(b200rank_engine* e);

/* Replaces `self.llm(input_ids, attention_mask, decoder_input_ids=[[pad]]).logits[:, :, (yes_id, no_id)]`
 * followed by softmax over the two logits — llmrankers/pointwise.py:117-124.
 * logits2: [n_docs][2] (yes, no) or NULL; scores: [n_docs] = P(yes). */
int b200rank_score_yes_no(b200rank_engine* e, const int32_t* ids, const int32_t* lengths, int n_docs, int stride,
                          int yes_id, int no_id, float* logits2, float* scores);

/* Replaces `self.llm(input_ids, attention_mask, labels=labels).logits` + CrossEntropyLoss(reduction='none')
 * summed over T and negated — llmrankers/pointwise.py:73-79. labels: [T] shared by all rows; decoder
 * inputs are shift_right(labels) ($TF:595-614). scores: [n_docs] = sum_t log p(label_t). */
int b200rank_score_qlm(b200rank_engine* e, const int32_t* ids, const int32_t* lengths, int n_docs, int stride,
                       const int32_t* labels, int T, float* scores);

/* Replaces `self.llm(input_ids, decoder_input_ids=prefix).logits[:, -1]` restricted to `cols`
 * — llmrankers/setwise.py:184-186 (normalize=1: softmax over the full vocabulary, then gather),
 * pointwise.py:173-178 / listwise.py:282-284 (normalize=0: raw logits). out: [n_docs][ncols]. */
int b200rank_logits_at(b200rank_engine* e, const int32_t* ids, const int32_t* lengths, int n_docs, int stride,
                       const int32_t* dec_prefix, int prefix_len, const int32_t* cols, int ncols, int normalize,
                       float* out);

/* Replaces `self.llm.generate(input_ids, decoder_input_ids=prefix, max_new_tokens=n)` (greedy, eos/pad from
 * the config) — llmrankers/setwise.py:93-95, pairwise.py:97-99,196-200. new_ids: [n_docs][max_new]; after a
 * row emits eos the remaining positions hold pad (transformers/generation/utils.py:2797). 1 <= max_new <= 64 and
 * prefix_len + max_new - 1 <= max_dec_len. Like the library's cached loop (generation/utils.py:2762-2804) a call keeps a
 * self-attention K/V cache: the prefix runs once, then one decoder position per generated token (B200RANK_KV_CACHE=0
 * re-runs the growing prefix instead; identical tokens for the rankers' two-token prefix). */
int b200rank_greedy(b200rank_engine* e, const int32_t* ids, const int32_t* lengths, int n_docs, int stride,
                    const int32_t* dec_prefix, int prefix_len, int max_new, int32_t* new_ids);

/* Split form of b200rank_score_yes_no for measurement and overlap: stage = pack + H2D (synchronous),
 * run = enqueue the device pass on the engine stream (asynchronous), fetch = D2H + synchronise.
 * The staged batch must fit one device pass (max_tokens / max_docs). */
int b200rank_stage(b200rank_engine* e, const int32_t* ids, const int32_t* lengths, int n_docs, int stride);
int b200rank_run_yes_no_staged(b200rank_engine* e, int yes_id, int no_id);
int b200rank_fetch_yes_no(b200rank_engine* e, float* logits2, float* scores);
int b200rank_sync(b200rank_engine* e);

/* Asynchronous form of b200rank_score_yes_no for throughput: at most two batches in flight. submit packs + copies the HOST
 * token ids and enqueues the encoder pass on the engine's main stream and the decoder pass on a second stream, so the
 * latency-bound decoder of batch i overlaps the encoder GEMMs of batch i+1 (which leave a few SMs free for it). wait blocks
 * until that batch's scores are in host memory. Each batch must fit one device pass, documents <= 240 tokens; a batch that
 * does not qualify is refused with B200RANK_ERR_ARG / B200RANK_ERR_CAPACITY before anything is enqueued (no ticket is issued,
 * batches in flight are unaffected): wait for the tickets in flight and score it with b200rank_score_yes_no instead.
 * (The reference's per-query loop `for qid, query, ranking in ...: ranker.rerank(query, ranking)` — run.py:184-192 — becomes
 * submit(query i+1); wait(query i).) */
int b200rank_submit_yes_no(b200rank_engine* e, const int32_t* ids, const int32_t* lengths, int n_docs, int stride,
                           int yes_id, int no_id, uint64_t* ticket);
int b200rank_wait_yes_no(b200rank_engine* e, uint64_t ticket, float* logits2, float* scores);

/* Measurement plumbing: CUDA events on the engine's own stream (torch.cuda.Event cannot see it). */
int b200rank_event_record(b200rank_engine* e, int which /* 0 = start, 1 = stop */);
int b200rank_event_elapsed_ms(b200rank_engine* e, float* ms); /* synchronises on the stop event */
int b200rank_launch_count(b200rank_engine* e, uint64_t* n);  /* kernels launched by this engine so far */
/* Per-launch timing: enable=1 brackets every kernel launch with CUDA events on the engine stream (labelled by kernel
 * and GEMM shape); report writes {"label": {"ms": total, "n": launches}, ...} as JSON. Used for roofline.achieved. */
int b200rank_profile(b200rank_engine* e, int enable);
int b200rank_profile_report(b200rank_engine* e, char* buf, int buflen);
int b200rank_flush_l2(b200rank_engine* e);                   /* overwrite a 256 MiB scratch (> 126 MB L2) */
int b200rank_device_info(b200rank_engine* e, int* sm_count, size_t* weight_bytes, size_t* workspace_bytes);

/* Kernel-level test hooks (used only by tests/ and profiles; host in, host out).
 * gemm: out = epilogue(A[M,K] . W[N,K]^T); epi: 0 bf16 store, 1 fp32 residual add (out pre-filled by caller),
 * 2 gated-gelu (W tile-interleaved as the engine packs wi_0/wi_1; out is [M, N/2] bf16), 3 fp32 store.
 * block_n: 0 = engine heuristic; use_simt != 0 runs the CUDA-core debug kernel instead of tcgen05. */
int b200rank_test_gemm(int device, const void* a_bf16, const void* w_bf16, int M, int N, int K, int epi, int block_n,
                       int use_simt, void* out, float* elapsed_ms);
/* encoder attention on packed qkv [tokens][3*inner] bf16; bias [H][257] fp32; out [tokens][inner] bf16.
 * mode: 0 = engine default, 1 = mma.sync 64-query tiles, 2 = mma.sync resident-KV, 3 = tcgen05, 5 = persistent tcgen05 (len <= 192), 6 = same with the row-split softmax,
 * 4 = mma.sync scores-in-registers (2, 3, 4: len <= 256) */
int b200rank_test_enc_attention(int device, const void* qkv_bf16, const int32_t* cu_seqlens, int n_docs, int num_heads,
                                const float* bias, void* out_bf16, int mode);
/* T5 relative-position bucket ($TF:189-234), host-side integer restatement used to expand the bias tables */
int b200rank_rel_bucket(int relative_position, int bidirectional, int num_buckets, int max_distance);

#ifdef __cplusplus
}
#endif
#endif /* B200RANK_H_ */
