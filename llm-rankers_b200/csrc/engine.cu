// Host side of libb200rank.so: weight arena + layout conversion, workspaces, the Flan-T5
// encoder/decoder pass as a sequence of sm_100a kernel launches on one stream, and the C-ABI
// declared in include/b200rank.h. No torch, no cuBLAS, no CPU fallback.
#include <climits>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/b200rank.h"
#include "attention_enc.cuh"
#include "attention_tc.cuh"
#include "attention_tc5.cuh"
#include "attention_dec.cuh"
#include "attention_wide.cuh"
#include "skinny_gemv.cuh"
#include "cross_ctx_t1.cuh"
#include "gemm_tcgen05.cuh"
#include "kernels_misc.cuh"

using namespace b200;
typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------ errors
static thread_local std::string g_last_error;
static int set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
#define CU_OK(expr)                                                                                         \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return set_error(B200RANK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                             __LINE__);                                                                     \
    } while (0)
#define RET_IF(expr)              \
    do {                          \
        int _r = (expr);          \
        if (_r != B200RANK_OK) return _r; \
    } while (0)

// -------------------------------------------------- driver entry point (TMA descriptors)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode_tiled = nullptr;
static int get_encode_tiled() {
    if (g_encode_tiled) return B200RANK_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn)
        return set_error(B200RANK_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    g_encode_tiled = reinterpret_cast<PFN_encodeTiled>(fn);
    return B200RANK_OK;
}

// 2-D row-major tensor map with 128-byte inner boxes and 128B swizzle: dims {cols (contiguous), rows}.
//   kind 0: bf16 GEMM operand, box {64, box_rows}      kind 1: bf16 epilogue output, box {64, 128}
//   kind 2: fp32 epilogue output (store or reduce-add), box {32, 128}
static int make_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows,
                     int kind = 0) {
    RET_IF(get_encode_tiled());
    const bool f32 = kind == 2;
    const size_t esz = f32 ? 4 : 2;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld_elems * esz};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr),
                                gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(B200RANK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%llu cols=%llu ld=%llu box=%u",
                         (int)r, ptr, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows);
    return B200RANK_OK;
}

// ------------------------------------------------------------------ kernel launch (programmatic dependent launch)
static bool pdl_enabled() {
    static int v = -1;
    if (v < 0) v = (getenv("B200RANK_PDL") && atoi(getenv("B200RANK_PDL")) == 0) ? 0 : 1;
    return v != 0;
}
// All kernels call pdl_trigger()/pdl_wait() (ptx.cuh), so any of them may be launched with the PDL attribute.
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// cudaFuncSetAttribute acts on the CURRENT device, and engines on different GPUs may live in one process and be driven from different
// host threads (include/b200rank.h): the "dynamic shared memory limit already raised" memo is per (kernel, device), not per process,
// and guarded against concurrent first launches.
struct SmemOptIn {
    std::mutex mu;
    uint64_t done = 0;   // bit d: raised on device d (a device ordinal >= 64 raises it on every launch)
    template <typename K>
    cudaError_t raise(K kern, int bytes) {
        int dev = 0;
        cudaError_t er = cudaGetDevice(&dev);
        if (er != cudaSuccess) return er;
        std::lock_guard<std::mutex> g(mu);
        if (dev >= 0 && dev < 64 && ((done >> dev) & 1ull)) return cudaSuccess;
        er = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (er == cudaSuccess && dev >= 0 && dev < 64) done |= 1ull << dev;
        return er;
    }
};

// ------------------------------------------------------------------ GEMM launch
template <int BN, int EPI, bool TMA_EPI, int CG>
static int launch_gemm_inst(cudaStream_t st, int num_sms, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout,
                            const GemmArgs& args) {
    static SmemOptIn smem_opt_in;
    auto kern = gemm_tcgen05_kernel<BN, EPI, TMA_EPI, CG>;
    using Cfg = GemmCfg<BN, CG>;
    CU_OK(smem_opt_in.raise(kern, Cfg::kSmemBytes));
    const int tiles = ((args.M + kGemmBlockM * CG - 1) / (kGemmBlockM * CG)) * ((args.N + BN - 1) / BN);
    const int grid = CG * std::min(tiles, num_sms / CG);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    CU_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, tout, args));
    return B200RANK_OK;
}

static int gemm_cta_group_pref() {
    static int pref = -1;
    if (pref < 0) {
        const char* s = getenv("B200RANK_GEMM_CG");
        pref = s ? atoi(s) : 2;
        if (pref != 1 && pref != 2) pref = 2;
    }
    return pref;
}

static int pick_block_n(int M, int N, int K, int epi, int num_sms) {
    if (epi == EPI_GATED_BF16) return 256;  // weight packing fixes the tile (HALF = 128)
    const int tiles_m = (M + kGemmBlockM - 1) / kGemmBlockM;
    const int cands[4] = {256, 128, 64, 32};
    if (M > 2 * kGemmBlockM) {
        // Several row tiles (encoder passes of any size, qlm decoder rows): minimise waves x time per tile. A tile costs its MMA
        // columns plus a fixed part (pipeline fill + epilogue drain ~ 2 us ~ 120 columns at K = 1024, relatively more for short K).
        // The largest-tile-that-fills-the-SMs rule below picked 256 for 156 tiles on 148 SMs (a second wave 5 % full) and 128 for
        // M 3300 x N 1024 (two waves where one wave of 256-wide tiles is 30 % cheaper: qlm decoder GEMMs ran at 375 TFLOP/s).
        const double fixed = 120.0 * 1024.0 / std::max(K, 64);
        int best = 256;
        double best_cost = 1e30;
        for (int i = 0; i < 4; ++i) {
            const int bn = cands[i];
            const long tiles = (long)tiles_m * ((N + bn - 1) / bn);
            const double cost = (double)((tiles + num_sms - 1) / num_sms) * (bn + fixed);
            if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }   // ties keep the larger tile
        }
        return best;
    }
    // one or two row tiles (the 100-row T = 1 decoder chain, generation prefixes): spread over the SMs
    for (int i = 0; i < 4; ++i) {
        const int bn = cands[i];
        if (tiles_m * ((N + bn - 1) / bn) >= num_sms) return bn;
    }
    return 32;
}

// CTA pairs (cta_group::2) pay off when there are enough 256-row tiles to keep all 74 pairs busy.
static int pick_cta_group(int M, int N, int bn, int num_sms) {
    if (bn != 256 || gemm_cta_group_pref() == 1) return 1;
    const int pair_tiles = ((M + 2 * kGemmBlockM - 1) / (2 * kGemmBlockM)) * ((N + bn - 1) / bn);
    return pair_tiles >= num_sms / 2 ? 2 : 1;
}

static int launch_gemm_tc(cudaStream_t st, int num_sms, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout,
                          const GemmArgs& a, int epi, int bn, bool tma_epi, int cg) {
    // the staged bf16 epilogue moves 64-column (128 B) tiles; a 32-column accumulator keeps the direct store path
    if (epi == EPI_BF16 && bn < 64) tma_epi = false;
    if (cg == 2) {
#define GEMM_CG2(EPI) \
    if (bn == 256 && epi == EPI && tma_epi) return launch_gemm_inst<256, EPI, true, 2>(st, num_sms, ta, tb, tout, a);
        GEMM_CG2(EPI_BF16) GEMM_CG2(EPI_RESID_F32) GEMM_CG2(EPI_GATED_BF16) GEMM_CG2(EPI_F32)
#undef GEMM_CG2
        return set_error(B200RANK_ERR_ARG, "no cta_group::2 GEMM instantiation for block_n=%d epi=%d tma_epi=%d", bn, epi, (int)tma_epi);
    }
#define GEMM_CASE(BN, EPI)                                                                              \
    if (bn == BN && epi == EPI)                                                                         \
        return tma_epi ? launch_gemm_inst<BN, EPI, true, 1>(st, num_sms, ta, tb, tout, a) : launch_gemm_inst<BN, EPI, false, 1>(st, num_sms, ta, tb, tout, a);
    GEMM_CASE(256, EPI_BF16) GEMM_CASE(128, EPI_BF16) GEMM_CASE(64, EPI_BF16) GEMM_CASE(32, EPI_BF16)
    GEMM_CASE(256, EPI_RESID_F32) GEMM_CASE(128, EPI_RESID_F32) GEMM_CASE(64, EPI_RESID_F32) GEMM_CASE(32, EPI_RESID_F32)
    GEMM_CASE(256, EPI_GATED_BF16)
    GEMM_CASE(256, EPI_F32) GEMM_CASE(128, EPI_F32) GEMM_CASE(64, EPI_F32) GEMM_CASE(32, EPI_F32)
#undef GEMM_CASE
    return set_error(B200RANK_ERR_ARG, "no GEMM instantiation for block_n=%d epi=%d", bn, epi);
}

// ------------------------------------------------------------------ model
struct LayerW {
    // encoder: ln1, wqkv, wo, ln2, wi, wff     decoder: + ln_c, wq_c, wo_c
    float *ln1 = nullptr, *ln2 = nullptr, *ln_c = nullptr;
    bf16 *wqkv = nullptr, *wo = nullptr, *wi = nullptr, *wff = nullptr, *wq_c = nullptr, *wo_c = nullptr;
    bf16* wov = nullptr;  // decoder only, derived at load: W_o . W_v (self-attention at T = 1 is exactly o(v(x)))
    bf16* wkT = nullptr;  // decoder only, derived at load: per-head transposed cross W_k, [H*d, 64] (q'_h = W_k,h^T q_h)
};

constexpr int kXAttnSplit = 8;   // key splits of the 2-4-position cross-attention (fixed: see run_decoder)
constexpr int kGreedyMaxNew = 64;  // b200rank_greedy: new tokens per call (bounded by max_dec_len as well)
constexpr int kKvCacheRows = 4096; // (document, position) rows of the greedy self-attention K/V cache

struct b200rank_engine {
    b200rank_config cfg;
    int device = 0, num_sms = 0;
    int d = 0, inner = 0, H = 0, F = 0, V = 0, Le = 0, Ld = 0;
    int dkv = 64;       // head width: 64 (specialised kernels) or 128 (attention_wide.cuh, experimental)
    bool gated = true;  // feed_forward_proj: gated-gelu (wi_0, wi_1) vs relu (wi)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool debug_simt = false, debug_sync = false, direct_epi = false;
    uint64_t launches = 0;

    // weight arena
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0;
    float* emb = nullptr;        // [V, d] fp32
    bf16* lm_head = nullptr;     // [V, d]
    float *enc_final_ln = nullptr, *dec_final_ln = nullptr;
    float *bias_enc = nullptr;   // [H][257]
    uint8_t* attn_bad_map = nullptr;  // [cap_tokens, H] bytes, all zero between launches: rows the tcgen05 attention leaves to its fix-up walk
    float *bias_enc_wide = nullptr;  // [H][512]: log2(e) * bias(clamp(w - 255)), the shared-memory window image of attention_tc5.cuh
    float *bias_dec = nullptr;   // [H][129]
    bf16* wckv = nullptr;        // [Ld * 2 * inner, d]  (k rows then v rows per decoder layer)
    std::vector<LayerW> enc, dec;
    std::map<std::string, bool> loaded;  // expected tensor names -> loaded?
    bool weights_ready = false;

    // workspaces
    int cap_tokens = 0, cap_docs = 0, cap_T = 0, cap_rows = 0, cap_logit_rows = 0;
    float* x = nullptr; bf16 *h = nullptr, *qkv = nullptr, *ao = nullptr, *g = nullptr, *ckv = nullptr;
    float* xd = nullptr; bf16 *hd = nullptr, *qkvd = nullptr, *aod = nullptr, *qd = nullptr, *gd = nullptr, *hlast = nullptr;
    bf16 *qp = nullptr, *ctxb = nullptr;  // [cap_docs, H*d]: re-associated T=1 cross-attention (q' and per-head context)
    float* logits = nullptr;       // [cap_logit_rows, V]
    float* small_out = nullptr;    // [cap_rows * 32] generic fp32 results
    float* small_out2 = nullptr;   // [cap_docs * 32]
    // self-attention K/V cache of greedy decoding (allocated on the first b200rank_greedy call): [Ld][kvc_rows][q | k | v] bf16, the row of
    // (document, decoder position) = document * (prefix_len + max_new) + position. The q third is where the step's fused q|k|v
    // projection lands (one GEMM, no copy); only k and v of earlier positions are read back.
    bf16* kvc = nullptr; int kvc_rows = 0;
    float* xattn_partial = nullptr;  // key-split cross-attention partials: [docs][H][nsplit][T][66] fp32 (few documents, long prompts)
    size_t xattn_partial_bytes = 0;
    int* d_ids = nullptr;          // [cap_tokens]
    int* d_cu = nullptr;           // [cap_docs + 1]
    int* d_dec_ids = nullptr;      // [cap_rows]
    int* d_cols = nullptr;         // [64]
    int* d_labels = nullptr;       // [cap_rows]
    int* d_int_out = nullptr;      // [cap_docs * (kGreedyMaxNew + 1)]: new ids | argmax scratch
    int* d_finished = nullptr;     // [cap_docs]
    uint8_t* l2_scratch = nullptr; size_t l2_scratch_bytes = 0;
    size_t workspace_bytes = 0;

    // two-deep pipeline (b200rank_submit_yes_no / _wait_yes_no): encoder passes run on `stream`, decoder passes on `stream_dec`;
    // the only tensors handed from one to the other are the encoder output and cu_seqlens, so those are double-buffered.
    cudaStream_t stream_main = nullptr, stream_dec = nullptr;
    bf16* enc_out[2] = {nullptr, nullptr};       // [cap_tokens, d] final-normed encoder output per slot
    int* d_cu_slot[2] = {nullptr, nullptr};
    bf16* enc_out_cur = nullptr;                  // what run_encoder writes / run_decoder reads
    int* d_cu_cur = nullptr;
    int* h_ids_slot[2] = {nullptr, nullptr}; int* h_cu_slot[2] = {nullptr, nullptr}; float* h_out_slot[2] = {nullptr, nullptr};
    cudaEvent_t ev_enc[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    struct Slot { int docs = 0, tokens = 0, maxlen = 0; bool busy = false; uint64_t ticket = 0; } slot[2];
    uint64_t next_ticket = 1;
    // Decoder graph (on by default, B200RANK_DEC_GRAPH=0 disables): the ~250-kernel T = 1 decoder chain of the pipelined pass is captured into a
    // CUDA graph per (slot, documents, padded longest document) on its second occurrence and replayed from the third on — one launch
    // instead of ~250 on the host, no inter-kernel launch gaps on the device. Kernel arguments (buffers, tensor maps, grids) are fixed
    // per key; ids / yes-no columns are uploaded outside the graph.
    bool dec_graph = false;
    struct DecGraph { int state = 0; cudaGraphExec_t exec = nullptr; uint64_t launches = 0; };   // 0 new, 1 seen, 2 captured, 3 never
    std::map<std::tuple<int, int, int>, DecGraph> dec_graphs;
    // the same for the decoder steps of the synchronous generation / likelihood entry points (b200rank_greedy, b200rank_logits_at): one
    // graph per (entry point, documents, decoder shape, step); a setwise / pairwise sort replays the same few shapes for every compare
    std::map<std::array<int, 8>, DecGraph> step_graphs;
    int gemm_sm_cap = 0;                          // > 0: persistent GEMMs use at most this many SMs (the rest serve the other stream)
    int pipe_reserve_sms = 0;  // measured on B200 (profiles/r01_bench_n1_v10_*): capping the encoder GEMM grids does not pay off

    // pinned host staging
    int* h_ids = nullptr; int* h_cu = nullptr; float* h_out = nullptr; int* h_int = nullptr;
    int* h_small = nullptr; size_t h_small_cap = 0, h_small_off = 0;  // pinned bump buffer for small async uploads
    bool unsynced = false;   // a synchronous-API call returned with work still enqueued on stream_main (run_yes_no_staged)
    int staged_docs = 0, staged_tokens = 0, staged_maxlen = 0;
    int staged_minlen = 0;   // shortest document of the staged pass (0 = not tracked on this path)

    std::map<std::tuple<const void*, uint64_t, uint64_t, uint64_t, uint32_t, int>, CUtensorMap> tmaps;

    // optional per-launch timing (b200rank_profile): event pairs around every kernel, keyed by a label
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;                        // pool, 2 per recorded launch
    std::vector<std::pair<std::string, size_t>> prof_records;    // (label, index of the start event)
    size_t prof_used = 0;
    std::map<std::string, std::pair<double, uint64_t>> prof_acc;  // label -> (total ms, launches)
};

static void prof_begin(b200rank_engine* e, const char* label) {
    if (!e->profiling) return;
    if (e->prof_used + 2 > e->prof_events.size()) {
        const size_t old = e->prof_events.size();
        e->prof_events.resize(old + 1024);
        for (size_t i = old; i < e->prof_events.size(); ++i) cudaEventCreate(&e->prof_events[i]);
    }
    e->prof_records.emplace_back(label, e->prof_used);
    cudaEventRecord(e->prof_events[e->prof_used], e->stream);
}
static void prof_end(b200rank_engine* e) {
    if (!e->profiling) return;
    cudaEventRecord(e->prof_events[e->prof_used + 1], e->stream);
    e->prof_used += 2;
}
static void prof_collect(b200rank_engine* e) {
    if (e->prof_records.empty()) return;
    cudaStreamSynchronize(e->stream);
    for (auto& r : e->prof_records) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e->prof_events[r.second], e->prof_events[r.second + 1]);
        auto& acc = e->prof_acc[r.first];
        acc.first += ms;
        acc.second += 1;
    }
    e->prof_records.clear();
    e->prof_used = 0;
}

static int engine_tmap(b200rank_engine* e, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                       int kind, const CUtensorMap** out) {
    auto key = std::make_tuple(ptr, rows, cols, ld, box_rows, kind);
    auto it = e->tmaps.find(key);
    if (it == e->tmaps.end()) {
        if (e->tmaps.size() > 16384) e->tmaps.clear();  // output maps are keyed by the live row count; bound the cache
        CUtensorMap m;
        RET_IF(make_tmap(&m, ptr, rows, cols, ld, box_rows, kind));
        it = e->tmaps.emplace(key, m).first;
    }
    *out = &it->second;
    return B200RANK_OK;
}

static int post_launch(b200rank_engine* e, const char* what) {
    e->launches++;
    if (strncmp(what, "gemm", 4) != 0) prof_end(e);  // gemm() closes its own interval
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return set_error(B200RANK_ERR_CUDA, "launch %s: %s", what, cudaGetErrorString(err));
    if (e->debug_sync) {
        err = cudaStreamSynchronize(e->stream);
        if (err != cudaSuccess) return set_error(B200RANK_ERR_CUDA, "kernel %s: %s", what, cudaGetErrorString(err));
    }
    return B200RANK_OK;
}

// acc[M,N] = A[M,K] . W[N,K]^T with fused epilogue. a_rows/w_rows: row capacity of the operands (TMA bounds).
static int gemm(b200rank_engine* e, const bf16* A, int lda, int a_rows, const bf16* W, int ldw, int w_rows, int M, int N,
                int K, int epi, void* out, int ldo, int force_bn, int n_per_batch, int a_cols) {
    if (M <= 0) return B200RANK_OK;
    if (N % 8 != 0 || K % 8 != 0 || lda % 8 != 0 || ldw % 8 != 0)
        return set_error(B200RANK_ERR_ARG, "gemm dims must be multiples of 8 (N=%d K=%d lda=%d ldw=%d)", N, K, lda, ldw);
    const int relu = (epi == EPI_RELU_BF16);
    int bn = force_bn ? force_bn : pick_block_n(M, N, K, relu ? EPI_BF16 : epi, e->num_sms);
    char label[96];
    if (e->profiling) snprintf(label, sizeof label, "gemm_tcgen05<bn%d,epi%d> M%d N%d K%d", bn, epi, M, N, K);  // cta group: pick_cta_group
    prof_begin(e, label);
    struct ProfEnd { b200rank_engine* e; ~ProfEnd() { prof_end(e); } } prof_end_guard{e};
    if (e->debug_simt) {
        if (n_per_batch) return set_error(B200RANK_ERR_ARG, "block-diagonal GEMM has no CUDA-core debug variant");
        const int n_out = epi == EPI_GATED_BF16 ? N / 2 : N;
        dim3 blk(32, 8), grd((n_out + 31) / 32, (M + 7) / 8);
        gemm_simt_debug_kernel<<<grd, blk, 0, e->stream>>>(A, lda, W, ldw, M, N, K, epi, 256, out, ldo);
        return post_launch(e, "gemm_simt_debug");
    }
    if (relu) epi = EPI_BF16;
    const int cg = e->direct_epi ? 1 : pick_cta_group(M, N, bn, e->num_sms);
    // The maps are COPIED out of the cache: a later lookup may clear it (it is bounded, and output maps are keyed by the live row
    // count, so a long-running process does reach the bound), which would leave pointers into it dangling.
    CUtensorMap ta, tb, tout;
    const CUtensorMap* cached = nullptr;
    RET_IF(engine_tmap(e, A, a_rows, a_cols > 0 ? a_cols : K, lda, kGemmBlockM, 0, &cached));
    ta = *cached;
    RET_IF(engine_tmap(e, W, w_rows, K, ldw, bn / cg, 0, &cached));
    tb = *cached;
    const bool out_f32 = (epi == EPI_RESID_F32 || epi == EPI_F32);
    const int n_out = (epi == EPI_GATED_BF16) ? N / 2 : N;
    // the output map carries the LIVE row count so TMA clips the ragged last M-tile
    RET_IF(engine_tmap(e, out, M, n_out, ldo, kGemmBlockM, out_f32 ? 2 : 1, &cached));
    tout = *cached;
    GemmArgs args{M, N, K, out, ldo, n_per_batch, relu};
    const int sms = (e->gemm_sm_cap > 0 && bn == 256 && M > 1024) ? std::min(e->gemm_sm_cap, e->num_sms) : e->num_sms;
    RET_IF(launch_gemm_tc(e->stream, sms, ta, tb, tout, args, epi, bn, !e->direct_epi, cg));
    return post_launch(e, "gemm_tcgen05");
}

// ------------------------------------------------------------------ relative position buckets
// modeling_t5.py:189-234. The "large" branch is max_exact + trunc(log(n/max_exact)/log(max_distance/max_exact)
// * (num_buckets - max_exact)); values that are mathematically integers (n = max_exact * ratio^(k/steps)) are snapped
// so that float rounding cannot move a bucket boundary (checked against the HF function in tests/).
extern "C" int b200rank_rel_bucket(int relative_position, int bidirectional, int num_buckets, int max_distance) {
    int ret = 0, n;
    if (bidirectional) {
        num_buckets /= 2;
        if (relative_position > 0) ret += num_buckets;
        n = std::abs(relative_position);
    } else {
        n = relative_position < 0 ? -relative_position : 0;
    }
    const int max_exact = num_buckets / 2;
    if (n < max_exact) return ret + n;
    double v = std::log(static_cast<double>(n) / max_exact) / std::log(static_cast<double>(max_distance) / max_exact) *
               (num_buckets - max_exact);
    const double rv = std::round(v);
    if (std::fabs(v - rv) < 1e-6) v = rv;
    int b = max_exact + static_cast<int>(v);
    b = std::min(b, num_buckets - 1);
    return ret + b;
}

// ------------------------------------------------------------------ create / destroy
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename T>
static int dev_alloc(b200rank_engine* e, T** p, size_t n) {
    const size_t bytes = align_up(std::max<size_t>(n, 1) * sizeof(T), 256);
    CU_OK(cudaMalloc(reinterpret_cast<void**>(p), bytes));
    CU_OK(cudaMemsetAsync(*p, 0, bytes, e->stream));
    e->workspace_bytes += bytes;
    return B200RANK_OK;
}

static void expect(b200rank_engine* e, const std::string& name) { e->loaded[name] = false; }

extern "C" const char* b200rank_version(void) { return "b200rank 0.1.0 (sm_100a, tcgen05/TMA)"; }
extern "C" const char* b200rank_last_error(void) { return g_last_error.c_str(); }

extern "C" void b200rank_destroy(b200rank_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    void* frees[] = {e->arena, e->x, e->h, e->qkv, e->ao, e->g, e->ckv, e->xd, e->hd, e->qkvd, e->aod, e->qd, e->gd, e->hlast,
                     e->logits, e->qp, e->ctxb, e->small_out, e->small_out2, e->xattn_partial, e->attn_bad_map, e->d_ids, e->d_dec_ids, e->d_cols, e->d_labels,
                     e->d_int_out, e->d_finished, e->l2_scratch, e->kvc};
    for (void* p : frees)
        if (p) cudaFree(p);
    if (e->h_ids) cudaFreeHost(e->h_ids);
    if (e->h_cu) cudaFreeHost(e->h_cu);
    if (e->h_out) cudaFreeHost(e->h_out);
    if (e->h_int) cudaFreeHost(e->h_int);
    if (e->h_small) cudaFreeHost(e->h_small);
    for (int b = 0; b < 2; ++b) {
        if (e->enc_out[b]) cudaFree(e->enc_out[b]);
        if (e->d_cu_slot[b]) cudaFree(e->d_cu_slot[b]);
        if (e->h_ids_slot[b]) cudaFreeHost(e->h_ids_slot[b]);
        if (e->h_cu_slot[b]) cudaFreeHost(e->h_cu_slot[b]);
        if (e->h_out_slot[b]) cudaFreeHost(e->h_out_slot[b]);
        if (e->ev_enc[b]) cudaEventDestroy(e->ev_enc[b]);
        if (e->ev_done[b]) cudaEventDestroy(e->ev_done[b]);
    }
    for (auto& kv : e->dec_graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    for (auto& kv : e->step_graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (e->stream_dec) cudaStreamDestroy(e->stream_dec);
    for (int i = 0; i < 2; ++i)
        if (e->ev[i]) cudaEventDestroy(e->ev[i]);
    for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
    // e->stream is an alias that points at stream_main / stream_dec in turn: destroy the owning handle
    cudaStream_t main_stream = e->stream_main ? e->stream_main : e->stream;
    if (main_stream) cudaStreamDestroy(main_stream);
    delete e;
}

static int create_impl(b200rank_engine* e) {
    const b200rank_config& c = e->cfg;
    CU_OK(cudaSetDevice(e->device));
    cudaDeviceProp prop;
    CU_OK(cudaGetDeviceProperties(&prop, e->device));
    if (prop.major != 10)
        return set_error(B200RANK_ERR_CUDA, "device %d is sm_%d%d; this library contains sm_100a code only", e->device, prop.major,
                         prop.minor);
    e->num_sms = prop.multiProcessorCount;
    {
        // Stream priorities: equal by default. Measured on B200 (profiles/r01_bench_n1_v11_*): giving the short decoder kernels the
        // HIGHEST priority costs ~4 % (they cut into the encoder GEMM waves); B200RANK_PIPE_PRIORITY=1 selects it, =2 the opposite
        // (encoder stream above the decoder stream: the chain only takes SMs no encoder CTA is waiting for).
        int prio_lo = 0, prio_hi = 0;
        CU_OK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        const int mode = getenv("B200RANK_PIPE_PRIORITY") ? atoi(getenv("B200RANK_PIPE_PRIORITY")) : 0;
        CU_OK(cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, mode == 2 ? prio_hi : prio_lo));
        e->stream_main = e->stream;
        CU_OK(cudaStreamCreateWithPriority(&e->stream_dec, cudaStreamNonBlocking, mode == 1 ? prio_hi : prio_lo));
    }
    for (int b = 0; b < 2; ++b) {
        CU_OK(cudaEventCreateWithFlags(&e->ev_enc[b], cudaEventDisableTiming));
        CU_OK(cudaEventCreateWithFlags(&e->ev_done[b], cudaEventDisableTiming));
    }
    if (getenv("B200RANK_PIPE_RESERVE_SMS")) e->pipe_reserve_sms = atoi(getenv("B200RANK_PIPE_RESERVE_SMS"));
    CU_OK(cudaEventCreate(&e->ev[0]));
    CU_OK(cudaEventCreate(&e->ev[1]));
    e->debug_simt = getenv("B200RANK_DEBUG_SIMT_GEMM") && atoi(getenv("B200RANK_DEBUG_SIMT_GEMM")) != 0;
    e->debug_sync = getenv("B200RANK_DEBUG_SYNC") && atoi(getenv("B200RANK_DEBUG_SYNC")) != 0;
    e->direct_epi = getenv("B200RANK_GEMM_DIRECT_EPI") && atoi(getenv("B200RANK_GEMM_DIRECT_EPI")) != 0;

    const size_t d = e->d, I = e->inner, F = e->F, V = e->V;
    // ---- weight arena layout
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    std::vector<std::pair<void**, size_t>> fix;  // (pointer slot, offset)
    auto reserve = [&](void** slot, size_t bytes) { fix.emplace_back(slot, take(bytes)); };
    reserve((void**)&e->emb, V * d * 4);
    reserve((void**)&e->lm_head, V * d * 2);
    reserve((void**)&e->enc_final_ln, d * 4);
    reserve((void**)&e->dec_final_ln, d * 4);
    reserve((void**)&e->bias_enc, (size_t)e->H * kAttnBiasLen * 4);
    reserve((void**)&e->bias_enc_wide, (size_t)e->H * 512 * 4);
    reserve((void**)&e->bias_dec, (size_t)e->H * (kAttnRelClamp + 1) * 4);
    reserve((void**)&e->wckv, (size_t)e->Ld * 2 * I * d * 2);
    e->enc.resize(e->Le);
    e->dec.resize(e->Ld);
    for (int l = 0; l < e->Le; ++l) {
        LayerW& w = e->enc[l];
        reserve((void**)&w.ln1, d * 4); reserve((void**)&w.ln2, d * 4);
        reserve((void**)&w.wqkv, 3 * I * d * 2); reserve((void**)&w.wo, d * I * 2);
        reserve((void**)&w.wi, (e->gated ? 2 : 1) * F * d * 2); reserve((void**)&w.wff, d * F * 2);
    }
    for (int l = 0; l < e->Ld; ++l) {
        LayerW& w = e->dec[l];
        reserve((void**)&w.ln1, d * 4); reserve((void**)&w.ln_c, d * 4); reserve((void**)&w.ln2, d * 4);
        reserve((void**)&w.wqkv, 3 * I * d * 2); reserve((void**)&w.wo, d * I * 2);
        reserve((void**)&w.wq_c, I * d * 2); reserve((void**)&w.wo_c, d * I * 2); reserve((void**)&w.wov, d * d * 2); reserve((void**)&w.wkT, (size_t)e->H * d * 64 * 2);
        reserve((void**)&w.wi, (e->gated ? 2 : 1) * F * d * 2); reserve((void**)&w.wff, d * F * 2);
    }
    e->arena_bytes = off;
    CU_OK(cudaMalloc(reinterpret_cast<void**>(&e->arena), e->arena_bytes));
    CU_OK(cudaMemsetAsync(e->arena, 0, e->arena_bytes, e->stream));
    for (auto& f : fix) *f.first = e->arena + f.second;

    // ---- expected tensors
    expect(e, "shared.weight"); expect(e, "lm_head.weight");
    expect(e, "encoder.final_layer_norm.weight"); expect(e, "decoder.final_layer_norm.weight");
    expect(e, "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight");
    expect(e, "decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight");
    char nm[256];
    for (int l = 0; l < e->Le; ++l) {
        for (const char* m : {"q", "k", "v", "o"}) { snprintf(nm, sizeof nm, "encoder.block.%d.layer.0.SelfAttention.%s.weight", l, m); expect(e, nm); }
        snprintf(nm, sizeof nm, "encoder.block.%d.layer.0.layer_norm.weight", l); expect(e, nm);
        for (const char* m : {"wi_0", "wi_1", "wi", "wo"}) {
            if ((m[2] == '_') != e->gated && m[1] == 'i') continue;  // gated: wi_0 + wi_1, T5 v1.0: wi
            snprintf(nm, sizeof nm, "encoder.block.%d.layer.1.DenseReluDense.%s.weight", l, m); expect(e, nm);
        }
        snprintf(nm, sizeof nm, "encoder.block.%d.layer.1.layer_norm.weight", l); expect(e, nm);
    }
    for (int l = 0; l < e->Ld; ++l) {
        for (const char* m : {"q", "k", "v", "o"}) {
            snprintf(nm, sizeof nm, "decoder.block.%d.layer.0.SelfAttention.%s.weight", l, m); expect(e, nm);
            snprintf(nm, sizeof nm, "decoder.block.%d.layer.1.EncDecAttention.%s.weight", l, m); expect(e, nm);
        }
        for (int j = 0; j < 3; ++j) { snprintf(nm, sizeof nm, "decoder.block.%d.layer.%d.layer_norm.weight", l, j); expect(e, nm); }
        for (const char* m : {"wi_0", "wi_1", "wi", "wo"}) {
            if ((m[2] == '_') != e->gated && m[1] == 'i') continue;
            snprintf(nm, sizeof nm, "decoder.block.%d.layer.2.DenseReluDense.%s.weight", l, m); expect(e, nm);
        }
    }

    // ---- workspaces (row capacities rounded to the 128-row GEMM tile)
    e->cap_tokens = (int)align_up(c.max_tokens > 0 ? c.max_tokens : 32768, 128);
    e->cap_docs = c.max_docs > 0 ? c.max_docs : 1024;
    e->cap_T = c.max_dec_len > 0 ? c.max_dec_len : 64;
    if (e->cap_T > 64) return set_error(B200RANK_ERR_ARG, "max_dec_len %d > 64 unsupported", e->cap_T);
    e->cap_logit_rows = (int)align_up(c.max_logit_rows > 0 ? c.max_logit_rows : 16384, 128);   // qlm: 496 documents x 33 label positions per device pass
    e->cap_rows = (int)align_up(std::max(e->cap_docs, e->cap_logit_rows), 128);
    const size_t Tk = e->cap_tokens, R = e->cap_rows;
    RET_IF(dev_alloc(e, &e->x, Tk * d)); RET_IF(dev_alloc(e, &e->h, Tk * d));
    RET_IF(dev_alloc(e, &e->qkv, Tk * 3 * I)); RET_IF(dev_alloc(e, &e->ao, Tk * I));
    RET_IF(dev_alloc(e, &e->g, Tk * F)); RET_IF(dev_alloc(e, &e->ckv, Tk * (size_t)e->Ld * 2 * I));
    RET_IF(dev_alloc(e, &e->xd, R * d)); RET_IF(dev_alloc(e, &e->hd, R * d));
    RET_IF(dev_alloc(e, &e->qkvd, R * 3 * I)); RET_IF(dev_alloc(e, &e->aod, R * I));
    RET_IF(dev_alloc(e, &e->qd, R * I)); RET_IF(dev_alloc(e, &e->gd, R * F)); RET_IF(dev_alloc(e, &e->hlast, R * d));
    RET_IF(dev_alloc(e, &e->logits, (size_t)e->cap_logit_rows * V));
    RET_IF(dev_alloc(e, &e->qp, (size_t)align_up(e->cap_docs, 128) * e->H * d)); RET_IF(dev_alloc(e, &e->ctxb, (size_t)align_up(e->cap_docs, 128) * e->H * d));
    RET_IF(dev_alloc(e, &e->small_out, R * 32)); RET_IF(dev_alloc(e, &e->small_out2, (size_t)e->cap_docs * 32));
    e->xattn_partial_bytes = (size_t)std::min(e->cap_docs, 256) * e->H * kXAttnSplit * 4 * 66 * sizeof(float);   // groups of <= 256 documents, T <= 4
    RET_IF(dev_alloc(e, &e->xattn_partial, e->xattn_partial_bytes / sizeof(float)));
    RET_IF(dev_alloc(e, &e->d_ids, Tk));
    RET_IF(dev_alloc(e, &e->attn_bad_map, (size_t)Tk * e->H));
    CU_OK(cudaMemset(e->attn_bad_map, 0, (size_t)Tk * e->H));
    for (int b = 0; b < 2; ++b) {
        RET_IF(dev_alloc(e, &e->enc_out[b], Tk * d));
        RET_IF(dev_alloc(e, &e->d_cu_slot[b], (size_t)e->cap_docs + 1));
        CU_OK(cudaHostAlloc(reinterpret_cast<void**>(&e->h_ids_slot[b]), Tk * sizeof(int), cudaHostAllocDefault));
        CU_OK(cudaHostAlloc(reinterpret_cast<void**>(&e->h_cu_slot[b]), ((size_t)e->cap_docs + 1) * sizeof(int), cudaHostAllocDefault));
        CU_OK(cudaHostAlloc(reinterpret_cast<void**>(&e->h_out_slot[b]), (size_t)e->cap_docs * 4 * sizeof(float), cudaHostAllocDefault));
    }
    e->enc_out_cur = e->enc_out[0];
    e->d_cu = e->d_cu_slot[0];
    e->d_cu_cur = e->d_cu;
    // the decoder graph is on unless B200RANK_DEC_GRAPH=0 (bit-identical to the eager chain: tests/test_engine_gpu.py)
    e->dec_graph = !(getenv("B200RANK_DEC_GRAPH") && atoi(getenv("B200RANK_DEC_GRAPH")) == 0);
    RET_IF(dev_alloc(e, &e->d_dec_ids, R)); RET_IF(dev_alloc(e, &e->d_cols, 64)); RET_IF(dev_alloc(e, &e->d_labels, R));
    RET_IF(dev_alloc(e, &e->d_int_out, (size_t)e->cap_docs * (kGreedyMaxNew + 1))); RET_IF(dev_alloc(e, &e->d_finished, (size_t)e->cap_docs));
    CU_OK(cudaHostAlloc(reinterpret_cast<void**>(&e->h_ids), Tk * sizeof(int), cudaHostAllocDefault));
    CU_OK(cudaHostAlloc(reinterpret_cast<void**>(&e->h_cu), ((size_t)e->cap_docs + 1) * sizeof(int), cudaHostAllocDefault));
    CU_OK(cudaHostAlloc(reinterpret_cast<void**>(&e->h_out), (size_t)e->cap_docs * 64 * sizeof(float), cudaHostAllocDefault));
    CU_OK(cudaHostAlloc(reinterpret_cast<void**>(&e->h_int), (size_t)e->cap_docs * kGreedyMaxNew * sizeof(int), cudaHostAllocDefault));
    e->h_small_cap = 4 * R + 4096;
    CU_OK(cudaHostAlloc(reinterpret_cast<void**>(&e->h_small), e->h_small_cap * sizeof(int), cudaHostAllocDefault));
    CU_OK(cudaStreamSynchronize(e->stream));
    return B200RANK_OK;
}

extern "C" int b200rank_create(const b200rank_config* cfg, int device, b200rank_engine** out) {
    if (!cfg || !out) return set_error(B200RANK_ERR_ARG, "null argument");
    *out = nullptr;
    // d_kv = 64 runs on the specialised kernels; d_kv = 128 (monot5-3b / duot5-3b) on the generic-width mma.sync attention of
    // attention_wide.cuh (validated on the B200 in round 2: tests/test_engine_gpu.py::test_wide_heads_vs_oracle)
    if (cfg->d_kv != 64 && cfg->d_kv != 128)
        return set_error(B200RANK_ERR_ARG, "d_kv=%d unsupported (64 and 128 are implemented)", cfg->d_kv);
    if (cfg->d_model <= 0 || cfg->d_ff <= 0 || cfg->vocab_size <= 0 || cfg->d_model % 64 || cfg->d_ff % 128 || cfg->d_model > 4096 ||
        cfg->d_ff > (1 << 17) || cfg->vocab_size % 8 || cfg->vocab_size > (1 << 22))
        return set_error(B200RANK_ERR_ARG, "unsupported dims d_model=%d d_ff=%d vocab=%d", cfg->d_model, cfg->d_ff, cfg->vocab_size);
    if (cfg->num_heads <= 0 || cfg->num_heads > 512 || cfg->num_layers <= 0 || cfg->num_layers > 256 || cfg->num_decoder_layers <= 0 ||
        cfg->num_decoder_layers > 256)
        return set_error(B200RANK_ERR_ARG, "bad layer/head counts");
    // relative-position buckets: T5's formula needs an even bucket count >= 4 (half the buckets per direction, half of those exact);
    // distances saturate at max_distance, and the bias tables of the attention kernels cover +-128 (b200rank_load_tensor re-checks)
    if (cfg->rel_buckets < 4 || cfg->rel_buckets > 1024 || (cfg->rel_buckets & 1) || cfg->rel_max_distance < 2 || cfg->rel_max_distance > kAttnRelClamp)
        return set_error(B200RANK_ERR_ARG, "unsupported relative attention: %d buckets, max distance %d (even bucket count >= 4, max distance 2..%d)",
                         cfg->rel_buckets, cfg->rel_max_distance, kAttnRelClamp);
    if (cfg->max_tokens < 0 || cfg->max_docs < 0 || cfg->max_dec_len < 0 || cfg->max_logit_rows < 0 || !(cfg->layer_norm_eps > 0.f))
        return set_error(B200RANK_ERR_ARG, "capacities (max_tokens / max_docs / max_dec_len / max_logit_rows) must be >= 0 (0 = default) and layer_norm_eps > 0");
    if (cfg->pad_id < 0 || cfg->pad_id >= cfg->vocab_size || cfg->eos_id < 0 || cfg->eos_id >= cfg->vocab_size)
        return set_error(B200RANK_ERR_ARG, "pad_id / eos_id outside the vocabulary");
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev <= 0)
        return set_error(B200RANK_ERR_CUDA, "no CUDA device available (%s); b200rank has no CPU fallback",
                         err == cudaSuccess ? "device count 0" : cudaGetErrorString(err));
    if (device < 0 || device >= ndev) return set_error(B200RANK_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    b200rank_engine* e = new b200rank_engine();
    e->cfg = *cfg;
    e->device = device;
    e->d = cfg->d_model; e->H = cfg->num_heads; e->inner = cfg->num_heads * cfg->d_kv; e->F = cfg->d_ff; e->V = cfg->vocab_size;
    e->dkv = cfg->d_kv;
    e->Le = cfg->num_layers; e->Ld = cfg->num_decoder_layers;
    e->gated = cfg->gated_gelu != 0;
    int r = create_impl(e);
    if (r != B200RANK_OK) { std::string keep = g_last_error; b200rank_destroy(e); g_last_error = keep; return r; }
    *out = e;
    return B200RANK_OK;
}

// ------------------------------------------------------------------ weight loading
static inline bf16 to_bf16(float f) { return __float2bfloat16_rn(f); }
static inline float src_at(const void* data, int dtype, size_t i) {
    if (dtype == B200RANK_DTYPE_F32) return reinterpret_cast<const float*>(data)[i];
    return __bfloat162float(reinterpret_cast<const bf16*>(data)[i]);
}

// copy a [rows, cols] host matrix as bf16 into dst rows given by row_map(r)
template <typename RowMap>
static int put_bf16_rows(b200rank_engine* e, bf16* dst, size_t dst_ld, const void* data, int dtype, int64_t rows, int64_t cols,
                         RowMap row_map) {
    std::vector<bf16> tmp(static_cast<size_t>(rows) * cols);
    for (size_t i = 0; i < tmp.size(); ++i) tmp[i] = to_bf16(src_at(data, dtype, i));
    // contiguous destination runs are common (identity map): detect and use one 2-D copy
    bool contiguous = true;
    const int64_t r0 = row_map(0);
    for (int64_t r = 1; r < rows && contiguous; ++r) contiguous = (row_map(r) == r0 + r);
    if (contiguous) {
        CU_OK(cudaMemcpy2D(dst + r0 * dst_ld, dst_ld * 2, tmp.data(), cols * 2, cols * 2, rows, cudaMemcpyHostToDevice));
    } else {
        // runs of consecutive rows
        int64_t r = 0;
        while (r < rows) {
            int64_t r1 = r + 1;
            while (r1 < rows && row_map(r1) == row_map(r) + (r1 - r)) ++r1;
            CU_OK(cudaMemcpy2D(dst + row_map(r) * dst_ld, dst_ld * 2, tmp.data() + r * cols, cols * 2, cols * 2, r1 - r,
                               cudaMemcpyHostToDevice));
            r = r1;
        }
    }
    return B200RANK_OK;
}
static int put_f32(float* dst, const void* data, int dtype, size_t n) {
    std::vector<float> tmp(n);
    for (size_t i = 0; i < n; ++i) tmp[i] = src_at(data, dtype, i);
    CU_OK(cudaMemcpy(dst, tmp.data(), n * 4, cudaMemcpyHostToDevice));
    return B200RANK_OK;
}

static int shape_check(const char* name, int64_t rows, int64_t cols, int64_t er, int64_t ec) {
    if (rows != er || cols != ec)
        return set_error(B200RANK_ERR_ARG, "tensor %s has shape [%lld, %lld], expected [%lld, %lld]", name, (long long)rows,
                         (long long)cols, (long long)er, (long long)ec);
    return B200RANK_OK;
}

static int gemm(b200rank_engine* e, const bf16* A, int lda, int a_rows, const bf16* W, int ldw, int w_rows, int M, int N,
                int K, int epi, void* out, int ldo, int force_bn = 0,
                int n_per_batch = 0, int a_cols = 0);

// Derived weights, computed once on the device when the last tensor arrives (they live in the arena, so the NCCL
// broadcast carries them): decoder W_ov[l] = W_o[l] . W_v[l]  (bf16 operands, fp32 accumulate, bf16 result).
// With a single decoder position the causal softmax is over one key and equals 1, so self-attention is o(v(norm(x)))
// (modeling_t5.py:350-377) and the q/k projections are dead work.
static int derive_weights(b200rank_engine* e) {
    const int d = e->d, I = e->inner;
    bf16* vt = nullptr;  // W_v^T [d, I]
    CU_OK(cudaMalloc(reinterpret_cast<void**>(&vt), (size_t)align_up(d, 256) * I * sizeof(bf16)));
    CU_OK(cudaMemsetAsync(vt, 0, (size_t)align_up(d, 256) * I * sizeof(bf16), e->stream));
    int rc = B200RANK_OK;
    for (int l = 0; l < e->Ld && rc == B200RANK_OK; ++l) {
        const LayerW& w = e->dec[l];
        launch_k(transpose_bf16_kernel, dim3(dim3((d + 31) / 32, (I + 31) / 32)), dim3(dim3(32, 8)), 0, e->stream, w.wqkv + (size_t)2 * I * d, I, d, d, vt, I);
        rc = post_launch(e, "transpose_bf16");
        // W_ov[i, j] = sum_k W_o[i, k] W_v[k, j]  ==  A[M=d, K=I] . W[N=d, K=I]^T with W = W_v^T
        if (rc == B200RANK_OK) rc = gemm(e, w.wo, I, d, vt, I, (int)align_up(d, 256), d, d, I, EPI_BF16, w.wov, d, 0);
        // per-head transposed cross-attention W_k: wkT[h*d + n, m] = W_k[h*64 + m, n]
        for (int h = 0; h < e->H && rc == B200RANK_OK && e->dkv == 64; ++h) {   // (the re-associated T = 1 path is d_kv = 64 only)
            const bf16* wk_h = e->wckv + ((size_t)l * 2 * I + (size_t)h * 64) * d;
            launch_k(transpose_bf16_kernel, dim3(dim3((d + 31) / 32, 2)), dim3(dim3(32, 8)), 0, e->stream, wk_h, 64, d, d, w.wkT + (size_t)h * d * 64, 64);
            rc = post_launch(e, "transpose_bf16");
        }
    }
    cudaError_t err = cudaStreamSynchronize(e->stream);
    cudaFree(vt);
    e->tmaps.clear();  // drop the tensor maps of the temporary
    if (rc != B200RANK_OK) return rc;
    if (err != cudaSuccess) return set_error(B200RANK_ERR_CUDA, "derive_weights: %s", cudaGetErrorString(err));
    return B200RANK_OK;
}

// [H][257] position-bias table -> [H][512] shared-memory window image of the tcgen05 attention kernel: entry w = log2(e) * bias(clamp(w - 255))
static std::vector<float> widen_enc_bias(const float* table, int H) {
    std::vector<float> wide((size_t)H * 512, 0.f);
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < 511; ++w) {
            const int rel = std::max(-kAttnRelClamp, std::min(kAttnRelClamp, w - 255));
            wide[(size_t)h * 512 + w] = table[(size_t)h * kAttnBiasLen + rel + kAttnRelClamp] * 1.4426950408889634f;
        }
    return wide;
}

extern "C" int b200rank_load_tensor(b200rank_engine* e, const char* hf_name, const void* data, int dtype, int64_t rows,
                                    int64_t cols) {
    if (!e || !hf_name || !data) return set_error(B200RANK_ERR_ARG, "null argument");
    if (dtype != B200RANK_DTYPE_F32 && dtype != B200RANK_DTYPE_BF16) return set_error(B200RANK_ERR_ARG, "bad dtype %d", dtype);
    CU_OK(cudaSetDevice(e->device));
    std::string name(hf_name);
    const int64_t d = e->d, I = e->inner, F = e->F, V = e->V;
    auto ident = [](int64_t r) { return r; };
    if (name == "encoder.embed_tokens.weight" || name == "decoder.embed_tokens.weight") return B200RANK_OK;  // aliases of shared.weight
    // T5 v1.0 checkpoints (t5-*, the monoT5 / duoT5 .bin lineage) carry a cross-attention relative bias that no forward reads;
    // transformers drops it silently (_keys_to_ignore_on_load_unexpected, modeling_t5.py), and so does this loader
    {
        static const std::string ignored = "EncDecAttention.relative_attention_bias.weight";
        if (name.size() >= ignored.size() && name.compare(name.size() - ignored.size(), ignored.size(), ignored) == 0) return B200RANK_OK;
    }
    if (e->loaded.find(name) == e->loaded.end()) return set_error(B200RANK_ERR_ARG, "unexpected tensor name %s", hf_name);

    int r = B200RANK_OK;
    if (name == "shared.weight") {
        RET_IF(shape_check(hf_name, rows, cols, V, d));
        r = put_f32(e->emb, data, dtype, (size_t)V * d);
    } else if (name == "lm_head.weight") {
        RET_IF(shape_check(hf_name, rows, cols, V, d));
        r = put_bf16_rows(e, e->lm_head, d, data, dtype, rows, cols, ident);
    } else if (name == "encoder.final_layer_norm.weight") {
        RET_IF(shape_check(hf_name, rows * cols, 1, d, 1));
        r = put_f32(e->enc_final_ln, data, dtype, d);
    } else if (name == "decoder.final_layer_norm.weight") {
        RET_IF(shape_check(hf_name, rows * cols, 1, d, 1));
        r = put_f32(e->dec_final_ln, data, dtype, d);
    } else {
        int l = -1, sub = -1;
        char stack[16] = {0}, rest[160] = {0};
        if (sscanf(hf_name, "%7[a-z].block.%d.layer.%d.%159s", stack, &l, &sub, rest) != 4)
            return set_error(B200RANK_ERR_ARG, "cannot parse tensor name %s", hf_name);
        const bool is_dec = strcmp(stack, "decoder") == 0;
        if (l < 0 || l >= (is_dec ? e->Ld : e->Le)) return set_error(B200RANK_ERR_ARG, "layer index out of range in %s", hf_name);
        LayerW& w = is_dec ? e->dec[l] : e->enc[l];
        std::string rs(rest);
        const int ff_sub = is_dec ? 2 : 1;
        if (rs == "layer_norm.weight") {
            RET_IF(shape_check(hf_name, rows * cols, 1, d, 1));
            float* dst = (sub == 0) ? w.ln1 : (sub == ff_sub ? w.ln2 : w.ln_c);
            r = put_f32(dst, data, dtype, d);
        } else if (rs == "SelfAttention.relative_attention_bias.weight") {
            RET_IF(shape_check(hf_name, rows, cols, e->cfg.rel_buckets, e->H));
            const int nb = e->cfg.rel_buckets, md = e->cfg.rel_max_distance;
            if (md > kAttnRelClamp) return set_error(B200RANK_ERR_ARG, "rel_max_distance %d > %d unsupported", md, kAttnRelClamp);
            if (!is_dec) {
                std::vector<float> t((size_t)e->H * kAttnBiasLen);
                for (int h = 0; h < e->H; ++h)
                    for (int dlt = -kAttnRelClamp; dlt <= kAttnRelClamp; ++dlt)
                        t[(size_t)h * kAttnBiasLen + dlt + kAttnRelClamp] =
                            src_at(data, dtype, (size_t)b200rank_rel_bucket(dlt, 1, nb, md) * e->H + h);
                CU_OK(cudaMemcpy(e->bias_enc, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
                std::vector<float> wide = widen_enc_bias(t.data(), e->H);
                CU_OK(cudaMemcpy(e->bias_enc_wide, wide.data(), wide.size() * 4, cudaMemcpyHostToDevice));
            } else {
                const int len = kAttnRelClamp + 1;
                std::vector<float> t((size_t)e->H * len);
                for (int h = 0; h < e->H; ++h)
                    for (int n = 0; n < len; ++n)
                        t[(size_t)h * len + n] = src_at(data, dtype, (size_t)b200rank_rel_bucket(-n, 0, nb, md) * e->H + h);
                CU_OK(cudaMemcpy(e->bias_dec, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
            }
        } else if (rs.rfind("SelfAttention.", 0) == 0 && sub == 0) {
            const char m = rs[14];
            if (m == 'o') { RET_IF(shape_check(hf_name, rows, cols, d, I)); r = put_bf16_rows(e, w.wo, I, data, dtype, rows, cols, ident); }
            else {
                RET_IF(shape_check(hf_name, rows, cols, I, d));
                const int64_t base = (m == 'q') ? 0 : (m == 'k' ? I : 2 * I);
                r = put_bf16_rows(e, w.wqkv, d, data, dtype, rows, cols, [base](int64_t rr) { return base + rr; });
            }
        } else if (rs.rfind("EncDecAttention.", 0) == 0 && is_dec && sub == 1) {
            const char m = rs[16];
            if (m == 'o') { RET_IF(shape_check(hf_name, rows, cols, d, I)); r = put_bf16_rows(e, w.wo_c, I, data, dtype, rows, cols, ident); }
            else if (m == 'q') { RET_IF(shape_check(hf_name, rows, cols, I, d)); r = put_bf16_rows(e, w.wq_c, d, data, dtype, rows, cols, ident); }
            else {
                RET_IF(shape_check(hf_name, rows, cols, I, d));
                const int64_t base = (int64_t)l * 2 * I + (m == 'k' ? 0 : I);
                r = put_bf16_rows(e, e->wckv, d, data, dtype, rows, cols, [base](int64_t rr) { return base + rr; });
            }
        } else if (rs.rfind("DenseReluDense.", 0) == 0 && sub == ff_sub) {
            std::string m = rs.substr(15);
            if (m == "wo.weight") { RET_IF(shape_check(hf_name, rows, cols, d, F)); r = put_bf16_rows(e, w.wff, F, data, dtype, rows, cols, ident); }
            else if ((m == "wi_0.weight" || m == "wi_1.weight") && e->gated) {
                RET_IF(shape_check(hf_name, rows, cols, F, d));
                // tile interleave for the gated epilogue: N-tile nb of 256 accumulator columns =
                // [wi_0 rows nb*128 .. +128 | wi_1 rows nb*128 .. +128]
                const int64_t add = (m == "wi_1.weight") ? 128 : 0;
                r = put_bf16_rows(e, w.wi, d, data, dtype, rows, cols, [add](int64_t f) { return (f / 128) * 256 + add + (f % 128); });
            } else if (m == "wi.weight" && !e->gated) {
                RET_IF(shape_check(hf_name, rows, cols, F, d));
                r = put_bf16_rows(e, w.wi, d, data, dtype, rows, cols, ident);
            } else return set_error(B200RANK_ERR_ARG, "unknown feed-forward tensor %s", hf_name);
        } else {
            return set_error(B200RANK_ERR_ARG, "unknown tensor %s", hf_name);
        }
    }
    if (r != B200RANK_OK) return r;
    e->loaded[name] = true;
    bool all = true;
    for (auto& kv : e->loaded) all = all && kv.second;
    // W_ov / wkT are functions of the decoder self-attention v / o and the cross-attention k weights: a reload of one of those into a
    // live engine (checkpoint swap) must refresh them, or the T = 1 path would mix old and new weights
    const bool feeds_derived = name.rfind("decoder.", 0) == 0 &&
                               (name.find("SelfAttention.v.") != std::string::npos || name.find("SelfAttention.o.") != std::string::npos ||
                                name.find("EncDecAttention.k.") != std::string::npos);
    if (all && (!e->weights_ready || feeds_derived)) {
        if (e->weights_ready) CU_OK(cudaDeviceSynchronize());   // nothing may still be reading the old derived weights
        e->weights_ready = false;
        RET_IF(derive_weights(e));
    }
    e->weights_ready = all;
    return B200RANK_OK;
}

extern "C" int b200rank_missing_tensors(b200rank_engine* e, char* buf, int buflen) {
    if (!e) return set_error(B200RANK_ERR_ARG, "null engine");
    int n = 0;
    std::string s;
    for (auto& kv : e->loaded)
        if (!kv.second) { if (n++) s += ","; s += kv.first; }
    if (buf && buflen > 0) { strncpy(buf, s.c_str(), buflen - 1); buf[buflen - 1] = 0; }
    return n;
}
extern "C" int b200rank_weights_blob(b200rank_engine* e, void** device_ptr, size_t* nbytes) {
    if (!e || !device_ptr || !nbytes) return set_error(B200RANK_ERR_ARG, "null argument");
    *device_ptr = e->arena; *nbytes = e->arena_bytes;
    return B200RANK_OK;
}
extern "C" int b200rank_mark_weights_loaded(b200rank_engine* e) {
    if (!e) return set_error(B200RANK_ERR_ARG, "null engine");
    for (auto& kv : e->loaded) kv.second = true;
    e->weights_ready = true;
    return B200RANK_OK;
}

// ------------------------------------------------------------------ forward pieces
static int k_embed(b200rank_engine* e, const int* ids, float* x, int n) {
    if (n <= 0) return B200RANK_OK;
    prof_begin(e, "embed"); launch_k(embed_kernel, dim3((n + 7) / 8), dim3(256), 0, e->stream, ids, e->emb, x, n, e->d, e->V);
    return post_launch(e, "embed");
}
// x = E[ids], h = T5LayerNorm(x; w) in one pass over the rows (B200RANK_FUSE_EMBED_NORM=0: the two kernels, bit-identical)
static int k_rmsnorm(b200rank_engine* e, const float* x, const float* w, bf16* h, int n);
static int k_embed_norm(b200rank_engine* e, const int* ids, float* x, const float* w, bf16* h, int n) {
    if (n <= 0) return B200RANK_OK;
    static int fuse = -1;
    if (fuse < 0) fuse = (getenv("B200RANK_FUSE_EMBED_NORM") && atoi(getenv("B200RANK_FUSE_EMBED_NORM")) == 0) ? 0 : 1;
    if (!fuse) {
        RET_IF(k_embed(e, ids, x, n));
        return k_rmsnorm(e, x, w, h, n);
    }
    const int grid = (n + 7) / 8;
    prof_begin(e, "embed_norm");
    if (e->d <= 1024) launch_k(embed_norm_kernel<8>, dim3(grid), dim3(256), 0, e->stream, ids, (const float*)e->emb, x, w, h, n, e->d, e->V, e->cfg.layer_norm_eps);
    else if (e->d <= 2048) launch_k(embed_norm_kernel<16>, dim3(grid), dim3(256), 0, e->stream, ids, (const float*)e->emb, x, w, h, n, e->d, e->V, e->cfg.layer_norm_eps);
    else launch_k(embed_norm_kernel<32>, dim3(grid), dim3(256), 0, e->stream, ids, (const float*)e->emb, x, w, h, n, e->d, e->V, e->cfg.layer_norm_eps);
    return post_launch(e, "embed_norm");
}
static int k_rmsnorm(b200rank_engine* e, const float* x, const float* w, bf16* h, int n) {
    if (n <= 0) return B200RANK_OK;
    const int grid = (n + 7) / 8;
    static int rev = -1;
    if (rev < 0) rev = (getenv("B200RANK_RMSNORM_REV") && atoi(getenv("B200RANK_RMSNORM_REV")) == 0) ? 0 : 1;
    // profile label: the encoder-sized launches (HBM-bound) apart from the 100-row decoder launches (launch-latency-bound)
    const char* label = n >= 4096 ? "rmsnorm" : "rmsnorm_small";
    prof_begin(e, label);
    if (e->d <= 1024) launch_k(rmsnorm_kernel<8>, dim3(grid), dim3(256), 0, e->stream, x, w, h, n, e->d, e->cfg.layer_norm_eps, rev);
    else if (e->d <= 2048) launch_k(rmsnorm_kernel<16>, dim3(grid), dim3(256), 0, e->stream, x, w, h, n, e->d, e->cfg.layer_norm_eps, rev);
    else launch_k(rmsnorm_kernel<32>, dim3(grid), dim3(256), 0, e->stream, x, w, h, n, e->d, e->cfg.layer_norm_eps, rev);
    return post_launch(e, label);
}

// x += A.W^T (in-L2 reduce-add epilogue), then h = bf16(T5LayerNorm(x) * norm_w) as a kernel of its own: fusing the norm into the GEMM
// (row-owning units re-reading their finished rows from L2) measured slower on the B200 in round 1 and was removed in round 2.
static int gemm_resid_then_norm(b200rank_engine* e, const bf16* A, int lda, int a_rows, const bf16* W, int ldw, int w_rows, int M, int K,
                                float* x, const float* norm_w, bf16* h) {
    const int d = e->d;
    RET_IF(gemm(e, A, lda, a_rows, W, ldw, w_rows, M, d, K, EPI_RESID_F32, x, d));
    return k_rmsnorm(e, x, norm_w, h, M);
}

// Encoder attention dispatch. mode: 0 = default for the build, 1 = mma.sync 64-query tiles (any length), 5 = round-1 persistent tcgen05
// kernel (len <= 192), 8 = round-2 persistent tcgen05 kernel (len <= 192; 16 softmax warps, one unshifted pass: attention_tc5.cuh).
// B200RANK_ATTN=tiled|tc2|tc5 overrides mode 0.
static int attn_default_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* s = getenv("B200RANK_ATTN");
        // default: tc5 for documents of <= 192 tokens (launch_enc_attention falls back to the mma.sync tiles above that):
        // 1.55 vs 2.0 ms (tc2) vs 2.3 ms (tiles) per 100 documents at S = 184 (profiles/r02_bench_attn_ab.txt)
        mode = !s ? 8 : (!strcmp(s, "tc2") ? 5 : (!strcmp(s, "tc5") ? 8 : 1));
    }
    return mode;
}
static int device_sm_count() {   // of the current device (test entry points without an engine; engines carry num_sms)
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : 1;
}
static int launch_enc_attention(b200rank_engine* e, const bf16* qkv, int ld, uint64_t qkv_rows, int inner, const int* d_cu, int nd,
                                int maxlen, int H, const float* bias, bf16* out, int ldo, cudaStream_t st, int mode, int minlen = 0,
                                const float* bias_wide = nullptr, uint8_t* bad_map = nullptr) {
    if (mode == 0) mode = attn_default_mode();
    if (mode == 8 && H > AttnTc5Cfg<3>::kMaxHeads) mode = 1;   // the bias windows of all heads must fit in shared memory (71 heads; T5-11B has 64)
    const bool persistent = (mode == 5 || mode == 8);
    if (!persistent) mode = 1;
    // Mixed batch under the default mode: documents of <= 192 tokens still get the tcgen05 kernel (it walks only those), the longer
    // ones the mma.sync tiles (which skip the short ones) — the kernel is chosen per document, so a document's result does not
    // depend on what it is batched with.
    const bool mixed = persistent && maxlen > 192;
    if (persistent && minlen > 192) mode = 1;   // known: no document qualifies for the tcgen05 kernel
    if (mode == 8) {
        // persistent tcgen05 kernel, 16 softmax warps / one-pass softmax (attention_tc5.cuh)
        CUtensorMap local;
        const CUtensorMap* tm = &local;
        if (e) RET_IF(engine_tmap(e, qkv, qkv_rows, (uint64_t)ld, (uint64_t)ld, 64, 0, &tm));
        else RET_IF(make_tmap(&local, qkv, qkv_rows, (uint64_t)ld, (uint64_t)ld, 64, 0));
        static SmemOptIn opt_in_tc5;
        auto kern5 = enc_attention_tc5_kernel<3>;
        CU_OK(opt_in_tc5.raise(kern5, 227 * 1024));
        const int n_items = nd * H;
        if (e) prof_begin(e, "enc_attention_tc5");
        const int sm_count = e ? e->num_sms : device_sm_count();
        const float* bw = e ? e->bias_enc_wide : bias_wide;
        uint8_t* bm = e ? e->attn_bad_map : bad_map;
        if (!bw || !bm) return set_error(B200RANK_ERR_ARG, "tc5 attention needs the widened bias table and the row fix-up map");
        CU_OK(launch_k(kern5, dim3(std::min(n_items, sm_count)), dim3(kAttn5Threads), AttnTc5Cfg<3>::smem_bytes(H), st, *tm, qkv, ld,
                       inner, d_cu, bw, out, ldo, H, n_items, 192, bm));
        if (e) RET_IF(post_launch(e, "enc_attention_tc5"));
        if (!mixed) return B200RANK_OK;
        if (e) prof_begin(e, "enc_attention");
        launch_k(enc_attention_kernel, dim3(dim3((maxlen + 63) / 64, H, nd)), dim3(128), 0, st, qkv, ld, inner, d_cu, bias, out, ldo, 192);
        return e ? post_launch(e, "enc_attention") : B200RANK_OK;
    }
    if (mode == 5) {
        // persistent tcgen05 kernel: one CTA per SM walks the (document, head) items
        CUtensorMap local;
        const CUtensorMap* tm = &local;
        if (e) RET_IF(engine_tmap(e, qkv, qkv_rows, (uint64_t)ld, (uint64_t)ld, 64, 0, &tm));
        else RET_IF(make_tmap(&local, qkv, qkv_rows, (uint64_t)ld, (uint64_t)ld, 64, 0));
        static SmemOptIn opt_in_two_pass;
        auto kern = enc_attention_tc2_kernel<3>;
        CU_OK(opt_in_two_pass.raise(kern, 227 * 1024));
        const int n_items = nd * H;
        static int spin = -1;
        if (spin < 0) spin = getenv("B200RANK_ATTN_SPIN") ? atoi(getenv("B200RANK_ATTN_SPIN")) : 0;
        if (e) prof_begin(e, "enc_attention_tc2");
        static int attn_sms = -1;   // B200RANK_ATTN_SMS=n: cap the persistent grid (leaves SMs to the decoder stream of the other query in flight)
        if (attn_sms < 0) attn_sms = getenv("B200RANK_ATTN_SMS") ? std::max(1, atoi(getenv("B200RANK_ATTN_SMS"))) : 0;
        const int sm_count = e ? e->num_sms : device_sm_count();
        const int grid_cap = attn_sms > 0 ? std::min(attn_sms, sm_count) : sm_count;
        CU_OK(launch_k(kern, dim3(std::min(n_items, grid_cap)), dim3(kAttnTcThreads), AttnTc2Cfg<3>::smem_bytes(H), st, *tm, inner, d_cu, bias,
                       out, ldo, H, n_items, spin, 192));
        if (e) RET_IF(post_launch(e, "enc_attention_tc2"));
        if (!mixed) return B200RANK_OK;
        if (e) prof_begin(e, "enc_attention");
        launch_k(enc_attention_kernel, dim3(dim3((maxlen + 63) / 64, H, nd)), dim3(128), 0, st, qkv, ld, inner, d_cu, bias, out, ldo, 192);
        return e ? post_launch(e, "enc_attention") : B200RANK_OK;
    }
    if (e) prof_begin(e, "enc_attention");
    launch_k(enc_attention_kernel, dim3(dim3((maxlen + 63) / 64, H, nd)), dim3(128), 0, st, qkv, ld, inner, d_cu, bias, out, ldo, 0);
    return e ? post_launch(e, "enc_attention") : B200RANK_OK;
}

// First feed-forward GEMM: gated-gelu (T5 v1.1 / Flan-T5, modeling_t5.py:115-132: gelu_new(wi_0 h) * wi_1 h, tile-interleaved
// weights, product in the epilogue) or plain relu (T5 v1.0 / monoT5, modeling_t5.py:88-103: relu(wi h)).
static int ffn_in(b200rank_engine* e, const bf16* h, int h_rows, const bf16* wi, int M, bf16* g) {
    const int d = e->d, F = e->F;
    if (e->gated) return gemm(e, h, d, h_rows, wi, d, 2 * F, M, 2 * F, d, EPI_GATED_BF16, g, F);
    return gemm(e, h, d, h_rows, wi, d, F, M, F, d, EPI_RELU_BF16, g, F);
}

// Generic-width attention (attention_wide.cuh): every attention site of a d_kv = 128 model.
template <int KIND>
static int launch_attention_wide(b200rank_engine* e, const char* label, const bf16* q, int ldq, int T, const bf16* kv, size_t ldkv, int k_off,
                                 int v_off, const int* cu, const float* bias, int bias_len, bf16* out, int ldo, int q_tiles, int nd) {
    if (e->dkv != 128) return set_error(B200RANK_ERR_ARG, "attention_wide: d_kv=%d not instantiated", e->dkv);
    static SmemOptIn smem_opt_in;
    auto kern = attention_wide_kernel<128, KIND>;
    CU_OK(smem_opt_in.raise(kern, AttnWideCfg<128>::kSmemBytes));
    prof_begin(e, label);
    CU_OK(launch_k(kern, dim3(q_tiles, e->H, nd), dim3(128), AttnWideCfg<128>::kSmemBytes, e->stream, q, ldq, T, kv, ldkv, k_off, v_off, cu, bias,
                   bias_len, out, ldo));
    return post_launch(e, label);
}

// Encoder over the staged batch + stacked cross-attention K|V projection of its output.
static int run_encoder(b200rank_engine* e, bool need_ckv = true) {
    const int n = e->staged_tokens, nd = e->staged_docs;
    const int d = e->d, I = e->inner, F = e->F, Tk = e->cap_tokens;
    RET_IF(k_embed_norm(e, e->d_ids, e->x, e->enc[0].ln1, e->h, n));
    for (int l = 0; l < e->Le; ++l) {
        const LayerW& w = e->enc[l];
        // h = norm1(x) was produced by the previous layer's last GEMM (or the line above for layer 0)
        RET_IF(gemm(e, e->h, d, Tk, w.wqkv, d, 3 * I, n, 3 * I, d, EPI_BF16, e->qkv, 3 * I));
        if (e->dkv == 64)
            RET_IF(launch_enc_attention(e, e->qkv, 3 * I, (uint64_t)Tk, I, e->d_cu_cur, nd, e->staged_maxlen, e->H, e->bias_enc, e->ao, I, e->stream, 0, e->staged_minlen));
        else
            RET_IF(launch_attention_wide<ATT_ENC>(e, "enc_attention_wide", e->qkv, 3 * I, 0, e->qkv, (size_t)3 * I, I, 2 * I, e->d_cu_cur, e->bias_enc,
                                                  kAttnBiasLen, e->ao, I, (e->staged_maxlen + 63) / 64, nd));
        RET_IF(gemm_resid_then_norm(e, e->ao, I, Tk, w.wo, I, d, n, I, e->x, w.ln2, e->h));
        RET_IF(ffn_in(e, e->h, Tk, w.wi, n, e->g));
        const bool last = (l + 1 == e->Le);  // the final layer norm lands in the slot buffer the decoder reads
        RET_IF(gemm_resid_then_norm(e, e->g, F, Tk, w.wff, F, d, n, F, e->x, last ? e->enc_final_ln : e->enc[l + 1].ln1,
                                    last ? e->enc_out_cur : e->h));
    }
    if (!need_ckv) return B200RANK_OK;  // re-associated T=1 decoder attends against the encoder output itself
    const int NC = e->Ld * 2 * I;
    RET_IF(gemm(e, e->enc_out_cur, d, Tk, e->wckv, d, NC, n, NC, d, EPI_BF16, e->ckv, NC));
    return B200RANK_OK;
}

// T = 1 and short documents: the decoder's cross-attention runs in re-associated form (cross_ctx_t1.cuh) and the encoder
// skips the stacked cross-K|V projection. B200RANK_DEC_REASSOC=0 selects the reference-shaped path (K/V GEMM + attention).
static bool reassoc_t1_possible(const b200rank_engine* e, int maxlen) {
    static int pref = -1;
    if (pref < 0) pref = (getenv("B200RANK_DEC_REASSOC") && atoi(getenv("B200RANK_DEC_REASSOC")) == 0) ? 0 : 1;
    return pref && maxlen <= 240 && !e->debug_simt && e->d % kCtxKC == 0 && e->dkv == 64;  // 240: 3-stage ring fits 227 KB
}
static bool use_reassoc_t1(const b200rank_engine* e, int T) { return T == 1 && reassoc_t1_possible(e, e->staged_maxlen); }

static bool cross_split_off() {   // read per call (not cached): tests flip B200RANK_CROSS_SPLIT in-process for the A/B
    const char* v = getenv("B200RANK_CROSS_SPLIT");
    return v && atoi(v) == 0;
}

// Many decoder positions (qlm labels, long prefixes): tensor-core attention kernels (attention_dec.cuh); the CUDA-core kernels
// remain for the 1-4 position prefixes of yes_no / generation. B200RANK_DEC_ATTN=simt forces the CUDA-core kernels everywhere.
static bool dec_attn_mma(int T) {
    static int pref = -1;
    if (pref < 0) pref = (getenv("B200RANK_DEC_ATTN") && !strcmp(getenv("B200RANK_DEC_ATTN"), "simt")) ? 0 : 1;
    return pref && T > 4;
}

// ---- decoder projections: tcgen05 GEMM (+ separate T5LayerNorm launch) or, for a handful of rows, skinny_gemv.cuh
// OFF by default (B200RANK_SKINNY=1 enables it): measured on a setwise compare (tests/gpu_prof_setwise.py) the fused kernels only
// save the T5LayerNorm launches (9.4 vs 10.5 ms of profiled GPU time per compare; a skinny launch costs as much as a 2-row tcgen05
// GEMM launch, both are latency-bound), while they give up a property the batched sort drivers rely on: with the GEMM path a row's
// result is bit-identical whatever else is in the batch, so rerank_many / level-parallel heaps reproduce rerank() exactly.
static bool use_skinny(const b200rank_engine* e, int R) {
    static int pref = -1;
    if (pref < 0) pref = (getenv("B200RANK_SKINNY") && atoi(getenv("B200RANK_SKINNY")) == 1) ? 1 : 0;
    return pref && R <= kSkinnyMaxRows && !e->debug_simt && e->d % 8 == 0 && e->F % 8 == 0 && e->inner % 8 == 0;
}

template <int EPI>
static int launch_skinny(b200rank_engine* e, const char* label, const float* x, const bf16* a, int lda, const float* ln_w, const bf16* W,
                         int ldw, int R, int n_out, int K, void* out, int ldo) {
    static SmemOptIn smem_opt_in;
    CU_OK(smem_opt_in.raise(skinny_gemv_kernel<EPI>, kSkinnyMaxRows * 10240 * 2));
    if ((size_t)R * K * 2 > (size_t)kSkinnyMaxRows * 10240 * 2) return set_error(B200RANK_ERR_CAPACITY, "skinny GEMV: K=%d too long", K);
    prof_begin(e, label);
    const int grid = std::min((n_out + 7) / 8, 8 * e->num_sms);
    CU_OK(launch_k(skinny_gemv_kernel<EPI>, dim3(grid), dim3(kSkinnyThreads), (size_t)R * K * 2, e->stream, x, a, lda, ln_w, e->cfg.layer_norm_eps, W, ldw,
                   R, n_out, K, out, ldo));
    return post_launch(e, label);
}

// out_bf16[R, N] = T5LayerNorm(xd; ln_w) . W^T   (W: [N, d])
// normed: hd already holds the normed rows (block 0, where the embedding kernel produced them)
static int dec_norm_proj(b200rank_engine* e, const float* ln_w, const bf16* W, int N, int R, bf16* out, int ldo, bool normed = false) {
    const int d = e->d;
    if (use_skinny(e, R)) return launch_skinny<SK_BF16>(e, "skinny_norm_proj", e->xd, nullptr, 0, ln_w, W, d, R, N, d, out, ldo);
    if (!normed) RET_IF(k_rmsnorm(e, e->xd, ln_w, e->hd, R));
    return gemm(e, e->hd, d, e->cap_rows, W, d, N, R, N, d, EPI_BF16, out, ldo);
}

// xd[R, d] += A[R, K] . W^T   (W: [d, K])
static int dec_resid_proj(b200rank_engine* e, const bf16* A, int lda, const bf16* W, int K, int R) {
    const int d = e->d;
    if (use_skinny(e, R)) return launch_skinny<SK_RESID_F32>(e, "skinny_resid_proj", nullptr, A, lda, nullptr, W, K, R, d, K, e->xd, d);
    return gemm(e, A, lda, e->cap_rows, W, K, d, R, d, K, EPI_RESID_F32, e->xd, d);
}

// gd[R, F] = act(T5LayerNorm(xd; ln2) . wi^T)  (gated-gelu over the tile-interleaved wi_0 | wi_1 packing, or relu)
static int dec_norm_ffn_in(b200rank_engine* e, const float* ln_w, const bf16* wi, int R) {
    const int d = e->d, F = e->F;
    if (use_skinny(e, R)) {
        if (e->gated) return launch_skinny<SK_GATED_BF16>(e, "skinny_norm_ffn_in", e->xd, nullptr, 0, ln_w, wi, d, R, F, d, e->gd, F);
        return launch_skinny<SK_RELU_BF16>(e, "skinny_norm_ffn_in", e->xd, nullptr, 0, ln_w, wi, d, R, F, d, e->gd, F);
    }
    RET_IF(k_rmsnorm(e, e->xd, ln_w, e->hd, R));
    return ffn_in(e, e->hd, e->cap_rows, wi, R, e->gd);
}

// Cross-attention of 1-4 decoder positions per document (generation / likelihood prefixes, KV-cached greedy steps). The keys of every
// (document, head) are ALWAYS split 8 ways (CTA z takes the 128-key chunks z, z+8, ...; flash-decoding partials + exact log-sum-exp
// merge): a setwise compare (1 document, 1.5 k keys) otherwise runs on 16 CTAs of a 148-SM GPU, 100 us per layer. The split is a
// function of the document alone — not of how many documents share the pass, nor of how many positions are in flight — so a row's
// result stays bit-identical across batch compositions (what lets rerank_many / the level-parallel heaps reproduce rerank()
// exactly) and between a re-run prefix and a cached step. The last launch's post_launch is left to the caller.
static int cross_attention_short(b200rank_engine* e, int doc0, int nd, int T, size_t ldkv, int k_off, int v_off) {
    const int I = e->inner;
    if (!cross_split_off()) {
        const int group = (int)(e->xattn_partial_bytes / ((size_t)e->H * kXAttnSplit * T * 66 * sizeof(float)));
        for (int g0 = 0; g0 < nd; g0 += group) {
            const int gn = std::min(group, nd - g0);
            if (g0) prof_begin(e, "cross_attention");
            cross_attention_kernel<4, 128><<<dim3(e->H, gn, kXAttnSplit), 128, 0, e->stream>>>(e->qd + (size_t)g0 * T * I, I, T, e->ckv, ldkv, k_off, v_off,
                                                                                                e->d_cu_cur + doc0 + g0, e->aod + (size_t)g0 * T * I, I, e->xattn_partial);
            RET_IF(post_launch(e, "cross_attention"));
            prof_begin(e, "cross_attention_combine");
            cross_attention_combine_kernel<<<dim3(e->H, gn), 64, 0, e->stream>>>(e->xattn_partial, kXAttnSplit, T, e->aod + (size_t)g0 * T * I, I);
            if (g0 + gn < nd) RET_IF(post_launch(e, "cross_attention_combine"));
        }
    } else {
        cross_attention_kernel<4, 128><<<dim3(e->H, nd), 128, 0, e->stream>>>(e->qd, I, T, e->ckv, ldkv, k_off, v_off, e->d_cu_cur + doc0, e->aod, I);
    }
    return B200RANK_OK;
}

// ---- greedy decoding with a self-attention K/V cache (generation/utils.py:2762-2804 runs one cached step per new token)
// B200RANK_KV_CACHE=0 selects the cache-less loop (the growing prefix re-run per step); read per call so that the GPU tests can
// compare the two in one process.
static bool kv_cache_off() {
    const char* v = getenv("B200RANK_KV_CACHE");
    return v && atoi(v) == 0;
}
static bf16* kvc_layer(const b200rank_engine* e, int l) { return e->kvc + (size_t)l * e->kvc_rows * 3 * e->inner; }

// k | v of the prefix rows (qkv workspace, row = doc * T + t) -> cache rows doc * row_stride + t (columns [inner, 3 inner))
__global__ void kv_cache_fill_kernel(const bf16* __restrict__ qkv, int ld, int inner, int T, bf16* __restrict__ cache, int row_stride, int n_rows) {
    pdl_trigger();
    pdl_wait();
    const int r = blockIdx.x;
    if (r >= n_rows) return;
    const uint4* src = reinterpret_cast<const uint4*>(qkv + (size_t)r * ld + inner);
    uint4* dst = reinterpret_cast<uint4*>(cache + ((size_t)(r / T) * row_stride + (r % T)) * ld + inner);
    for (int i = threadIdx.x; i < 2 * inner / 8; i += blockDim.x) dst[i] = src[i];
}

// Decoder over documents [doc0, doc0+nd) with T positions each; dec ids in d_dec_ids[nd*T].
// Leaves the final-normed hidden states in hd[nd*T, d].
// kv_stride > 0 (first step of a KV-cached greedy call): the self-attention k | v rows of the T prefix positions are also written to
// the cache, row stride kv_stride per document, for run_decoder_cached_step to attend to.
static int run_decoder(b200rank_engine* e, int doc0, int nd, int T, int kv_stride = 0) {
    const int R = nd * T;
    const int d = e->d, I = e->inner, F = e->F, cap = e->cap_rows;
    if (R > cap) return set_error(B200RANK_ERR_CAPACITY, "decoder rows %d exceed capacity %d", R, cap);
    if (T > e->cap_T) return set_error(B200RANK_ERR_CAPACITY, "decoder length %d exceeds max_dec_len %d", T, e->cap_T);
    RET_IF(k_embed_norm(e, e->d_dec_ids, e->xd, e->dec[0].ln1, e->hd, R));   // block 0's first norm rides on the embedding gather
    const size_t ldkv = (size_t)e->Ld * 2 * I;
    const int max_len = e->staged_maxlen;
    const bool reassoc = use_reassoc_t1(e, T);
    if (reassoc) {
        static SmemOptIn smem_opt_in;
        CU_OK(smem_opt_in.raise(cross_ctx_t1_kernel, cross_ctx_smem_bytes(240)));
    }
    for (int l = 0; l < e->Ld; ++l) {
        const LayerW& w = e->dec[l];
        if (T == 1) {
            // one decoder position: softmax over a single key is 1, so the block is x += W_o W_v norm(x) (derive_weights).
            // Input and output are both xd, so the norm stays a kernel of its own (a fused norm would race with the adds).
            if (l > 0) RET_IF(k_rmsnorm(e, e->xd, w.ln1, e->hd, R));
            // the fused W_o W_v product never forms k and v: project them for the cache from the same normed rows (the k | v rows
            // of Wqkv; a GEMM row does not depend on the tile shape, so these are the values a re-run prefix would compute)
            if (kv_stride)
                RET_IF(gemm(e, e->hd, d, cap, w.wqkv + (size_t)I * d, d, 2 * I, R, 2 * I, d, EPI_BF16, kvc_layer(e, l) + I, kv_stride * 3 * I));
            RET_IF(dec_resid_proj(e, e->hd, d, w.wov, d, R));
        } else {
            RET_IF(dec_norm_proj(e, w.ln1, w.wqkv, 3 * I, R, e->qkvd, 3 * I, /*normed=*/l == 0));
            // k | v of the prefix go to the cache: inside the CUDA-core attention kernel for short prefixes, by a copy otherwise
            if (kv_stride && (e->dkv != 64 || dec_attn_mma(T))) {
                prof_begin(e, "kv_cache_fill");
                launch_k(kv_cache_fill_kernel, dim3(R), dim3(128), 0, e->stream, (const bf16*)e->qkvd, 3 * I, I, T, kvc_layer(e, l), kv_stride, R);
                RET_IF(post_launch(e, "kv_cache_fill"));
            }
            if (e->dkv != 64) {
                RET_IF(launch_attention_wide<ATT_DEC_SELF>(e, "dec_self_attention_wide", e->qkvd, 3 * I, T, e->qkvd, (size_t)3 * I, I, 2 * I, nullptr,
                                                           e->bias_dec, kAttnRelClamp + 1, e->aod, I, (T + 63) / 64, nd));
            } else if (dec_attn_mma(T)) {
                prof_begin(e, "dec_self_attention_mma");
                launch_k(dec_attention_mma_kernel<true>, dim3((T + 63) / 64, e->H, nd), dim3(128), 0, e->stream, e->qkvd, 3 * I, T, e->qkvd, (size_t)3 * I, I, 2 * I,
                         (const int*)nullptr, e->bias_dec, kAttnRelClamp + 1, e->aod, I);
                RET_IF(post_launch(e, "dec_self_attention_mma"));
            } else {
                prof_begin(e, "dec_self_attention"); dec_self_attention_kernel<<<dim3(e->H, nd), 128, 0, e->stream>>>(e->qkvd, 3 * I, I, T, 0, T, e->bias_dec, kAttnRelClamp + 1, e->aod, I,
                                                                                                          kv_stride ? kvc_layer(e, l) : nullptr, kv_stride);
                RET_IF(post_launch(e, "dec_self_attention"));
            }
            RET_IF(dec_resid_proj(e, e->aod, I, w.wo, I, R));
        }
        RET_IF(dec_norm_proj(e, w.ln_c, w.wq_c, I, R, e->qd, I));
        if (reassoc) {
            // cross_ctx_t1.cuh: scores = (W_k,h^T q_h) . e_j and out_h = W_v,h (sum_j p_j e_j): no K/V projection of the encoder
            const int HD = e->H * d, dcap = (int)align_up(e->cap_docs, 128);
            RET_IF(gemm(e, e->qd, I, cap, w.wkT, 64, HD, R, HD, 64, EPI_BF16, e->qp, HD, 0, /*n_per_batch=*/d, /*a_cols=*/I));
            const int s_pad = (max_len + 15) & ~15;
            prof_begin(e, "cross_ctx_t1");
            launch_k(cross_ctx_t1_kernel, dim3(nd), dim3(kCtxThreads), cross_ctx_smem_bytes(s_pad), e->stream, e->qp, e->enc_out_cur, e->d_cu_cur + doc0, e->ctxb, e->H, d, s_pad);
            RET_IF(post_launch(e, "cross_ctx_t1"));
            const bf16* wv_l = e->wckv + ((size_t)l * 2 * I + I) * d;
            RET_IF(gemm(e, e->ctxb, HD, dcap, wv_l, d, I, R, I, d, EPI_BF16, e->aod, I, /*force_bn=*/32, /*n_per_batch=*/64, /*a_cols=*/HD));
            RET_IF(dec_resid_proj(e, e->aod, I, w.wo_c, I, R));
            RET_IF(dec_norm_ffn_in(e, w.ln2, w.wi, R));
            RET_IF(dec_resid_proj(e, e->gd, F, w.wff, F, R));
            continue;
        }
        const int k_off = l * 2 * I, v_off = l * 2 * I + I;
        if (e->dkv != 64) {
            RET_IF(launch_attention_wide<ATT_CROSS>(e, "cross_attention_wide", e->qd, I, T, e->ckv, ldkv, k_off, v_off, e->d_cu_cur + doc0, nullptr, 0,
                                                    e->aod, I, (T + 63) / 64, nd));
            RET_IF(dec_resid_proj(e, e->aod, I, w.wo_c, I, R));
            RET_IF(dec_norm_ffn_in(e, w.ln2, w.wi, R));
            RET_IF(dec_resid_proj(e, e->gd, F, w.wff, F, R));
            continue;
        }
        if (dec_attn_mma(T)) {
            prof_begin(e, "cross_attention_mma");
            launch_k(dec_attention_mma_kernel<false>, dim3((T + 63) / 64, e->H, nd), dim3(128), 0, e->stream, e->qd, I, T, e->ckv, ldkv, k_off, v_off,
                     (const int*)(e->d_cu_cur + doc0), (const float*)nullptr, 0, e->aod, I);
            RET_IF(post_launch(e, "cross_attention_mma"));
            RET_IF(dec_resid_proj(e, e->aod, I, w.wo_c, I, R));
            RET_IF(dec_norm_ffn_in(e, w.ln2, w.wi, R));
            RET_IF(dec_resid_proj(e, e->gd, F, w.wff, F, R));
            continue;
        }
        prof_begin(e, "cross_attention");
        if (T == 1 && e->H % 4 == 0 && max_len <= 256)
            cross_attention_t1_kernel<8><<<dim3(e->H / 4, nd), 128, 0, e->stream>>>(e->qd, I, e->ckv, ldkv, k_off, v_off, e->d_cu_cur + doc0, e->aod, I);
        else if (T == 1 && e->H % 4 == 0 && max_len <= 2048)
            cross_attention_t1_kernel<64><<<dim3(e->H / 4, nd), 128, 0, e->stream>>>(e->qd, I, e->ckv, ldkv, k_off, v_off, e->d_cu_cur + doc0, e->aod, I);
        else if (T <= 4) {
            RET_IF(cross_attention_short(e, doc0, nd, T, ldkv, k_off, v_off));
        }
        else
            cross_attention_kernel<40, 64><<<dim3(e->H, nd), 128, 0, e->stream>>>(e->qd, I, T, e->ckv, ldkv, k_off, v_off, e->d_cu_cur + doc0, e->aod, I);
        RET_IF(post_launch(e, "cross_attention"));
        RET_IF(dec_resid_proj(e, e->aod, I, w.wo_c, I, R));
        RET_IF(dec_norm_ffn_in(e, w.ln2, w.wi, R));
        RET_IF(dec_resid_proj(e, e->gd, F, w.wff, F, R));
    }
    RET_IF(k_rmsnorm(e, e->xd, e->dec_final_ln, e->hd, R));
    return B200RANK_OK;
}

// One NEW decoder position per document (position pos >= 1; its token in d_dec_ids[nd]) against the self-attention K/V cache that
// run_decoder(kv_stride) and the earlier steps filled. Per layer: the fused q|k|v projection of the new row lands in the cache row
// of (document, pos) — the output tensor map strides by a document's kv_stride rows — the generalised dec_self_attention_kernel
// attends that row's q to the k/v of positions 0..pos, and cross-attention runs with one query row per document. Every kernel and
// every per-row summation order is the one the cache-less loop uses for the same position, so the generated tokens are identical.
// Leaves the final-normed hidden states of the new rows in hd[nd, d].
static int run_decoder_cached_step(b200rank_engine* e, int doc0, int nd, int pos, int kv_stride) {
    const int R = nd;
    const int I = e->inner, F = e->F;
    const size_t ldkv = (size_t)e->Ld * 2 * I;
    RET_IF(k_embed_norm(e, e->d_dec_ids, e->xd, e->dec[0].ln1, e->hd, R));
    for (int l = 0; l < e->Ld; ++l) {
        const LayerW& w = e->dec[l];
        bf16* cache = kvc_layer(e, l);
        RET_IF(dec_norm_proj(e, w.ln1, w.wqkv, 3 * I, R, cache + (size_t)pos * 3 * I, kv_stride * 3 * I, /*normed=*/l == 0));
        prof_begin(e, "dec_self_attention");
        dec_self_attention_kernel<<<dim3(e->H, nd), 128, 0, e->stream>>>(cache, 3 * I, I, kv_stride, pos, 1, e->bias_dec, kAttnRelClamp + 1, e->aod, I);
        RET_IF(post_launch(e, "dec_self_attention"));
        RET_IF(dec_resid_proj(e, e->aod, I, w.wo, I, R));
        RET_IF(dec_norm_proj(e, w.ln_c, w.wq_c, I, R, e->qd, I));
        prof_begin(e, "cross_attention");
        RET_IF(cross_attention_short(e, doc0, nd, 1, ldkv, l * 2 * I, l * 2 * I + I));
        RET_IF(post_launch(e, "cross_attention"));
        RET_IF(dec_resid_proj(e, e->aod, I, w.wo_c, I, R));
        RET_IF(dec_norm_ffn_in(e, w.ln2, w.wi, R));
        RET_IF(dec_resid_proj(e, e->gd, F, w.wff, F, R));
    }
    RET_IF(k_rmsnorm(e, e->xd, e->dec_final_ln, e->hd, R));
    return B200RANK_OK;
}

static float logit_scale(const b200rank_engine* e) {
    return e->cfg.scale_decoder_outputs ? 1.0f / std::sqrt(static_cast<float>(e->d)) : 1.0f;
}

// Runs `enqueue` (kernel launches on e->stream with arguments that are a function of `key` alone — no copies, no synchronisation) as a
// CUDA graph: eagerly on the key's first occurrence, captured on the second, replayed from then on. One host call instead of ~280, no
// launch gaps on the device; the kernels and their order are the eager ones, so the results are bit-identical. B200RANK_DEC_GRAPH=0,
// profiling and B200RANK_DEBUG_SYNC run everything eagerly; so does a key that arrives when 256 graphs exist already.
template <class F>
static int run_as_graph(b200rank_engine* e, const std::array<int, 8>& key, F&& enqueue) {
    b200rank_engine::DecGraph* g = nullptr;
    if (e->dec_graph && !e->profiling && !e->debug_sync) {
        auto it = e->step_graphs.find(key);
        if (it != e->step_graphs.end()) g = &it->second;
        else if (e->step_graphs.size() < 256) g = &e->step_graphs[key];
    }
    if (g && g->state == 2) {
        cudaError_t er = cudaGraphLaunch(g->exec, e->stream);
        if (er != cudaSuccess) return set_error(B200RANK_ERR_CUDA, "decoder step graph launch: %s", cudaGetErrorString(er));
        e->launches += g->launches;
        return B200RANK_OK;
    }
    if (g && g->state == 1) {
        const uint64_t l0 = e->launches;
        cudaGraph_t graph = nullptr;
        cudaError_t er = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal);
        const int crc = (er == cudaSuccess) ? enqueue() : B200RANK_ERR_CUDA;
        if (er == cudaSuccess) er = cudaStreamEndCapture(e->stream, &graph);
        int rc = B200RANK_OK;
        if (crc == B200RANK_OK && er == cudaSuccess && graph && cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess) {
            g->state = 2;
            g->launches = e->launches - l0;
            er = cudaGraphLaunch(g->exec, e->stream);
            if (er != cudaSuccess) rc = set_error(B200RANK_ERR_CUDA, "decoder step graph launch: %s", cudaGetErrorString(er));
        } else {            // capture is not possible here: never try this key again, run the step the ordinary way
            g->state = 3;
            g->exec = nullptr;
            cudaGetLastError();
            e->launches = l0;
            rc = enqueue();
        }
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (g && g->state == 0) g->state = 1;
    return enqueue();
}

// Full-vocabulary row reductions (log-prob of a label / argmax / softmax gather): the register-resident single-pass kernel when the
// row fits (V <= 32768, a multiple of 4, 16 B-aligned rows), the multi-pass kernel otherwise. B200RANK_VOCAB_ROW=multipass forces the latter.
static int k_vocab_row(b200rank_engine* e, const char* what, int rows, int mode, const int* labels, const int* cols, int ncols, float* out_f, int* out_i) {
    static int regs = -1;
    if (regs < 0) regs = (getenv("B200RANK_VOCAB_ROW") && !strcmp(getenv("B200RANK_VOCAB_ROW"), "multipass")) ? 0 : 1;
    prof_begin(e, "vocab_row");
    if (regs && e->V % 4 == 0 && e->V <= 32768)
        launch_k(vocab_row_regs_kernel<8>, dim3(rows), dim3(1024), 0, e->stream, (const float*)e->logits, e->V, (size_t)e->V, mode, logit_scale(e), labels, cols, ncols, out_f, out_i);
    else
        launch_k(vocab_row_kernel, dim3(rows), dim3(256), 0, e->stream, (const float*)e->logits, e->V, (size_t)e->V, mode, logit_scale(e), labels, cols, ncols, out_f, out_i);
    return post_launch(e, what);
}

// ------------------------------------------------------------------ staging
static int check_ready(b200rank_engine* e) {
    if (!e) return set_error(B200RANK_ERR_ARG, "null engine");
    if (!e->weights_ready) return set_error(B200RANK_ERR_STATE, "weights are not fully loaded (%d tensors missing)", b200rank_missing_tensors(e, nullptr, 0));
    CU_OK(cudaSetDevice(e->device));
    if (e->slot[0].busy || e->slot[1].busy)
        return set_error(B200RANK_ERR_STATE, "pipelined batches are in flight: call b200rank_wait_yes_no() for every ticket before using the synchronous API");
    // the pinned bump buffer restarts only when nothing enqueued can still read it (b200rank_run_yes_no_staged returns without
    // synchronising); otherwise it keeps advancing as a ring — upload_ints synchronises when it wraps
    if (!e->unsynced) e->h_small_off = 0;
    e->stream = e->stream_main;
    e->enc_out_cur = e->enc_out[0];
    e->d_cu_cur = e->d_cu_slot[0];
    e->gemm_sm_cap = 0;
    return B200RANK_OK;
}

// Pack documents [d0, d1) into the pinned staging buffers and copy to the device (async on the stream).
static int stage_range(b200rank_engine* e, const int32_t* ids, const int32_t* lengths, int stride, int d0, int d1) {
    int tok = 0, maxlen = 0, minlen = INT_MAX;
    e->h_cu[0] = 0;
    for (int i = d0; i < d1; ++i) {
        const int len = lengths[i];
        memcpy(e->h_ids + tok, ids + (size_t)i * stride, (size_t)len * sizeof(int));
        tok += len;
        maxlen = std::max(maxlen, len);
        minlen = std::min(minlen, len);
        e->h_cu[i - d0 + 1] = tok;
    }
    e->staged_docs = d1 - d0; e->staged_tokens = tok; e->staged_maxlen = maxlen; e->staged_minlen = d1 > d0 ? minlen : 0;
    CU_OK(cudaMemcpyAsync(e->d_ids, e->h_ids, (size_t)tok * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    CU_OK(cudaMemcpyAsync(e->d_cu, e->h_cu, (size_t)(d1 - d0 + 1) * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    return B200RANK_OK;
}

// Greedy split of documents into device passes that respect max_tokens / max_docs / doc_limit.
static int next_group(b200rank_engine* e, const int32_t* lengths, int n_docs, int stride, int d0, int doc_limit, int* d1_out) {
    int tok = 0, d1 = d0;
    while (d1 < n_docs && (d1 - d0) < doc_limit) {
        const int len = lengths[d1];
        if (len <= 0 || len > stride) return set_error(B200RANK_ERR_ARG, "lengths[%d]=%d out of range (stride %d)", d1, len, stride);
        if (len > e->cap_tokens) return set_error(B200RANK_ERR_CAPACITY, "document %d has %d tokens > max_tokens %d", d1, len, e->cap_tokens);
        if (tok + len > e->cap_tokens) break;
        tok += len; ++d1;
    }
    *d1_out = d1;
    return B200RANK_OK;
}

// Small async upload through a pinned bump buffer (reset at the start of every API call) so the host never
// blocks behind queued kernels the way a pageable cudaMemcpyAsync would.
static int upload_ints(b200rank_engine* e, int* dst, const std::vector<int>& v) {
    if (e->h_small_off + v.size() > e->h_small_cap) {
        CU_OK(cudaStreamSynchronize(e->stream));
        if (e->stream == e->stream_main) e->unsynced = false;
        e->h_small_off = 0;
        if (v.size() > e->h_small_cap) return set_error(B200RANK_ERR_CAPACITY, "upload of %zu ints exceeds staging", v.size());
    }
    int* src = e->h_small + e->h_small_off;
    memcpy(src, v.data(), v.size() * sizeof(int));
    e->h_small_off += v.size();
    CU_OK(cudaMemcpyAsync(dst, src, v.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    return B200RANK_OK;
}

// ------------------------------------------------------------------ yes/no
static int yes_no_device(b200rank_engine* e, int yes_id, int no_id) {
    const int nd = e->staged_docs;
    if (yes_id < 0 || yes_id >= e->V || no_id < 0 || no_id >= e->V) return set_error(B200RANK_ERR_ARG, "yes/no ids out of vocabulary");
    RET_IF(run_encoder(e, /*need_ckv=*/!use_reassoc_t1(e, 1)));
    std::vector<int> dec(nd, e->cfg.pad_id);  // decoder_input_ids = [[pad]] per row, pointwise.py:102
    RET_IF(upload_ints(e, e->d_dec_ids, dec));
    RET_IF(upload_ints(e, e->d_cols, std::vector<int>{yes_id, no_id}));
    RET_IF(run_decoder(e, 0, nd, 1));
    prof_begin(e, "lm_head_cols"); launch_k(lm_head_cols_kernel, dim3(nd), dim3(64), 0, e->stream, e->hd, e->d, 1, 0, e->lm_head, e->d_cols, 2, logit_scale(e), e->small_out);
    RET_IF(post_launch(e, "lm_head_cols"));
    prof_begin(e, "yes_no_score"); launch_k(yes_no_score_kernel, dim3((nd + 127) / 128), dim3(128), 0, e->stream, e->small_out, e->small_out2, nd);
    RET_IF(post_launch(e, "yes_no_score"));
    return B200RANK_OK;
}

static int fetch_yes_no(b200rank_engine* e, float* logits2, float* scores) {
    const int nd = e->staged_docs;
    CU_OK(cudaMemcpyAsync(e->h_out, e->small_out, (size_t)nd * 2 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU_OK(cudaMemcpyAsync(e->h_out + 2 * (size_t)nd, e->small_out2, (size_t)nd * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU_OK(cudaStreamSynchronize(e->stream));
    if (e->stream == e->stream_main) e->unsynced = false;
    if (logits2) memcpy(logits2, e->h_out, (size_t)nd * 2 * sizeof(float));
    if (scores) memcpy(scores, e->h_out + 2 * (size_t)nd, (size_t)nd * sizeof(float));
    return B200RANK_OK;
}

extern "C" int b200rank_score_yes_no(b200rank_engine* e, const int32_t* ids, const int32_t* lengths, int n_docs, int stride,
                                     int yes_id, int no_id, float* logits2, float* scores) {
    RET_IF(check_ready(e));
    if (!ids || !lengths || n_docs < 0 || stride <= 0) return set_error(B200RANK_ERR_ARG, "bad arguments");
    int d0 = 0;
    while (d0 < n_docs) {
        int d1 = 0;
        RET_IF(next_group(e, lengths, n_docs, stride, d0, e->cap_docs, &d1));
        RET_IF(stage_range(e, ids, lengths, stride, d0, d1));
        RET_IF(yes_no_device(e, yes_id, no_id));
        RET_IF(fetch_yes_no(e, logits2 ? logits2 + 2 * (size_t)d0 : nullptr, scores ? scores + d0 : nullptr));
        d0 = d1;
    }
    return B200RANK_OK;
}

extern "C" int b200rank_stage(b200rank_engine* e, const int32_t* ids, const int32_t* lengths, int n_docs, int stride) {
    RET_IF(check_ready(e));
    if (!ids || !lengths || n_docs <= 0 || stride <= 0) return set_error(B200RANK_ERR_ARG, "bad arguments");
    int d1 = 0;
    RET_IF(next_group(e, lengths, n_docs, stride, 0, e->cap_docs, &d1));
    if (d1 != n_docs) return set_error(B200RANK_ERR_CAPACITY, "staged batch does not fit one device pass (%d of %d docs)", d1, n_docs);
    RET_IF(stage_range(e, ids, lengths, stride, 0, n_docs));
    CU_OK(cudaStreamSynchronize(e->stream));
    return B200RANK_OK;
}
extern "C" int b200rank_run_yes_no_staged(b200rank_engine* e, int yes_id, int no_id) {
    RET_IF(check_ready(e));
    if (e->staged_docs <= 0) return set_error(B200RANK_ERR_STATE, "nothing staged");
    e->unsynced = true;
    return yes_no_device(e, yes_id, no_id);
}
extern "C" int b200rank_fetch_yes_no(b200rank_engine* e, float* logits2, float* scores) {
    RET_IF(check_ready(e));
    if (e->staged_docs <= 0) return set_error(B200RANK_ERR_STATE, "nothing staged");
    return fetch_yes_no(e, logits2, scores);
}
extern "C" int b200rank_sync(b200rank_engine* e) {
    if (!e) return set_error(B200RANK_ERR_ARG, "null engine");
    CU_OK(cudaSetDevice(e->device));
    CU_OK(cudaStreamSynchronize(e->stream));
    if (e->stream == e->stream_main) e->unsynced = false;
    return B200RANK_OK;
}


// ------------------------------------------------------------------ two-deep pipeline (submit / wait)
// submit: pack + H2D + encoder pass on the main stream, then the decoder pass + D2H on the decoder stream; returns without
// synchronising. While the latency-bound decoder chain of batch i (a few hundred small dependent kernels on a handful of SMs)
// runs, the main stream already executes the encoder GEMMs of batch i+1 on the remaining SMs (persistent GEMM grids are capped
// at num_sms - pipe_reserve_sms while a pipeline is active).
static int check_ready_async(b200rank_engine* e) {
    if (!e) return set_error(B200RANK_ERR_ARG, "null engine");
    if (!e->weights_ready) return set_error(B200RANK_ERR_STATE, "weights are not fully loaded");
    CU_OK(cudaSetDevice(e->device));
    return B200RANK_OK;
}

extern "C" int b200rank_submit_yes_no(b200rank_engine* e, const int32_t* ids, const int32_t* lengths, int n_docs, int stride,
                                      int yes_id, int no_id, uint64_t* ticket) {
    RET_IF(check_ready_async(e));
    if (!ticket) return set_error(B200RANK_ERR_ARG, "bad arguments");
    if (yes_id < 0 || yes_id >= e->V || no_id < 0 || no_id >= e->V) return set_error(B200RANK_ERR_ARG, "yes/no ids out of vocabulary");
    const int b = static_cast<int>(e->next_ticket & 1);
    if (e->slot[b].busy) return set_error(B200RANK_ERR_STATE, "two batches are already in flight: wait for ticket %llu first", (unsigned long long)e->slot[b].ticket);
    const bool resident = (ids == nullptr);  // ids == NULL: score the batch b200rank_stage() left in device memory (no H2D)
    int tok = 0, maxlen = 0;
    if (resident) {
        if (e->staged_docs <= 0) return set_error(B200RANK_ERR_STATE, "nothing staged (b200rank_stage) for a resident submit");
        n_docs = e->staged_docs; tok = e->staged_tokens; maxlen = e->staged_maxlen;
    } else {
        if (!lengths || n_docs <= 0 || stride <= 0) return set_error(B200RANK_ERR_ARG, "bad arguments");
        // the batch must fit one device pass and take the T = 1 fast path (no stacked cross-K|V buffer is double-buffered)
        if (n_docs > e->cap_docs) return set_error(B200RANK_ERR_CAPACITY, "%d documents exceed max_docs %d", n_docs, e->cap_docs);
        for (int i = 0; i < n_docs; ++i) {
            const int len = lengths[i];
            if (len <= 0 || len > stride) return set_error(B200RANK_ERR_ARG, "lengths[%d]=%d out of range (stride %d)", i, len, stride);
            tok += len; maxlen = std::max(maxlen, len);
        }
        if (tok > e->cap_tokens) return set_error(B200RANK_ERR_CAPACITY, "%d tokens exceed max_tokens %d", tok, e->cap_tokens);
    }
    // the pipelined pass skips the stacked cross-K|V projection, so it is only valid where the decoder takes the re-associated T = 1 form
    if (!reassoc_t1_possible(e, maxlen))
        return set_error(B200RANK_ERR_ARG, "pipelined submit supports documents of at most 240 tokens on d_kv = 64 models (re-associated T = 1 decoder)");

    cudaStream_t s_enc = e->stream_main;
    if (e->unsynced) {   // a b200rank_run_yes_no_staged pass may still read the pinned bump buffer this call is about to rewrite
        CU_OK(cudaStreamSynchronize(e->stream_main));
        e->unsynced = false;
    }
    // Whatever way this function is left, the engine's "current" stream / slot pointers go back to the synchronous defaults; and if it
    // is left on an error after work was enqueued, that work is drained first — the slot is not marked busy then, so the next submit
    // would otherwise repack this slot's pinned staging under a copy that is still in flight.
    struct SubmitGuard {
        b200rank_engine* e; cudaStream_t s_enc; bool ok = false;
        ~SubmitGuard() {
            if (!ok) { cudaStreamSynchronize(s_enc); cudaStreamSynchronize(e->stream_dec); }
            e->stream = e->stream_main; e->enc_out_cur = e->enc_out[0]; e->d_cu_cur = e->d_cu_slot[0]; e->gemm_sm_cap = 0;
        }
    } submit_guard{e, s_enc};
    e->stream = s_enc;
    CU_OK(cudaStreamWaitEvent(s_enc, e->ev_done[b], 0));  // the decoder that last read this slot has finished
    if (resident) {
        if (b != 0) CU_OK(cudaMemcpyAsync(e->d_cu_slot[b], e->d_cu_slot[0], (size_t)(n_docs + 1) * sizeof(int), cudaMemcpyDeviceToDevice, s_enc));
    } else {
        // host packing into this slot's pinned staging (its previous H2D finished before the slot was waited on)
        int* hid = e->h_ids_slot[b]; int* hcu = e->h_cu_slot[b];
        hcu[0] = 0; tok = 0;
        for (int i = 0; i < n_docs; ++i) {
            memcpy(hid + tok, ids + (size_t)i * stride, (size_t)lengths[i] * sizeof(int));
            tok += lengths[i];
            hcu[i + 1] = tok;
        }
        CU_OK(cudaMemcpyAsync(e->d_ids, hid, (size_t)tok * sizeof(int), cudaMemcpyHostToDevice, s_enc));
        CU_OK(cudaMemcpyAsync(e->d_cu_slot[b], hcu, (size_t)(n_docs + 1) * sizeof(int), cudaMemcpyHostToDevice, s_enc));
        e->staged_docs = n_docs; e->staged_tokens = tok; e->staged_maxlen = maxlen;
    }
    e->staged_minlen = 0;   // not tracked here: a mixed batch (192 < maxlen <= 240) launches both attention kernels
    e->slot[b].docs = n_docs; e->slot[b].tokens = tok; e->slot[b].maxlen = maxlen;
    e->enc_out_cur = e->enc_out[b];
    e->d_cu_cur = e->d_cu_slot[b];
    e->gemm_sm_cap = std::max(2, (e->num_sms - std::max(0, e->pipe_reserve_sms)) & ~1);
    int rc = run_encoder(e, /*need_ckv=*/false);
    e->gemm_sm_cap = 0;
    if (rc == B200RANK_OK) { cudaError_t er = cudaEventRecord(e->ev_enc[b], s_enc); if (er != cudaSuccess) rc = set_error(B200RANK_ERR_CUDA, "event record: %s", cudaGetErrorString(er)); }

    // ---- decoder + head + D2H on the decoder stream
    if (rc == B200RANK_OK) {
        e->stream = e->stream_dec;
        cudaStreamWaitEvent(e->stream_dec, e->ev_enc[b], 0);
        e->h_small_off = (size_t)b * (e->h_small_cap / 2);  // per-slot half of the pinned bump buffer
        std::vector<int> dec(n_docs, e->cfg.pad_id);
        rc = upload_ints(e, e->d_dec_ids, dec);
        if (rc == B200RANK_OK) rc = upload_ints(e, e->d_cols, std::vector<int>{yes_id, no_id});
        // decoder chain + the two-column head + the score: kernels only (what a graph of this pass holds)
        auto enqueue_decoder = [&]() -> int {
            // measurement switch (never set in production: the scores are then garbage): how fast do the encoder passes run back to
            // back when no decoder chain competes for SMs at their kernel boundaries? ("decoder interference")
            static const bool skip = getenv("B200RANK_DEBUG_SKIP_DECODER") && atoi(getenv("B200RANK_DEBUG_SKIP_DECODER")) != 0;
            if (skip) return B200RANK_OK;
            int r = run_decoder(e, 0, n_docs, 1);
            if (r == B200RANK_OK) {
                prof_begin(e, "lm_head_cols");
                launch_k(lm_head_cols_kernel, dim3(n_docs), dim3(64), 0, e->stream, e->hd, e->d, 1, 0, e->lm_head, e->d_cols, 2, logit_scale(e), e->small_out);
                r = post_launch(e, "lm_head_cols");
            }
            if (r == B200RANK_OK) {
                prof_begin(e, "yes_no_score");
                launch_k(yes_no_score_kernel, dim3((n_docs + 127) / 128), dim3(128), 0, e->stream, e->small_out, e->small_out2, n_docs);
                r = post_launch(e, "yes_no_score");
            }
            return r;
        };
        if (rc == B200RANK_OK) {
            const bool try_graph = e->dec_graph && !e->profiling && !e->debug_sync && e->dec_graphs.size() < 64;
            b200rank_engine::DecGraph* g = try_graph ? &e->dec_graphs[std::make_tuple(b, n_docs, (maxlen + 15) & ~15)] : nullptr;
            if (g && g->state == 2) {
                cudaError_t er = cudaGraphLaunch(g->exec, e->stream_dec);
                if (er != cudaSuccess) rc = set_error(B200RANK_ERR_CUDA, "decoder graph launch: %s", cudaGetErrorString(er));
                e->launches += g->launches;
            } else if (g && g->state == 1) {
                const uint64_t l0 = e->launches;
                cudaGraph_t graph = nullptr;
                cudaError_t er = cudaStreamBeginCapture(e->stream_dec, cudaStreamCaptureModeThreadLocal);
                const int crc = (er == cudaSuccess) ? enqueue_decoder() : B200RANK_ERR_CUDA;
                if (er == cudaSuccess) er = cudaStreamEndCapture(e->stream_dec, &graph);
                if (crc == B200RANK_OK && er == cudaSuccess && graph && cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess) {
                    g->state = 2;
                    g->launches = e->launches - l0;
                    er = cudaGraphLaunch(g->exec, e->stream_dec);
                    if (er != cudaSuccess) rc = set_error(B200RANK_ERR_CUDA, "decoder graph launch: %s", cudaGetErrorString(er));
                } else {            // capture is not possible here: never try this key again, run the pass the ordinary way
                    g->state = 3;
                    g->exec = nullptr;
                    cudaGetLastError();
                    e->launches = l0;
                    rc = enqueue_decoder();
                }
                if (graph) cudaGraphDestroy(graph);
            } else {
                if (g && g->state == 0) g->state = 1;
                rc = enqueue_decoder();
            }
        }
        if (rc == B200RANK_OK) {
            float* ho = e->h_out_slot[b];
            cudaMemcpyAsync(ho, e->small_out, (size_t)n_docs * 2 * sizeof(float), cudaMemcpyDeviceToHost, e->stream_dec);
            cudaMemcpyAsync(ho + 2 * (size_t)n_docs, e->small_out2, (size_t)n_docs * sizeof(float), cudaMemcpyDeviceToHost, e->stream_dec);
            cudaError_t er = cudaEventRecord(e->ev_done[b], e->stream_dec);
            if (er != cudaSuccess) rc = set_error(B200RANK_ERR_CUDA, "event record: %s", cudaGetErrorString(er));
        }
    }
    if (rc != B200RANK_OK) return rc;   // SubmitGuard drains what was enqueued and restores the synchronous defaults
    submit_guard.ok = true;
    e->slot[b].busy = true;
    e->slot[b].ticket = e->next_ticket;
    *ticket = e->next_ticket++;
    return B200RANK_OK;
}

extern "C" int b200rank_wait_yes_no(b200rank_engine* e, uint64_t ticket, float* logits2, float* scores) {
    RET_IF(check_ready_async(e));
    const int b = static_cast<int>(ticket & 1);
    if (!e->slot[b].busy || e->slot[b].ticket != ticket) return set_error(B200RANK_ERR_STATE, "ticket %llu is not in flight", (unsigned long long)ticket);
    CU_OK(cudaEventSynchronize(e->ev_done[b]));
    const int nd = e->slot[b].docs;
    if (logits2) memcpy(logits2, e->h_out_slot[b], (size_t)nd * 2 * sizeof(float));
    if (scores) memcpy(scores, e->h_out_slot[b] + 2 * (size_t)nd, (size_t)nd * sizeof(float));
    e->slot[b].busy = false;
    return B200RANK_OK;
}

// ------------------------------------------------------------------ qlm
extern "C" int b200rank_score_qlm(b200rank_engine* e, const int32_t* ids, const int32_t* lengths, int n_docs, int stride,
                                  const int32_t* labels, int T, float* scores) {
    RET_IF(check_ready(e));
    if (!ids || !lengths || !labels || !scores || n_docs < 0 || stride <= 0 || T <= 0) return set_error(B200RANK_ERR_ARG, "bad arguments");
    if (T > e->cap_T) return set_error(B200RANK_ERR_CAPACITY, "label length %d exceeds max_dec_len %d", T, e->cap_T);
    for (int t = 0; t < T; ++t)
        if (labels[t] < 0 || labels[t] >= e->V) return set_error(B200RANK_ERR_ARG, "label id out of vocabulary");
    const int doc_limit = std::min(e->cap_docs, e->cap_logit_rows / T);
    if (doc_limit <= 0) return set_error(B200RANK_ERR_CAPACITY, "max_logit_rows %d < T %d", e->cap_logit_rows, T);
    // decoder inputs = shift_right(labels): [decoder_start(=pad), labels[0..T-2]]  (modeling_t5.py:595-614)
    std::vector<int> dec_row(T), lab_row(labels, labels + T);
    dec_row[0] = e->cfg.pad_id;
    for (int t = 1; t < T; ++t) dec_row[t] = labels[t - 1];
    int d0 = 0;
    while (d0 < n_docs) {
        int d1 = 0;
        RET_IF(next_group(e, lengths, n_docs, stride, d0, doc_limit, &d1));
        const int nd = d1 - d0, R = nd * T;
        RET_IF(stage_range(e, ids, lengths, stride, d0, d1));
        RET_IF(run_encoder(e));
        std::vector<int> dec((size_t)R), lab((size_t)R);
        for (int i = 0; i < nd; ++i) { std::copy(dec_row.begin(), dec_row.end(), dec.begin() + (size_t)i * T); std::copy(lab_row.begin(), lab_row.end(), lab.begin() + (size_t)i * T); }
        RET_IF(upload_ints(e, e->d_dec_ids, dec));
        RET_IF(upload_ints(e, e->d_labels, lab));
        RET_IF(run_decoder(e, 0, nd, T));
        RET_IF(gemm(e, e->hd, e->d, e->cap_rows, e->lm_head, e->d, e->V, R, e->V, e->d, EPI_F32, e->logits, e->V));
        RET_IF(k_vocab_row(e, "vocab_row_logprob", R, 0, e->d_labels, nullptr, 0, e->small_out, nullptr));
        prof_begin(e, "sum_rows"); launch_k(sum_rows_kernel, dim3((nd + 127) / 128), dim3(128), 0, e->stream, e->small_out, T, e->small_out2, nd);
        RET_IF(post_launch(e, "sum_rows"));
        CU_OK(cudaMemcpyAsync(e->h_out, e->small_out2, (size_t)nd * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
        CU_OK(cudaStreamSynchronize(e->stream));
        memcpy(scores + d0, e->h_out, (size_t)nd * sizeof(float));
        d0 = d1;
    }
    return B200RANK_OK;
}

// ------------------------------------------------------------------ logits_at
extern "C" int b200rank_logits_at(b200rank_engine* e, const int32_t* ids, const int32_t* lengths, int n_docs, int stride,
                                  const int32_t* dec_prefix, int prefix_len, const int32_t* cols, int ncols, int normalize,
                                  float* out) {
    RET_IF(check_ready(e));
    if (!ids || !lengths || !dec_prefix || !cols || !out || n_docs < 0 || stride <= 0 || prefix_len <= 0 || ncols <= 0 || ncols > 32)
        return set_error(B200RANK_ERR_ARG, "bad arguments (ncols must be 1..32)");
    const int T = prefix_len;
    if (T > e->cap_T) return set_error(B200RANK_ERR_CAPACITY, "prefix length %d exceeds max_dec_len %d", T, e->cap_T);
    for (int c = 0; c < ncols; ++c)
        if (cols[c] < 0 || cols[c] >= e->V) return set_error(B200RANK_ERR_ARG, "column id out of vocabulary");
    const int doc_limit = std::min(std::min(e->cap_docs, e->cap_rows / T), normalize ? e->cap_logit_rows : e->cap_docs);
    int d0 = 0;
    while (d0 < n_docs) {
        int d1 = 0;
        RET_IF(next_group(e, lengths, n_docs, stride, d0, doc_limit, &d1));
        const int nd = d1 - d0;
        RET_IF(stage_range(e, ids, lengths, stride, d0, d1));
        RET_IF(run_encoder(e));
        std::vector<int> dec((size_t)nd * T);
        for (int i = 0; i < nd; ++i) std::copy(dec_prefix, dec_prefix + T, dec.begin() + (size_t)i * T);
        RET_IF(upload_ints(e, e->d_dec_ids, dec));
        RET_IF(upload_ints(e, e->d_cols, std::vector<int>(cols, cols + ncols)));
        auto enqueue_pass = [&]() -> int {
            RET_IF(run_decoder(e, 0, nd, T));
            if (!normalize) {
                prof_begin(e, "lm_head_cols"); launch_k(lm_head_cols_kernel, dim3(nd), dim3(128), 0, e->stream, e->hd, e->d, T, T - 1, e->lm_head, e->d_cols, ncols, logit_scale(e), e->small_out);
                return post_launch(e, "lm_head_cols");
            }
            prof_begin(e, "gather_rows"); launch_k(gather_rows_kernel, dim3(nd), dim3(128), 0, e->stream, e->hd, e->d, T, T - 1, e->hlast, nd);
            RET_IF(post_launch(e, "gather_rows"));
            RET_IF(gemm(e, e->hlast, e->d, e->cap_rows, e->lm_head, e->d, e->V, nd, e->V, e->d, EPI_F32, e->logits, e->V));
            return k_vocab_row(e, "vocab_row_softmax_gather", nd, 2, nullptr, e->d_cols, ncols, e->small_out, nullptr);
        };
        RET_IF(run_as_graph(e, {2, nd, T, ncols, normalize ? 1 : 0, T == 1 ? ((e->staged_maxlen + 15) & ~15) : 0, 0, cross_split_off() ? 1 : 0}, enqueue_pass));
        CU_OK(cudaMemcpyAsync(e->h_out, e->small_out, (size_t)nd * ncols * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
        CU_OK(cudaStreamSynchronize(e->stream));
        memcpy(out + (size_t)d0 * ncols, e->h_out, (size_t)nd * ncols * sizeof(float));
        d0 = d1;
    }
    return B200RANK_OK;
}

// ------------------------------------------------------------------ greedy
// dec_ids[doc, T_cap] updated on device: write argmax (or pad once finished) at position t.
__global__ void greedy_update_kernel(const int* __restrict__ argmax, int* __restrict__ dec_rows, int* __restrict__ finished,
                                     int* __restrict__ new_ids, int nd, int t_write, int t_stride, int step, int max_new,
                                     int eos, int pad) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd) return;
    int tok = argmax[i];
    if (finished[i]) tok = pad;            // generation/utils.py:2797: next_tokens * unfinished + pad * (1 - unfinished)
    if (tok == eos) finished[i] = 1;
    new_ids[i * max_new + step] = tok;
    if (t_write < t_stride) dec_rows[i * t_stride + t_write] = tok;
}
// expand dec rows [nd, t_stride] -> contiguous [nd, T] ids for this step: positions t0 .. t0 + T - 1 of every document
__global__ void dec_rows_to_ids_kernel(const int* __restrict__ dec_rows, int t_stride, int t0, int T, int* __restrict__ dec_ids, int nd) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd * T) return;
    dec_ids[i] = dec_rows[(i / T) * t_stride + t0 + (i % T)];
}

extern "C" int b200rank_greedy(b200rank_engine* e, const int32_t* ids, const int32_t* lengths, int n_docs, int stride,
                               const int32_t* dec_prefix, int prefix_len, int max_new, int32_t* new_ids) {
    RET_IF(check_ready(e));
    if (!ids || !lengths || !dec_prefix || !new_ids || n_docs < 0 || stride <= 0 || prefix_len <= 0 || max_new <= 0 || max_new > kGreedyMaxNew)
        return set_error(B200RANK_ERR_ARG, "bad arguments (max_new must be 1..%d)", kGreedyMaxNew);
    const int Tmax = prefix_len + max_new - 1;  // longest decoder input actually run
    if (Tmax > e->cap_T) return set_error(B200RANK_ERR_CAPACITY, "prefix+new %d exceeds max_dec_len %d", Tmax, e->cap_T);
    const int t_stride = prefix_len + max_new;
    // Self-attention K/V cache (the cached loop of generation/utils.py:2762-2804): step 0 runs the prefix and fills the cache, every
    // later step runs ONE new position per document, so the decoder work of a call is linear in the tokens generated. The cached
    // kernel holds <= 64 positions and 64-wide heads; anything else (and B200RANK_KV_CACHE=0) re-runs the growing prefix per step.
    const bool cached = !kv_cache_off() && max_new > 1 && e->dkv == 64 && Tmax <= 64;
    if (cached && !e->kvc) {
        e->kvc_rows = getenv("B200RANK_KV_CACHE_ROWS") ? std::max(128, atoi(getenv("B200RANK_KV_CACHE_ROWS"))) : kKvCacheRows;
        RET_IF(dev_alloc(e, &e->kvc, (size_t)e->Ld * e->kvc_rows * 3 * e->inner));
    }
    int doc_limit = std::min(std::min(e->cap_docs, e->cap_rows / t_stride), e->cap_logit_rows);
    if (cached) doc_limit = std::min(doc_limit, e->kvc_rows / t_stride);
    if (doc_limit <= 0) return set_error(B200RANK_ERR_CAPACITY, "prefix+new %d leaves no room for a document (decoder rows %d)", t_stride, e->cap_rows);
    int* d_argmax = e->d_int_out + (size_t)e->cap_docs * kGreedyMaxNew;  // behind the new ids
    int d0 = 0;
    while (d0 < n_docs) {
        int d1 = 0;
        RET_IF(next_group(e, lengths, n_docs, stride, d0, doc_limit, &d1));
        const int nd = d1 - d0;
        RET_IF(stage_range(e, ids, lengths, stride, d0, d1));
        RET_IF(run_encoder(e));
        // decoder token rows live in d_labels ([nd, t_stride]); finished flags in d_finished; new ids in d_int_out
        std::vector<int> rows((size_t)nd * t_stride, e->cfg.pad_id);
        for (int i = 0; i < nd; ++i) std::copy(dec_prefix, dec_prefix + prefix_len, rows.begin() + (size_t)i * t_stride);
        RET_IF(upload_ints(e, e->d_labels, rows));
        CU_OK(cudaMemsetAsync(e->d_finished, 0, (size_t)nd * sizeof(int), e->stream));
        for (int step = 0; step < max_new; ++step) {
            const int T = prefix_len + step;   // decoder positions so far; the new token is written at position T
            auto enqueue_step = [&]() -> int {
                const bf16* h_last = nullptr;
                if (cached && step > 0) {
                    prof_begin(e, "dec_rows_to_ids"); launch_k(dec_rows_to_ids_kernel, dim3((nd + 255) / 256), dim3(256), 0, e->stream, e->d_labels, t_stride, T - 1, 1, e->d_dec_ids, nd);
                    RET_IF(post_launch(e, "dec_rows_to_ids"));
                    RET_IF(run_decoder_cached_step(e, 0, nd, T - 1, t_stride));
                    h_last = e->hd;
                } else {
                    // cache-less loop: the decoder prefix is re-run — token-for-token the same greedy choice
                    prof_begin(e, "dec_rows_to_ids"); launch_k(dec_rows_to_ids_kernel, dim3((nd * T + 255) / 256), dim3(256), 0, e->stream, e->d_labels, t_stride, 0, T, e->d_dec_ids, nd);
                    RET_IF(post_launch(e, "dec_rows_to_ids"));
                    RET_IF(run_decoder(e, 0, nd, T, cached ? t_stride : 0));
                    prof_begin(e, "gather_rows"); launch_k(gather_rows_kernel, dim3(nd), dim3(128), 0, e->stream, e->hd, e->d, T, T - 1, e->hlast, nd);
                    RET_IF(post_launch(e, "gather_rows"));
                    h_last = e->hlast;
                }
                RET_IF(gemm(e, h_last, e->d, e->cap_rows, e->lm_head, e->d, e->V, nd, e->V, e->d, EPI_F32, e->logits, e->V));
                RET_IF(k_vocab_row(e, "vocab_row_argmax", nd, 1, nullptr, nullptr, 0, nullptr, d_argmax));
                prof_begin(e, "greedy_update"); launch_k(greedy_update_kernel, dim3((nd + 127) / 128), dim3(128), 0, e->stream, d_argmax, e->d_labels, e->d_finished, e->d_int_out, nd, T,
                                                                              t_stride, step, max_new, e->cfg.eos_id, e->cfg.pad_id);
                return post_launch(e, "greedy_update");
            };
            // every launch argument of a step follows from (documents, prefix, budget, step) — token ids, lengths and finished flags are
            // read on the device — plus the longest document where one decoder position selects its kernels by it
            RET_IF(run_as_graph(e, {1, nd, prefix_len, max_new, step, T == 1 ? ((e->staged_maxlen + 15) & ~15) : 0, cached ? 1 : 0, cross_split_off() ? 1 : 0}, enqueue_step));
        }
        CU_OK(cudaMemcpyAsync(e->h_int, e->d_int_out, (size_t)nd * max_new * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        CU_OK(cudaStreamSynchronize(e->stream));
        memcpy(new_ids + (size_t)d0 * max_new, e->h_int, (size_t)nd * max_new * sizeof(int));
        d0 = d1;
    }
    return B200RANK_OK;
}

// ------------------------------------------------------------------ measurement plumbing
extern "C" int b200rank_event_record(b200rank_engine* e, int which) {
    if (!e || which < 0 || which > 1) return set_error(B200RANK_ERR_ARG, "bad arguments");
    CU_OK(cudaSetDevice(e->device));
    if (which == 1) {  // "stop" = after everything enqueued so far, including decoder passes of pipelined batches
        for (int b = 0; b < 2; ++b) CU_OK(cudaStreamWaitEvent(e->stream_main, e->ev_done[b], 0));
    }
    CU_OK(cudaEventRecord(e->ev[which], e->stream_main));
    return B200RANK_OK;
}
extern "C" int b200rank_event_elapsed_ms(b200rank_engine* e, float* ms) {
    if (!e || !ms) return set_error(B200RANK_ERR_ARG, "bad arguments");
    CU_OK(cudaSetDevice(e->device));
    CU_OK(cudaEventSynchronize(e->ev[1]));
    CU_OK(cudaEventElapsedTime(ms, e->ev[0], e->ev[1]));
    return B200RANK_OK;
}
extern "C" int b200rank_launch_count(b200rank_engine* e, uint64_t* n) {
    if (!e || !n) return set_error(B200RANK_ERR_ARG, "bad arguments");
    *n = e->launches;
    return B200RANK_OK;
}
extern "C" int b200rank_profile(b200rank_engine* e, int enable) {
    if (!e) return set_error(B200RANK_ERR_ARG, "null engine");
    CU_OK(cudaSetDevice(e->device));
    prof_collect(e);
    e->profiling = enable != 0;
    if (enable) e->prof_acc.clear();
    return B200RANK_OK;
}
extern "C" int b200rank_profile_report(b200rank_engine* e, char* buf, int buflen) {
    if (!e || !buf || buflen <= 2) return set_error(B200RANK_ERR_ARG, "bad arguments");
    CU_OK(cudaSetDevice(e->device));
    prof_collect(e);
    std::string js = "{";
    bool first = true;
    for (auto& kv : e->prof_acc) {
        char line[256];
        snprintf(line, sizeof line, "%s\"%s\": {\"ms\": %.6f, \"n\": %llu}", first ? "" : ", ", kv.first.c_str(), kv.second.first,
                 (unsigned long long)kv.second.second);
        js += line;
        first = false;
    }
    js += "}";
    if ((int)js.size() + 1 > buflen) return set_error(B200RANK_ERR_ARG, "profile report needs %zu bytes", js.size() + 1);
    memcpy(buf, js.c_str(), js.size() + 1);
    return B200RANK_OK;
}
__global__ void fill_kernel(uint4* p, size_t n, uint32_t v) {
    pdl_trigger();
    pdl_wait();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(v, v, v, v);
}
extern "C" int b200rank_flush_l2(b200rank_engine* e) {
    if (!e) return set_error(B200RANK_ERR_ARG, "null engine");
    CU_OK(cudaSetDevice(e->device));
    if (!e->l2_scratch) {
        e->l2_scratch_bytes = 256ull << 20;
        CU_OK(cudaMalloc(reinterpret_cast<void**>(&e->l2_scratch), e->l2_scratch_bytes));
    }
    fill_kernel<<<e->num_sms * 8, 256, 0, e->stream>>>(reinterpret_cast<uint4*>(e->l2_scratch), e->l2_scratch_bytes / 16, (uint32_t)e->launches);
    CU_OK(cudaGetLastError());
    return B200RANK_OK;
}
extern "C" int b200rank_device_info(b200rank_engine* e, int* sm_count, size_t* weight_bytes, size_t* workspace_bytes) {
    if (!e) return set_error(B200RANK_ERR_ARG, "null engine");
    if (sm_count) *sm_count = e->num_sms;
    if (weight_bytes) *weight_bytes = e->arena_bytes;
    if (workspace_bytes) *workspace_bytes = e->workspace_bytes;
    return B200RANK_OK;
}

// ------------------------------------------------------------------ kernel-level test hooks
extern "C" int b200rank_test_gemm(int device, const void* a_bf16, const void* w_bf16, int M, int N, int K, int epi, int block_n,
                                  int use_simt, void* out, float* elapsed_ms) {
    if (!a_bf16 || !w_bf16 || !out || M <= 0 || N <= 0 || K <= 0) return set_error(B200RANK_ERR_ARG, "bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return set_error(B200RANK_ERR_CUDA, "no CUDA device available; b200rank has no CPU fallback");
    CU_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_OK(cudaGetDeviceProperties(&prop, device));
    const int n_out = epi == EPI_GATED_BF16 ? N / 2 : N;
    const size_t out_elem = (epi == EPI_BF16 || epi == EPI_GATED_BF16) ? 2 : 4;
    bf16 *dA = nullptr, *dW = nullptr; void* dO = nullptr;
    const size_t Mp = align_up(M, 128), Np = align_up(N, 256);
    CU_OK(cudaMalloc((void**)&dA, Mp * K * 2)); CU_OK(cudaMalloc((void**)&dW, Np * K * 2)); CU_OK(cudaMalloc(&dO, (size_t)M * n_out * out_elem));
    CU_OK(cudaMemset(dA, 0, Mp * K * 2)); CU_OK(cudaMemset(dW, 0, Np * K * 2));
    CU_OK(cudaMemcpy(dA, a_bf16, (size_t)M * K * 2, cudaMemcpyHostToDevice));
    CU_OK(cudaMemcpy(dW, w_bf16, (size_t)N * K * 2, cudaMemcpyHostToDevice));
    CU_OK(cudaMemcpy(dO, out, (size_t)M * n_out * out_elem, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CU_OK(cudaEventCreate(&e0)); CU_OK(cudaEventCreate(&e1));
    int rc = B200RANK_OK;
    CU_OK(cudaEventRecord(e0, 0));
    if (use_simt) {
        dim3 blk(32, 8), grd((n_out + 31) / 32, (M + 7) / 8);
        gemm_simt_debug_kernel<<<grd, blk>>>(dA, K, dW, K, M, N, K, epi, 256, dO, n_out);
    } else {
        const int bn = block_n ? block_n : pick_block_n(M, N, K, epi, prop.multiProcessorCount);
        CUtensorMap ta, tb, tout;
        const bool direct = getenv("B200RANK_GEMM_DIRECT_EPI") && atoi(getenv("B200RANK_GEMM_DIRECT_EPI")) != 0;
        const int cg = direct ? 1 : pick_cta_group(M, N, bn, prop.multiProcessorCount);
        rc = make_tmap(&ta, dA, Mp, K, K, kGemmBlockM);
        if (rc == B200RANK_OK) rc = make_tmap(&tb, dW, Np, K, K, bn / cg);
        if (rc == B200RANK_OK) rc = make_tmap(&tout, dO, M, n_out, n_out, kGemmBlockM, out_elem == 4 ? 2 : 1);
        GemmArgs args{M, N, K, dO, n_out, 0, 0};
        if (rc == B200RANK_OK) rc = launch_gemm_tc(0, prop.multiProcessorCount, ta, tb, tout, args, epi, bn, !direct, cg);
    }
    if (rc == B200RANK_OK) {
        cudaEventRecord(e1, 0);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) rc = set_error(B200RANK_ERR_CUDA, "test_gemm kernel failed: %s", cudaGetErrorString(err));
        else {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (elapsed_ms) *elapsed_ms = ms;
            cudaMemcpy(out, dO, (size_t)M * n_out * out_elem, cudaMemcpyDeviceToHost);
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dA); cudaFree(dW); cudaFree(dO);
    return rc;
}

extern "C" int b200rank_test_enc_attention(int device, const void* qkv_bf16, const int32_t* cu_seqlens, int n_docs, int num_heads,
                                           const float* bias, void* out_bf16, int mode) {
    if (!qkv_bf16 || !cu_seqlens || !bias || !out_bf16 || n_docs <= 0 || num_heads <= 0) return set_error(B200RANK_ERR_ARG, "bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return set_error(B200RANK_ERR_CUDA, "no CUDA device available; b200rank has no CPU fallback");
    CU_OK(cudaSetDevice(device));
    const int inner = num_heads * 64, tokens = cu_seqlens[n_docs];
    int maxlen = 0;
    for (int i = 0; i < n_docs; ++i) maxlen = std::max(maxlen, cu_seqlens[i + 1] - cu_seqlens[i]);
    bf16 *dq = nullptr, *dout = nullptr; int* dcu = nullptr; float* dbias = nullptr;
    const size_t qrows = align_up(tokens, 64) + 256;  // slack rows: tiles may read past the last document (masked)
    CU_OK(cudaMalloc((void**)&dq, qrows * 3 * inner * 2)); CU_OK(cudaMalloc((void**)&dout, (size_t)tokens * inner * 2));
    CU_OK(cudaMemset(dq, 0, qrows * 3 * inner * 2));
    CU_OK(cudaMalloc((void**)&dcu, (n_docs + 1) * 4)); CU_OK(cudaMalloc((void**)&dbias, (size_t)num_heads * kAttnBiasLen * 4));
    CU_OK(cudaMemcpy(dq, qkv_bf16, (size_t)tokens * 3 * inner * 2, cudaMemcpyHostToDevice));
    CU_OK(cudaMemcpy(dcu, cu_seqlens, (n_docs + 1) * 4, cudaMemcpyHostToDevice));
    CU_OK(cudaMemcpy(dbias, bias, (size_t)num_heads * kAttnBiasLen * 4, cudaMemcpyHostToDevice));
    CU_OK(cudaMemset(dout, 0, (size_t)tokens * inner * 2));
    float* dwide = nullptr;
    uint8_t* dbad = nullptr;
    CU_OK(cudaMalloc((void**)&dbad, (size_t)tokens * num_heads));
    CU_OK(cudaMemset(dbad, 0, (size_t)tokens * num_heads));
    {
        std::vector<float> wide = widen_enc_bias(bias, num_heads);
        CU_OK(cudaMalloc((void**)&dwide, wide.size() * 4));
        CU_OK(cudaMemcpy(dwide, wide.data(), wide.size() * 4, cudaMemcpyHostToDevice));
    }
    int rc = launch_enc_attention(nullptr, dq, 3 * inner, qrows, inner, dcu, n_docs, maxlen, num_heads, dbias, dout, inner, 0, mode, 0, dwide, dbad);
    cudaError_t err = cudaDeviceSynchronize();
    if (err == cudaSuccess && rc == B200RANK_OK) {   // the fix-up walk must leave the map clean for the next launch
        std::vector<uint8_t> hb((size_t)tokens * num_heads);
        cudaMemcpy(hb.data(), dbad, hb.size(), cudaMemcpyDeviceToHost);
        for (uint8_t b : hb) if (b) { rc = set_error(B200RANK_ERR_CUDA, "attention fix-up map not cleared"); break; }
    }
    cudaFree(dwide); cudaFree(dbad);
    if (rc != B200RANK_OK) { cudaFree(dq); cudaFree(dout); cudaFree(dcu); cudaFree(dbias); return rc; }
    if (err != cudaSuccess) rc = set_error(B200RANK_ERR_CUDA, "enc_attention kernel failed: %s", cudaGetErrorString(err));
    else cudaMemcpy(out_bf16, dout, (size_t)tokens * inner * 2, cudaMemcpyDeviceToHost);
    cudaFree(dq); cudaFree(dout); cudaFree(dcu); cudaFree(dbias);
    return rc;
}
