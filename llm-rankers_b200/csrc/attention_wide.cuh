// Attention for heads that are not 64 wide (d_kv = 128: the T5-3B shape of castorini/monot5-3b-msmarco and duot5-3b-msmarco,
// the only checkpoints of the reference's T5 rankers that are not d_kv = 64 — pointwise.py:136-186, pairwise.py:296-352).
// ONE kernel family serves the three attention sites of the model, selected by KIND:
//   ATT_ENC      encoder self-attention over packed documents: queries and keys are rows cu[doc] + i of the fused qkv buffer,
//                bidirectional relative-position bias table [H][257] indexed by clamp(j - i, +-128) + 128
//   ATT_DEC_SELF decoder self-attention: rows doc*T + t, causal, unidirectional bias table [H][bias_len] indexed by min(i - j, last)
//   ATT_CROSS    cross-attention: queries rows doc*T + t of q, keys rows cu[doc] + j of the stacked cross-K|V buffer, no bias
// (modeling_t5.py:236-251 bias, :308 no 1/sqrt(d) scaling, :313-334 mask + fp32 softmax + P.V.)
// Machinery: the mma.sync m16n8k16 tile scheme of enc_attention_kernel / dec_attention_mma_kernel (4 warps x 16 query rows, 64-key
// blocks double-buffered with cp.async, XOR-swizzled shared tiles read with ldmatrix, fp32 online softmax) with the head width as a
// template parameter: [64][DKV] tiles, DKV/16 k-steps in Q.K^T, DKV/8 output tiles in P.V. Shared memory is dynamic (80 KB at 128).
// This is the capability path for the wide-head checkpoints, not a tuned one: the d_kv = 64 models keep their specialised kernels.
#pragma once
#include "attention_enc.cuh"

namespace b200 {

enum AttnKind : int { ATT_ENC = 0, ATT_DEC_SELF = 1, ATT_CROSS = 2 };

template <int DKV>
struct AttnWideCfg {
    static_assert(DKV % 16 == 0 && DKV >= 64 && DKV <= 128, "head width: 64..128 in steps of 16");
    static constexpr int kTileElems = 64 * DKV;                               // one [64][DKV] bf16 tile
    static constexpr int kSmemBytes = 5 * kTileElems * 2 + kAttnBiasLen * 4 + 128;   // Q, 2 x K, 2 x V, bias window, alignment slack
};

// element offset of (row, 8-element chunk) inside a [64][DKV] bf16 tile; the low three chunk bits are XOR-swizzled by the row
template <int DKV>
__device__ __forceinline__ int sw_off_w(int row, int chunk) { return row * DKV + ((chunk ^ (row & 7)) << 3); }

// Loads rows row0 .. row0+63 (zero-filled from `len` on) of a [*, DKV] slice with leading dimension ld into a swizzled tile.
template <int DKV>
__device__ __forceinline__ void load_tile_w(__nv_bfloat16* smem_tile, const __nv_bfloat16* gbase, size_t ld, int row0, int len, int tid) {
    constexpr int kChunksPerRow = DKV / 8;
#pragma unroll
    for (int it = 0; it < (64 * kChunksPerRow) / 128; ++it) {
        const int idx = tid + it * 128;
        const int r = idx / kChunksPerRow, c = idx % kChunksPerRow;
        const bool ok = (row0 + r) < len;
        const __nv_bfloat16* src = gbase + static_cast<size_t>(ok ? (row0 + r) : 0) * ld + c * 8;
        cp_async16(smem_tile + sw_off_w<DKV>(r, c), src, ok);
    }
}

// grid (query tiles of 64, H, n_docs), 128 threads, dynamic shared memory AttnWideCfg<DKV>::kSmemBytes.
template <int DKV, int KIND>
__global__ void __launch_bounds__(128, 2)
attention_wide_kernel(const __nv_bfloat16* __restrict__ q, int ldq, int T, const __nv_bfloat16* __restrict__ kv, size_t ldkv, int k_off,
                      int v_off, const int* __restrict__ cu, const float* __restrict__ bias, int bias_len,
                      __nv_bfloat16* __restrict__ out, int ldo) {
    using Cfg = AttnWideCfg<DKV>;
    constexpr int KS = DKV / 16;    // k-steps of Q.K^T
    constexpr int DT = DKV / 8;     // 8-wide output tiles of P.V
    pdl_trigger();
    pdl_wait();
    extern __shared__ uint8_t smem_raw_w[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>((reinterpret_cast<uintptr_t>(smem_raw_w) + 127) & ~uintptr_t(127));
    __nv_bfloat16* sK = sQ + Cfg::kTileElems;           // [2]
    __nv_bfloat16* sV = sK + 2 * Cfg::kTileElems;       // [2]
    float* sBias = reinterpret_cast<float*>(sV + 2 * Cfg::kTileElems);

    const int qt = blockIdx.x, h = blockIdx.y, doc = blockIdx.z;
    const int q0 = qt * 64;
    // queries: packed rows of the document (encoder) or T rows per document (decoder); keys: packed rows except decoder self-attention
    const int q_row0 = (KIND == ATT_ENC) ? cu[doc] : doc * T;
    const int q_len = (KIND == ATT_ENC) ? cu[doc + 1] - q_row0 : T;
    if (q0 >= q_len) return;
    const int key_row0 = (KIND == ATT_DEC_SELF) ? doc * T : cu[doc];
    const int kv_len = (KIND == ATT_DEC_SELF) ? T : cu[doc + 1] - key_row0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;

    const __nv_bfloat16* gq = q + static_cast<size_t>(q_row0) * ldq + h * DKV;
    const __nv_bfloat16* gk = kv + static_cast<size_t>(key_row0) * ldkv + k_off + h * DKV;
    const __nv_bfloat16* gv = kv + static_cast<size_t>(key_row0) * ldkv + v_off + h * DKV;
    if (KIND == ATT_ENC) {
        for (int i = tid; i < kAttnBiasLen; i += 128) sBias[i] = bias[h * kAttnBiasLen + i];
    } else if (KIND == ATT_DEC_SELF) {
        for (int i = tid; i <= kAttnRelClamp; i += 128) sBias[i] = bias[h * bias_len + min(i, bias_len - 1)];
    }
    // causal: keys beyond the last query of this tile are never attended
    const int k_end = (KIND == ATT_DEC_SELF) ? min(kv_len, q0 + 64) : kv_len;
    const int nkb = (k_end + 63) / 64;
    load_tile_w<DKV>(sQ, gq, ldq, q0, q_len, tid);
    load_tile_w<DKV>(sK, gk, ldkv, 0, kv_len, tid);
    load_tile_w<DKV>(sV, gv, ldkv, 0, kv_len, tid);
    cp_async_commit();

    uint32_t qf[KS][4];
    float o[DT][4];
#pragma unroll
    for (int i = 0; i < DT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    const int qi0 = q0 + warp * 16 + g;  // query index of accumulator rows c0/c1; +8 for c2/c3

    for (int kb = 0; kb < nkb; ++kb) {
        const int buf = kb & 1;
        __nv_bfloat16* sKb = sK + buf * Cfg::kTileElems;
        __nv_bfloat16* sVb = sV + buf * Cfg::kTileElems;
        if (kb + 1 < nkb) {
            load_tile_w<DKV>(sK + (buf ^ 1) * Cfg::kTileElems, gk, ldkv, (kb + 1) * 64, kv_len, tid);
            load_tile_w<DKV>(sV + (buf ^ 1) * Cfg::kTileElems, gv, ldkv, (kb + 1) * 64, kv_len, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (kb == 0) {
            // Q fragments (A operand): rows warp*16 + (lane % 16), 8-element chunk 2*ks + lane/16
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) ldmatrix_x4(qf[ks], sQ + sw_off_w<DKV>(warp * 16 + (lane & 15), 2 * ks + (lane >> 4)));
        }
        // ---- S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                // lanes 0-7: keys np*16+0..7 @ d-chunk 2ks ; 8-15: same keys @ 2ks+1 ; 16-23: keys +8 @ 2ks ; 24-31: keys +8 @ 2ks+1
                uint32_t kf[4];
                const int krow = np * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int kch = 2 * ks + ((lane >> 3) & 1);
                ldmatrix_x4(kf, sKb + sw_off_w<DKV>(krow, kch));
                mma_bf16_16816(s[2 * np], qf[ks], kf[0], kf[1]);
                mma_bf16_16816(s[2 * np + 1], qf[ks], kf[2], kf[3]);
            }
        }
        // ---- bias / masks, online softmax. Key 0 is visible to every query (bidirectional and cross: kv_len >= 1; causal: 0 <= i),
        // so the running maximum is finite after the first block and a fully masked later block contributes exp(-inf) = 0.
        const int kbase = kb * 64;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = kbase + nt * 8 + 2 * t4 + (e & 1);
                const int qi = qi0 + ((e >> 1) << 3);
                float v = s[nt][e];
                bool ok = j < kv_len;
                if (KIND == ATT_ENC) {
                    v += sBias[max(-kAttnRelClamp, min(kAttnRelClamp, j - qi)) + kAttnRelClamp];
                } else if (KIND == ATT_DEC_SELF) {
                    ok = ok && j <= qi;
                    v += sBias[min(max(qi - j, 0), kAttnRelClamp)];
                }
                v = ok ? v : -INFINITY;
                s[nt][e] = v;
                mx[e >> 1] = fmaxf(mx[e >> 1], v);
            }
        }
        float scale[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            scale[r] = __expf(m_run[r] - m_new);
            m_run[r] = m_new;
            l_run[r] *= scale[r];
        }
        uint32_t pf[4][4];  // P as A-operand fragments, one per 16-key step
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = __expf(s[nt][0] - m_run[0]);
            const float p1 = __expf(s[nt][1] - m_run[0]);
            const float p2 = __expf(s[nt][2] - m_run[1]);
            const float p3 = __expf(s[nt][3] - m_run[1]);
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
        }
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) {
            o[dt][0] *= scale[0];
            o[dt][1] *= scale[0];
            o[dt][2] *= scale[1];
            o[dt][3] *= scale[1];
        }
        // ---- O += P V   (V^T fragments via ldmatrix.trans)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {          // 16 keys per step
#pragma unroll
            for (int dp = 0; dp < DT / 2; ++dp) {   // pairs of 8-wide d tiles
                // lanes 0-7: keys ks*16+0..7 @ d-chunk 2dp ; 8-15: keys +8 @ 2dp ; 16-23: keys 0..7 @ 2dp+1 ; 24-31: keys +8 @ 2dp+1
                uint32_t vf[4];
                const int vrow = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                const int vch = 2 * dp + (lane >> 4);
                ldmatrix_x4_trans(vf, sVb + sw_off_w<DKV>(vrow, vch));
                mma_bf16_16816(o[2 * dp], pf[ks], vf[0], vf[1]);
                mma_bf16_16816(o[2 * dp + 1], pf[ks], vf[2], vf[3]);
            }
        }
        __syncthreads();
    }

    // ---- normalise and store
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
    __nv_bfloat16* obase = out + static_cast<size_t>(q_row0) * ldo + h * DKV;
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
        const int col = dt * 8 + 2 * t4;
        if (qi0 < q_len)
            *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(qi0) * ldo + col) = pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
        if (qi0 + 8 < q_len)
            *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(qi0 + 8) * ldo + col) = pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
    }
}

}  // namespace b200
