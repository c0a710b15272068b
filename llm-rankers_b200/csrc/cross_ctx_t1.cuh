// Cross-attention for ONE decoder position (T = 1, pointwise yes_no) in re-associated form.
//
// The reference projects every encoder position through W_k and W_v of every decoder layer (13.6 % of all FLOPs,
// modeling_t5.py:298-299) and then attends with a single query. With one query per (document, head) the same numbers are
//     scores[h, j] = q_h . (W_k,h e_j)      = (W_k,h^T q_h) . e_j        = q'_h . e_j
//     out_h        = W_v,h (sum_j p[h,j] e_j)                             = W_v,h ctx_h
// so the K/V projections of the S encoder positions collapse into two per-document products against the encoder output E
// itself: scores = Q' E^T and ctx = P E (Q' is [H, d], P is [H, S]). The stacked cross-K|V GEMM and its 1.8 GB of K/V traffic
// per 100 documents disappear; q' = W_k,h^T q_h and out_h = W_v,h ctx_h are tiny block-diagonal GEMMs (gemm_tcgen05).
//
// This kernel does the two E products: one CTA per document, 8 warps, bf16 mma.sync m16n8k16 with fp32 accumulation.
// E (S x d) is streamed twice through shared memory in 128-column chunks (cp.async double buffering): pass 1 accumulates
// scores over the chunks, then an exact fp32 softmax over the S keys (no position bias in cross-attention,
// modeling_t5.py:313-315; the packed layout has no padded keys), pass 2 produces ctx chunk by chunk.
// Heads are processed 16 at a time (the M of the MMA); H < 16 pads with zero rows.
#pragma once
#include "attention_enc.cuh"
#include "kernels_misc.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int kCtxKC = 128;            // columns of E per staged chunk
constexpr int kCtxLdE = kCtxKC + 8;    // padded smem row (272 B): ldmatrix rows land in different banks
constexpr int kCtxThreads = 256;
constexpr int kCtxStages = 3;          // cp.async ring depth: the kernel is L2-latency bound, two chunks stay in flight

__host__ __device__ constexpr int cross_ctx_smem_bytes(int s_pad) {
    return kCtxStages * s_pad * kCtxLdE * 2      // sE[stages][s_pad][kCtxLdE] bf16
           + kCtxStages * 16 * kCtxLdE * 2       // sQ[stages][16][kCtxLdE]   bf16
           + 16 * s_pad * 4             // sS[16][s_pad]        fp32 scores
           + 16 * (s_pad + 8) * 2;      // sP[16][s_pad + 8]    bf16 probabilities
}

// qp : [n_docs, H * d] bf16, head h at columns h*d ..            (q' = W_k,h^T q_h)
// E  : packed encoder output [tokens, d] bf16 (final-normed), document `doc` = rows cu[doc] .. cu[doc+1]
// ctx: [n_docs, H * d] bf16, head h at columns h*d ..            (sum_j p[h,j] e_j)
__global__ void __launch_bounds__(kCtxThreads)
cross_ctx_t1_kernel(const __nv_bfloat16* __restrict__ qp, const __nv_bfloat16* __restrict__ E, const int* __restrict__ cu,
                    __nv_bfloat16* __restrict__ ctx, int H, int d, int s_pad) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(128) uint8_t ctx_smem[];
    __nv_bfloat16* sE = reinterpret_cast<__nv_bfloat16*>(ctx_smem);
    __nv_bfloat16* sQ = sE + kCtxStages * s_pad * kCtxLdE;
    float* sS = reinterpret_cast<float*>(sQ + kCtxStages * 16 * kCtxLdE);
    __nv_bfloat16* sP = reinterpret_cast<__nv_bfloat16*>(sS + 16 * s_pad);
    const int ldP = s_pad + 8;

    const int doc = blockIdx.x;
    const int tok0 = cu[doc];
    const int S = cu[doc + 1] - tok0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int nchunks = d / kCtxKC;
    const int n_tiles = s_pad >> 3;          // 8-key tiles
    const size_t ldq = static_cast<size_t>(H) * d;
    const __nv_bfloat16* Ebase = E + static_cast<size_t>(tok0) * d;

    for (int h0 = 0; h0 < H; h0 += 16) {
        const int hn = min(16, H - h0);
        // The two passes read the same column chunks of E: chunk id cc in [0, 2*nchunks) maps to columns (cc % nchunks)*128 and
        // ring slot cc % kCtxStages, so the cp.async pipeline keeps running through the softmax between the passes.
        const int total = 2 * nchunks;
        auto stage = [&](int cc) {
            if (cc >= total) { cp_async_commit(); return; }  // keep the group count uniform
            const int c = cc % nchunks, buf = cc % kCtxStages;
            __nv_bfloat16* dE = sE + buf * s_pad * kCtxLdE;
            for (int idx = tid; idx < s_pad * 16; idx += kCtxThreads) {
                const int r = idx >> 4, ch = idx & 15;
                const bool ok = r < S;
                cp_async16(dE + r * kCtxLdE + ch * 8, Ebase + static_cast<size_t>(ok ? r : 0) * d + c * kCtxKC + ch * 8, ok);
            }
            __nv_bfloat16* dQ = sQ + buf * 16 * kCtxLdE;
            {
                const int r = tid >> 4, ch = tid & 15;  // 16 rows x 16 chunks = 256 threads
                const bool ok = r < hn;
                cp_async16(dQ + r * kCtxLdE + ch * 8,
                           qp + static_cast<size_t>(doc) * ldq + static_cast<size_t>(h0 + (ok ? r : 0)) * d + c * kCtxKC + ch * 8, ok);
            }
            cp_async_commit();
        };

        // ---------------- pass 1: scores[16, S] = Q' E^T, accumulated over the column chunks
        float sc[4][4];  // up to 4 key tiles per warp (s_pad <= 256 -> 32 tiles / 8 warps)
#pragma unroll
        for (int i = 0; i < 4; ++i) { sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f; }
#pragma unroll
        for (int i = 0; i < kCtxStages - 1; ++i) stage(i);
        for (int c = 0; c < nchunks; ++c) {
            const int buf = c % kCtxStages;
            stage(c + kCtxStages - 1);          // slot (c-1) % stages: its readers passed the barrier that ended iteration c-1
            cp_async_wait<kCtxStages - 1>();    // chunk c has landed
            __syncthreads();
            const __nv_bfloat16* bE = sE + buf * s_pad * kCtxLdE;
            const __nv_bfloat16* bQ = sQ + buf * 16 * kCtxLdE;
#pragma unroll
            for (int ks = 0; ks < kCtxKC / 16; ks += 2) {
                uint32_t a0[4], a1[4];
                ldmatrix_x4(a0, bQ + (lane & 15) * kCtxLdE + ks * 16 + (lane >> 4) * 8);
                ldmatrix_x4(a1, bQ + (lane & 15) * kCtxLdE + (ks + 1) * 16 + (lane >> 4) * 8);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int nt = warp + 8 * i;
                    if (nt < n_tiles) {
                        // 4 matrices: keys nt*8.. x k-chunks (2ks, 2ks+1, 2ks+2, 2ks+3)
                        uint32_t bfr[4];
                        ldmatrix_x4(bfr, bE + (nt * 8 + (lane & 7)) * kCtxLdE + ks * 16 + (lane >> 3) * 8);
                        mma_bf16_16816(sc[i], a0, bfr[0], bfr[1]);
                        mma_bf16_16816(sc[i], a1, bfr[2], bfr[3]);
                    }
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int nt = warp + 8 * i;
            if (nt < n_tiles) {
                const int col = nt * 8 + 2 * t4;
                sS[g * s_pad + col] = sc[i][0];
                sS[g * s_pad + col + 1] = sc[i][1];
                sS[(g + 8) * s_pad + col] = sc[i][2];
                sS[(g + 8) * s_pad + col + 1] = sc[i][3];
            }
        }
        __syncthreads();
        // ---------------- softmax over the S keys (fp32), P -> bf16; rows 2*warp and 2*warp+1
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int row = 2 * warp + rr;
            float m = -INFINITY;
            for (int j = lane; j < S; j += 32) m = fmaxf(m, sS[row * s_pad + j]);
            m = warp_max(m);
            float l = 0.f;
            for (int j = lane; j < S; j += 32) {
                const float p = __expf(sS[row * s_pad + j] - m);
                sS[row * s_pad + j] = p;
                l += p;
            }
            l = warp_sum(l);
            const float inv = 1.f / l;
            for (int j = lane; j < s_pad; j += 32) sP[row * ldP + j] = __float2bfloat16(j < S ? sS[row * s_pad + j] * inv : 0.f);
        }
        __syncthreads();
        // ---------------- pass 2: ctx[16, d] = P E, one 128-column chunk at a time (16 n-tiles, 2 per warp)
        for (int c = 0; c < nchunks; ++c) {
            const int cc = nchunks + c, buf = cc % kCtxStages;
            stage(cc + kCtxStages - 1);
            cp_async_wait<kCtxStages - 1>();
            __syncthreads();
            const __nv_bfloat16* bE = sE + buf * s_pad * kCtxLdE;
            float o[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
            for (int ks = 0; ks < (s_pad >> 4); ++ks) {
                uint32_t a[4], bfr[4];
                ldmatrix_x4(a, sP + (lane & 15) * ldP + ks * 16 + (lane >> 4) * 8);
                // E^T fragments: keys ks*16 + (0..7 | 8..15) x columns (2*warp | 2*warp+1)*8
                const int krow = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                const int ncol = (2 * warp + (lane >> 4)) * 8;
                ldmatrix_x4_trans(bfr, bE + krow * kCtxLdE + ncol);
                mma_bf16_16816(o[0], a, bfr[0], bfr[1]);
                mma_bf16_16816(o[1], a, bfr[2], bfr[3]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int col = c * kCtxKC + (2 * warp + i) * 8 + 2 * t4;
                if (g < hn)
                    *reinterpret_cast<uint32_t*>(ctx + static_cast<size_t>(doc) * ldq + static_cast<size_t>(h0 + g) * d + col) = pack_bf16(o[i][0], o[i][1]);
                if (g + 8 < hn)
                    *reinterpret_cast<uint32_t*>(ctx + static_cast<size_t>(doc) * ldq + static_cast<size_t>(h0 + g + 8) * d + col) = pack_bf16(o[i][2], o[i][3]);
            }
            __syncthreads();
        }
    }
}

}  // namespace b200
