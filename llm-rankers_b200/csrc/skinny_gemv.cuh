// Decoder projections for a HANDFUL of rows (R = documents x decoder positions <= 8): the generation prefixes of a setwise /
// pairwise compare (1-2 prompts, 2-3 positions), duoT5 pairs, small synchronous rerank() calls.
// A 128-row tcgen05 tile is 98 % padding there and each GEMM launch pays TMEM allocation, barrier setup, tensor-map fetches and a
// TMA -> MMA -> TMEM -> epilogue pipeline fill for a 2 KB result: ~10 us per launch, 170 launches per decoder pass, plus a separate
// T5LayerNorm launch in front of half of them. This kernel is the memory-bound formulation: every warp owns output columns, streams
// their weight rows once (coalesced 16 B loads, K contiguous as in the GEMM's B operand), keeps the R activation rows in shared
// memory as bf16 (the same rounding point as the GEMM's A operand) and does the fp32 dot products on the CUDA cores; the
// T5LayerNorm in front (modeling_t5.py:55-68) is recomputed by every CTA from the fp32 residual rows (R x K x 4 B from L2) instead
// of being a launch of its own. Epilogues mirror gemm_tcgen05.cuh: bf16 store, relu, in-place fp32 residual add, gated-gelu over
// the tile-interleaved wi_0 | wi_1 packing.
#pragma once
#include "kernels_misc.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int kSkinnyMaxRows = 8;
constexpr int kSkinnyThreads = 256;   // 8 warps: one activation row each while normalising, one output column each afterwards

enum SkinnyEpi : int { SK_BF16 = 0, SK_RESID_F32 = 1, SK_GATED_BF16 = 2, SK_RELU_BF16 = 3 };

// x    : fp32 rows [R, K] to be layer-normed with ln_w (x != nullptr), or
// a    : bf16 rows [R, lda] used as they are (x == nullptr)
// W    : bf16 [N, ldw], row n = output column n (SK_GATED: physical rows (f/128)*256 + f%128 and +128 hold wi_0[f] and wi_1[f])
// out  : SK_BF16 / SK_RELU / SK_GATED: bf16 [R, ldo];  SK_RESID_F32: fp32 [R, ldo], out += acc
// n_out: number of output columns (SK_GATED: F, the weight has 2F rows)
template <int EPI>
__global__ void __launch_bounds__(kSkinnyThreads)
skinny_gemv_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ a, int lda, const float* __restrict__ ln_w, float eps,
                   const __nv_bfloat16* __restrict__ W, int ldw, int R, int n_out, int K, void* __restrict__ out, int ldo) {
    pdl_trigger();
    extern __shared__ __align__(16) uint8_t skinny_smem[];
    __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(skinny_smem);   // [R][K]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_wait();
    // ---- stage the activation rows as bf16 (with the T5LayerNorm applied when x is given): warp r handles row r
    if (warp < R) {
        __nv_bfloat16* dst = sA + static_cast<size_t>(warp) * K;
        if (x != nullptr) {
            const float* src = x + static_cast<size_t>(warp) * K;
            float ss = 0.f;
            for (int k = lane * 4; k < K; k += 128) {
                const float4 v = *reinterpret_cast<const float4*>(src + k);
                ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
            ss = warp_sum(ss);
            const float r = rsqrtf(ss / static_cast<float>(K) + eps);
            for (int k = lane * 4; k < K; k += 128) {
                const float4 v = *reinterpret_cast<const float4*>(src + k);
                const float4 g = __ldg(reinterpret_cast<const float4*>(ln_w + k));
                uint2 o;
                o.x = pack_bf16(v.x * r * g.x, v.y * r * g.y);
                o.y = pack_bf16(v.z * r * g.z, v.w * r * g.w);
                *reinterpret_cast<uint2*>(dst + k) = o;
            }
        } else {
            const __nv_bfloat16* src = a + static_cast<size_t>(warp) * lda;
            for (int k = lane * 8; k < K; k += 256) *reinterpret_cast<uint4*>(dst + k) = *reinterpret_cast<const uint4*>(src + k);
        }
    }
    __syncthreads();
    // ---- every warp: output columns col = blockIdx.x * 8 + warp, + gridDim.x * 8, ...
    for (int col = blockIdx.x * 8 + warp; col < n_out; col += gridDim.x * 8) {
        const __nv_bfloat16* w0 = W + static_cast<size_t>(EPI == SK_GATED_BF16 ? (col / 128) * 256 + (col % 128) : col) * ldw;
        const __nv_bfloat16* w1 = w0 + static_cast<size_t>(128) * ldw;   // SK_GATED only
        float acc0[kSkinnyMaxRows], acc1[kSkinnyMaxRows];
#pragma unroll
        for (int r = 0; r < kSkinnyMaxRows; ++r) { acc0[r] = 0.f; acc1[r] = 0.f; }
        for (int k = lane * 8; k < K; k += 256) {
            const uint4 wv = *reinterpret_cast<const uint4*>(w0 + k);
            const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(&wv);
            float wf[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(wp[i]); wf[2 * i] = f.x; wf[2 * i + 1] = f.y; }
            float gf[8];
            if (EPI == SK_GATED_BF16) {
                const uint4 gv = *reinterpret_cast<const uint4*>(w1 + k);
                const __nv_bfloat162* gp = reinterpret_cast<const __nv_bfloat162*>(&gv);
#pragma unroll
                for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(gp[i]); gf[2 * i] = f.x; gf[2 * i + 1] = f.y; }
            }
#pragma unroll
            for (int r = 0; r < kSkinnyMaxRows; ++r) {
                if (r < R) {
                    const uint4 av = *reinterpret_cast<const uint4*>(sA + static_cast<size_t>(r) * K + k);
                    const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&av);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 f = __bfloat1622float2(ap[i]);
                        acc0[r] = fmaf(f.x, wf[2 * i], acc0[r]);
                        acc0[r] = fmaf(f.y, wf[2 * i + 1], acc0[r]);
                        if (EPI == SK_GATED_BF16) {
                            acc1[r] = fmaf(f.x, gf[2 * i], acc1[r]);
                            acc1[r] = fmaf(f.y, gf[2 * i + 1], acc1[r]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kSkinnyMaxRows; ++r) {
            if (r < R) {
                const float s0 = warp_sum(acc0[r]);
                const float s1 = EPI == SK_GATED_BF16 ? warp_sum(acc1[r]) : 0.f;
                if (lane == 0) {
                    if (EPI == SK_RESID_F32) {
                        reinterpret_cast<float*>(out)[static_cast<size_t>(r) * ldo + col] += s0;
                    } else {
                        const float v = EPI == SK_GATED_BF16 ? gelu_new(s0) * s1 : (EPI == SK_RELU_BF16 ? fmaxf(s0, 0.f) : s0);
                        reinterpret_cast<__nv_bfloat16*>(out)[static_cast<size_t>(r) * ldo + col] = __float2bfloat16(v);
                    }
                }
            }
        }
    }
}

}  // namespace b200
