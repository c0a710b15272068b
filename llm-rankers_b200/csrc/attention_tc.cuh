// tcgen05 encoder self-attention for short documents (len <= 256: the pointwise / pairwise regime).
//
// One CTA per (document, head). Q, K, V head slices (len x 64 bf16, 128 B rows) are TMA-loaded once into shared memory
// (128B swizzle). Then, per 128-query tile t (at most two):
//   MMA-1 (tcgen05, one thread):  S_t[128 x 64*NKB] = Q_t . K^T      fp32 in TMEM (columns t*256 ..)
//   softmax (warpgroup t, one thread per query row): TMEM -> registers, + relative-position bias, key-length mask,
//            exact two-pass softmax in fp32 (the whole row is resident, no online rescaling), P -> bf16 -> shared memory
//            in the K-major 128B-swizzled layout of an MMA A operand
//   MMA-2 (tcgen05):              O_t[128 x 64] = P_t . V            V consumed in place as an MN-major B operand
//   epilogue (warpgroup t):       O_t / rowsum -> bf16 -> global
// The two query tiles overlap: while warpgroup 0 does the softmax of tile 0 the tensor core computes S_1, etc.
// No 1/sqrt(d) scaling (modeling_t5.py:308); bias + mask + fp32 softmax as modeling_t5.py:313-334; padded keys do not
// exist in the packed layout (rows past `len` belong to the next document: they are masked to -inf / never stored).
#pragma once
#include <cuda.h>
#include "attention_enc.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int kAttnTcThreads = 320;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: softmax tile 0, warps 6-9: softmax tile 1

template <int NKB>
struct AttnTcCfg {
    static constexpr int kRows = 64 * NKB;              // padded keys / queries held in smem
    static constexpr int kQKVBytes = kRows * 128;       // one of Q, K, V
    static constexpr int kPBytes = NKB * 128 * 128;     // P tile: NKB k-blocks of [128 rows x 128 B]
    static constexpr int kSmemBytes = 3 * kQKVBytes + 2 * kPBytes + 8 * 8 + 16 + kAttnBiasLen * 4 + 1024;
};

// Instruction descriptor with an MN-major B operand (bit 16), otherwise as make_idesc_bf16.
__host__ __device__ constexpr uint32_t make_idesc_bf16_bmn(uint32_t m, uint32_t n) {
    return make_idesc_bf16(m, n) | (1u << 16);
}

template <int NKB>
__global__ void __launch_bounds__(kAttnTcThreads, 1)
enc_attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, int inner, const int* __restrict__ cu,
                        const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int ldo) {
    using Cfg = AttnTcCfg<NKB>;
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + Cfg::kQKVBytes;
    uint8_t* sV = sK + Cfg::kQKVBytes;
    uint8_t* sP = sV + Cfg::kQKVBytes;                      // [2][kPBytes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * Cfg::kPBytes);
    uint64_t* bar_qk = bars + 0;
    uint64_t* bar_v = bars + 1;
    uint64_t* bar_s = bars + 2;   // [2] S_t ready in TMEM
    uint64_t* bar_p = bars + 4;   // [2] P_t written to smem (128 arrivals)
    uint64_t* bar_o = bars + 6;   // [2] O_t ready in TMEM
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 8);
    float* sBias = reinterpret_cast<float*>(tmem_base_smem + 4);

    const int h = blockIdx.x, doc = blockIdx.y;
    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp_idx == 1) tmem_alloc(tmem_base_smem, 512);  // prologue work that needs no global data comes before pdl_wait
    pdl_wait();
    const int tok0 = cu[doc];
    const int len = cu[doc + 1] - tok0;
    const int ntiles = (len + 127) >> 7;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_qkv);
        mbar_init(bar_qk, 1);
        mbar_init(bar_v, 1);
        for (int t = 0; t < 2; ++t) {
            mbar_init(&bar_s[t], 1);
            mbar_init(&bar_p[t], 128);
            mbar_init(&bar_o[t], 1);
        }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < kAttnBiasLen; i += kAttnTcThreads) sBias[i] = bias[h * kAttnBiasLen + i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    const int nkb_used = (len + 63) >> 6;  // 64-key blocks that hold real keys

    if (warp_idx == 0) {
        if (lane == 0) {
            // Q and K first (MMA-1 needs both), then V. Rows past the document come from the next document (masked later).
            mbar_arrive_expect_tx(bar_qk, 2 * nkb_used * 8192);
            for (int b = 0; b < nkb_used; ++b) {
                tma_load_2d(sQ + b * 8192, &tmap_qkv, bar_qk, h * 64, tok0 + b * 64, kEvictFirst);
                tma_load_2d(sK + b * 8192, &tmap_qkv, bar_qk, inner + h * 64, tok0 + b * 64, kEvictFirst);
            }
            mbar_arrive_expect_tx(bar_v, nkb_used * 8192);
            for (int b = 0; b < nkb_used; ++b)
                tma_load_2d(sV + b * 8192, &tmap_qkv, bar_v, 2 * inner + h * 64, tok0 + b * 64, kEvictFirst);
        }
    } else if (warp_idx == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(128, 64 * NKB);
            constexpr uint32_t idesc_o = make_idesc_bf16_bmn(128, 64);
            mbar_wait(bar_qk, 0);
            tc_fence_after();
            for (int t = 0; t < ntiles; ++t) {
                const uint64_t da = make_sw128_kmajor_desc(smem_u32(sQ + t * 128 * 128));
                const uint64_t db = make_sw128_kmajor_desc(smem_u32(sK));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + t * 256, da + 2 * k, db + 2 * k, idesc_s, k != 0);
                umma_commit(&bar_s[t]);
            }
            mbar_wait(bar_v, 0);
            for (int t = 0; t < ntiles; ++t) {
                mbar_wait(&bar_p[t], 0);
                tc_fence_after();
                for (int kb = 0; kb < nkb_used; ++kb) {
                    const uint64_t da = make_sw128_kmajor_desc(smem_u32(sP + t * Cfg::kPBytes + kb * 16384));
                    // V block: rows = keys (the MMA K dimension), 128 B of head dims contiguous = MN-major B operand;
                    // 16 keys per MMA = 2048 B -> +128 in (addr >> 4)
                    const uint64_t db = make_sw128_kmajor_desc(smem_u32(sV + kb * 8192));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + t * 256, da + 2 * k, db + 128 * k, idesc_o, (kb | k) != 0);
                }
                umma_commit(&bar_o[t]);
            }
        }
    } else {
        const int t = (warp_idx - 2) >> 2;            // query tile of this warpgroup
        const int quarter = warp_idx & 3;             // TMEM lane quarter of this warp
        if (t < ntiles) {
            const int row_in_tile = quarter * 32 + lane;
            const int qi = t * 128 + row_in_tile;     // query index inside the document
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * 256;
            const int ncols = nkb_used * 64;
            mbar_wait(&bar_s[t], 0);
            tc_fence_after();
            // ---- pass 1: row maximum of scores + bias over the real keys
            float m = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < ncols; c += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c, r);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int j = c + e;
                    int rel = j - qi;
                    rel = max(-kAttnRelClamp, min(kAttnRelClamp, rel));
                    const float v = __uint_as_float(r[e]) + sBias[rel + kAttnRelClamp];
                    m = fmaxf(m, (j < len) ? v : -INFINITY);
                }
            }
            // ---- pass 2: p = exp(v - m), row sum, bf16 P into the swizzled A-operand tile
            const float m_l2 = m * 1.4426950408889634f;
            float l = 0.f;
            uint8_t* prow = sP + t * Cfg::kPBytes + row_in_tile * 128;
#pragma unroll 1
            for (int c = 0; c < ncols; c += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c, r);
                tmem_ld_wait();
                uint32_t packed[16];
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    float p[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int j = c + e + u;
                        int rel = j - qi;
                        rel = max(-kAttnRelClamp, min(kAttnRelClamp, rel));
                        const float v = __uint_as_float(r[e + u]) + sBias[rel + kAttnRelClamp];
                        p[u] = (j < len) ? exp2f(v * 1.4426950408889634f - m_l2) : 0.f;
                        l += p[u];
                    }
                    packed[e >> 1] = pack_bf16(p[0], p[1]);
                }
                // 32 keys = 64 B = 4 chunks of 16 B inside k-block c/64, chunk index ((c % 64) / 8 + i)
                uint8_t* kblk = prow + (c >> 6) * 16384;
                const int chunk0 = (c & 63) >> 3;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    st_shared_v4(kblk + (((chunk0 + i) ^ (row_in_tile & 7)) << 4), packed[4 * i], packed[4 * i + 1], packed[4 * i + 2],
                                 packed[4 * i + 3]);
            }
            tc_fence_before();      // this thread's TMEM reads of S_t are complete before MMA-2 overwrites the columns with O_t
            fence_proxy_async();    // generic-proxy smem writes -> visible to the tensor core (async proxy)
            mbar_arrive(&bar_p[t]);
            // ---- epilogue: O_t / l -> bf16 -> global
            mbar_wait(&bar_o[t], 0);
            tc_fence_after();
            uint32_t o0[32], o1[32];
            tmem_ld32(taddr, o0);
            tmem_ld32(taddr + 32, o1);
            tmem_ld_wait();
            if (qi < len) {
                const float inv = 1.f / l;
                uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(tok0 + qi) * ldo + h * 64);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 v;
                    v.x = pack_bf16(__uint_as_float(o0[8 * i + 0]) * inv, __uint_as_float(o0[8 * i + 1]) * inv);
                    v.y = pack_bf16(__uint_as_float(o0[8 * i + 2]) * inv, __uint_as_float(o0[8 * i + 3]) * inv);
                    v.z = pack_bf16(__uint_as_float(o0[8 * i + 4]) * inv, __uint_as_float(o0[8 * i + 5]) * inv);
                    v.w = pack_bf16(__uint_as_float(o0[8 * i + 6]) * inv, __uint_as_float(o0[8 * i + 7]) * inv);
                    dst[i] = v;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 v;
                    v.x = pack_bf16(__uint_as_float(o1[8 * i + 0]) * inv, __uint_as_float(o1[8 * i + 1]) * inv);
                    v.y = pack_bf16(__uint_as_float(o1[8 * i + 2]) * inv, __uint_as_float(o1[8 * i + 3]) * inv);
                    v.z = pack_bf16(__uint_as_float(o1[8 * i + 4]) * inv, __uint_as_float(o1[8 * i + 5]) * inv);
                    v.w = pack_bf16(__uint_as_float(o1[8 * i + 6]) * inv, __uint_as_float(o1[8 * i + 7]) * inv);
                    dst[4 + i] = v;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200
