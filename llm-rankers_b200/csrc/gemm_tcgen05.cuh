// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  acc[M,N] = A[M,K] . W[N,K]^T
//
//   * A (activations) and W (nn.Linear weight, [out,in]) are both K-major bf16 in HBM.
//   * warp 0 (one lane): TMA producer  — cp.async.bulk.tensor tiles (128B swizzle) into a
//     multi-stage shared-memory ring, completion on "full" mbarriers.
//   * warp 1 (one lane): MMA issuer    — tcgen05.mma 128 x BLOCK_N x 16, fp32 accumulators in
//     TMEM (double-buffered: 2 x BLOCK_N columns), tcgen05.commit releases ring slots.
//   * warps 2..5: epilogue             — tcgen05.ld (32 lanes x 32 columns per warp),
//     fused epilogue (bf16 store / fp32 residual add / gated-GELU product / fp32 store),
//     overlapped with the next tile's MMAs through the second TMEM buffer.
//
// This replaces the cuBLAS GEMMs under T5Attention.{q,k,v,o} and T5DenseGatedActDense
// (transformers/models/t5/modeling_t5.py:277,298-299,338 and :115-132).
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace b200 {

enum EpiMode : int {
    EPI_BF16 = 0,       // out_bf16[m, n]            = acc
    EPI_RESID_F32 = 1,  // out_f32[m, n]            += acc              (residual stream, in place)
    EPI_GATED_BF16 = 2, // out_bf16[m, n/2 ...]      = gelu_new(acc[:, :BN/2]) * acc[:, BN/2:]  per N-tile
    EPI_F32 = 3,        // out_f32[m, n]             = acc
};

struct GemmArgs {
    int M, N, K;   // N counts accumulator columns (= weight rows); K is the contraction length
    void* out;     // bf16* or float* depending on the epilogue
    int ldo;       // leading dimension of out, in elements
};

constexpr int kGemmBlockM = 128;
constexpr int kGemmBlockK = 64;  // 64 bf16 = 128 B = one swizzle atom
constexpr int kGemmThreads = 192;

template <int BLOCK_N>
struct GemmCfg {
    static constexpr int kABytes = kGemmBlockM * kGemmBlockK * 2;
    static constexpr int kBBytes = BLOCK_N * kGemmBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    // fill ~192 KB with the ring; at least 3, at most 8 stages
    static constexpr int kStagesRaw = (192 * 1024) / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : (kStagesRaw < 3 ? 3 : kStagesRaw);
    static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
    static constexpr int kBarrierBytes = (2 * kStages + 4) * 8 + 16;
    static constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + 1024;  // +1024: manual alignment slack
};

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmArgs args) {
    using Cfg = GemmCfg<BLOCK_N>;
    constexpr int kStages = Cfg::kStages;
    static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "bad BLOCK_N");
    static_assert(EPI != EPI_GATED_BF16 || BLOCK_N % 64 == 0, "gated epilogue needs BLOCK_N % 64 == 0");

    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled TMA/UMMA tiles need 1024 B alignment.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * Cfg::kABytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;   // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int tiles_m = (args.M + kGemmBlockM - 1) / kGemmBlockM;
    const int tiles_n = (args.N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = tiles_m * tiles_n;
    const int num_kb = (args.K + kGemmBlockK - 1) / kGemmBlockK;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 128);
        }
        fence_barrier_init();
    } else if (warp_idx == 1) {
        tmem_alloc(tmem_base_smem, Cfg::kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp_idx == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * kGemmBlockM;
                const int n0 = (tile % tiles_n) * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    tma_load_2d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * kGemmBlockK, m0,
                                kEvictNormal);
                    tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * kGemmBlockK, n0,
                                kEvictLast);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        // -------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(kGemmBlockM, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
                const int acc = iter & 1;
                const uint32_t acc_phase = (iter >> 1) & 1;
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t desc_a = make_sw128_kmajor_desc(smem_u32(smem_a + stage * Cfg::kABytes));
                    const uint64_t desc_b = make_sw128_kmajor_desc(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
                    for (int k = 0; k < kGemmBlockK / 16; ++k) {
                        // advance 16 elements (32 B) along K inside the swizzle atom: +2 in (addr >> 4)
                        umma_bf16(tmem_d, desc_a + 2 * k, desc_b + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kb == num_kb - 1) umma_commit(&tmem_full_bar[acc]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue
        const int quarter = warp_idx & 3;  // TMEM lane quarter this warp may access
        const int row_in_tile = quarter * 32 + lane;
        int iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int m0 = (tile / tiles_n) * kGemmBlockM;
            const int nb = tile % tiles_n;
            const int row = m0 + row_in_tile;
            const bool row_ok = row < args.M;
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;

            if constexpr (EPI == EPI_GATED_BF16) {
                constexpr int HALF = BLOCK_N / 2;
                __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(args.out) + static_cast<size_t>(row) * args.ldo;
                const int n_out_total = args.N / 2;
#pragma unroll 1
                for (int c = 0; c < HALF; c += 32) {
                    uint32_t g[32], l[32];
                    tmem_ld32(taddr + c, g);
                    tmem_ld32(taddr + HALF + c, l);
                    tmem_ld_wait();
                    if (c + 32 == HALF) {
                        tc_fence_before();
                        mbar_arrive(&tmem_empty_bar[acc]);
                    }
                    const int col0 = nb * HALF + c;
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            if (col0 + j < n_out_total) {
                                uint4 v;
                                uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    float a0 = gelu_new(__uint_as_float(g[j + 2 * e])) * __uint_as_float(l[j + 2 * e]);
                                    float a1 = gelu_new(__uint_as_float(g[j + 2 * e + 1])) * __uint_as_float(l[j + 2 * e + 1]);
                                    pv[e] = pack_bf16(a0, a1);
                                }
                                *reinterpret_cast<uint4*>(out + col0 + j) = v;
                            }
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < BLOCK_N; c += 32) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c, r);
                    tmem_ld_wait();
                    if (c + 32 == BLOCK_N) {
                        tc_fence_before();
                        mbar_arrive(&tmem_empty_bar[acc]);
                    }
                    const int col0 = nb * BLOCK_N + c;
                    if (row_ok) {
                        if constexpr (EPI == EPI_BF16) {
                            __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(args.out) + static_cast<size_t>(row) * args.ldo;
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                if (col0 + j < args.N) {
                                    uint4 v;
                                    v.x = pack_bf16(__uint_as_float(r[j + 0]), __uint_as_float(r[j + 1]));
                                    v.y = pack_bf16(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                                    v.z = pack_bf16(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5]));
                                    v.w = pack_bf16(__uint_as_float(r[j + 6]), __uint_as_float(r[j + 7]));
                                    *reinterpret_cast<uint4*>(out + col0 + j) = v;
                                }
                            }
                        } else if constexpr (EPI == EPI_RESID_F32) {
                            float* out = reinterpret_cast<float*>(args.out) + static_cast<size_t>(row) * args.ldo;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                if (col0 + j < args.N) {
                                    float4 x = *reinterpret_cast<float4*>(out + col0 + j);
                                    x.x += __uint_as_float(r[j + 0]);
                                    x.y += __uint_as_float(r[j + 1]);
                                    x.z += __uint_as_float(r[j + 2]);
                                    x.w += __uint_as_float(r[j + 3]);
                                    *reinterpret_cast<float4*>(out + col0 + j) = x;
                                }
                            }
                        } else {  // EPI_F32
                            float* out = reinterpret_cast<float*>(args.out) + static_cast<size_t>(row) * args.ldo;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                if (col0 + j < args.N) {
                                    float4 x;
                                    x.x = __uint_as_float(r[j + 0]);
                                    x.y = __uint_as_float(r[j + 1]);
                                    x.z = __uint_as_float(r[j + 2]);
                                    x.w = __uint_as_float(r[j + 3]);
                                    *reinterpret_cast<float4*>(out + col0 + j) = x;
                                }
                            }
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

}  // namespace b200
