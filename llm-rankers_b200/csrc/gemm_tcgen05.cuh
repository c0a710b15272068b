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
    EPI_RELU_BF16 = 5,  // host-side alias: launched as EPI_BF16 with GemmArgs::relu = 1 (out_bf16 = max(acc, 0))
};

struct GemmArgs {
    int M, N, K;   // N counts accumulator columns (= weight rows); K is the contraction length
    void* out;     // bf16* or float* depending on the epilogue
    int ldo;       // leading dimension of out, in elements
    // Block-diagonal GEMM (n_per_batch > 0): column block b = n / n_per_batch of the output contracts A[:, b*K : (b+1)*K]
    // with W rows of that block, i.e. out[:, b-th block] = A_b . W_b^T for per-head weight slices (decoder T=1 fast path).
    int n_per_batch;
    int relu;      // EPI_BF16 only: out = max(acc, 0) (T5 v1.0 DenseReluDense, modeling_t5.py:88-103)
};

constexpr int kGemmBlockM = 128;
constexpr int kGemmBlockK = 64;  // 64 bf16 = 128 B = one swizzle atom
constexpr int kGemmThreads = 192;

template <int BLOCK_N, int CG = 1>
struct GemmCfg {
    static constexpr int kABytes = kGemmBlockM * kGemmBlockK * 2;
    static constexpr int kBRows = BLOCK_N / CG;                      // a CTA pair splits the B tile
    static constexpr int kBBytes = kBRows * kGemmBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    // fill ~192 KB with the ring; at least 3, at most 8 stages (BLOCK_N=256: 4 x 48 KB, or 6 x 32 KB for a CTA pair)
    static constexpr int kStagesRaw = (192 * 1024) / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : (kStagesRaw < 3 ? 3 : kStagesRaw);
    static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
    static constexpr int kBarrierBytes = (2 * kStages + 4) * 8 + 16;
    static constexpr int kStagingBytes = 2 * kGemmBlockM * 128;  // two 128-row x 128-byte epilogue staging tiles (TMA store source)
    static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarrierBytes + 1024;  // +1024: manual alignment slack
};

// TMA_EPI = true : accumulators leave through shared-memory staging tiles (128 rows x 128 B, 128B swizzle) and
//                  cp.async.bulk.tensor stores — or cp.reduce.async.bulk.tensor .add for the fp32 residual, which
//                  performs x += acc inside L2 so the SMs never read the residual stream.
// TMA_EPI = false: per-thread 16 B global stores straight from registers (round-1 bring-up path, kept for bisecting
//                  with B200RANK_GEMM_DIRECT_EPI=1).
// Tried on the B200 in round 2 and dropped (same-box A/B, profiles/r02_gemm_epilogue_ab.txt): software-pipelined TMEM loads in the
// epilogues (no change) and TWO epilogue warpgroups splitting the tile's columns (slower: O-projection 1.10 -> 1.23 ms per step, one
// ring stage less) — the residual-epilogue GEMMs are not bound by epilogue threads but by L2: operand tiles at 64 B/clk/SM plus the
// read-modify-write of the fp32 residual in L2.
// CG = 2: the kernel runs as clusters of two CTAs (cta_group::2). One tcgen05.mma then covers 256 x BLOCK_N: each CTA keeps
//         its own 128 accumulator rows in its own TMEM and stages its own 128 rows of A but only HALF of the B tile — the
//         tensor core reads the other half from the peer's shared memory. That halves the L2->SM and shared-memory traffic
//         of the B operand (the 1-CTA 128x256 tile is bound by it: profiles/r01_ncu_summary_v4.txt). Only the leader CTA
//         issues MMAs; both CTAs run TMA producer and epilogue roles.
template <int BLOCK_N, int EPI, bool TMA_EPI, int CG = 1>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const GemmArgs args) {
    using Cfg = GemmCfg<BLOCK_N, CG>;
    static_assert(CG == 1 || CG == 2, "cta group");
    pdl_trigger();
    constexpr int kStages = Cfg::kStages;
    static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "bad BLOCK_N");
    static_assert(EPI != EPI_GATED_BF16 || BLOCK_N % 64 == 0, "gated epilogue needs BLOCK_N % 64 == 0");
    static_assert(!TMA_EPI || EPI != EPI_BF16 || BLOCK_N % 64 == 0 || BLOCK_N == 32, "bf16 staged epilogue moves 64-column tiles");
    static_assert(!TMA_EPI || EPI != EPI_GATED_BF16 || BLOCK_N % 128 == 0, "gated staged epilogue moves 64 output columns");

    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled TMA/UMMA tiles need 1024 B alignment.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * Cfg::kABytes;
    uint8_t* smem_stage = smem + kStages * Cfg::kStageBytes;  // 1024 B aligned (stage sizes are multiples of 1024)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_stage + Cfg::kStagingBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;   // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // a "unit" is one CTA (CG = 1) or one CTA pair (CG = 2); a tile is (128 * CG) x BLOCK_N
    const int cta_rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
    const int unit = blockIdx.x / CG, num_units = gridDim.x / CG;
    const int tiles_m = (args.M + kGemmBlockM * CG - 1) / (kGemmBlockM * CG);
    const int tiles_n = (args.N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = tiles_m * tiles_n;
    const int num_kb = (args.K + kGemmBlockK - 1) / kGemmBlockK;
    constexpr bool kResid = (EPI == EPI_RESID_F32);
    // it-th tile of this unit -> (m-block, n-block): tiles round-robin over units, n fastest
    auto get_tile = [&](int it, int& mb, int& nb) -> bool {
        const int tile = unit + it * num_units;
        mb = tile / tiles_n;
        nb = tile % tiles_n;
        return tile < num_tiles;
    };

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if constexpr (TMA_EPI) tma_prefetch_desc(&tmap_out);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], CG);      // CG = 2: leader's arrive.expect_tx + the peer's remote arrive (leader's copy is the live one)
            mbar_init(&empty_bar[s], 1);      // tcgen05.commit (multicast to both CTAs when CG = 2)
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 128 * CG);  // every epilogue thread of the unit (leader's copy is the live one)
        }
        fence_barrier_init();
    } else if (warp_idx == 1) {
        if constexpr (CG == 2) tmem_alloc_cg2(tmem_base_smem, Cfg::kTmemCols);
        else tmem_alloc(tmem_base_smem, Cfg::kTmemCols);
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all();  // peer barriers must be initialised before any remote arrive / multicast commit
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();  // barriers, TMEM and descriptors are set up; from here on we touch global memory written by the previous kernel

    if (warp_idx == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int mb, nb_;
            for (int it = 0; get_tile(it, mb, nb_); ++it) {
                const int m0 = mb * (kGemmBlockM * CG) + cta_rank * kGemmBlockM;
                const int n0 = nb_ * BLOCK_N + cta_rank * Cfg::kBRows;
                const int a_k0 = args.n_per_batch > 0 ? ((nb_ * BLOCK_N) / args.n_per_batch) * args.K : 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if constexpr (CG == 2) {
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
                        else mbar_arrive_cluster(&full_bar[stage], 0);
                        tma_load_2d_cg2(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], a_k0 + kb * kGemmBlockK, m0, kEvictNormal);
                        tma_load_2d_cg2(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * kGemmBlockK, n0, kEvictLast);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                        tma_load_2d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], a_k0 + kb * kGemmBlockK, m0, kEvictNormal);
                        tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * kGemmBlockK, n0, kEvictLast);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        // -------------------------------------------------------------- MMA issuer
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(kGemmBlockM * CG, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int mb, nb_;
            for (int iter = 0; get_tile(iter, mb, nb_); ++iter) {
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
                        if constexpr (CG == 2) umma_bf16_cg2(tmem_d, desc_a + 2 * k, desc_b + 2 * k, idesc, (kb | k) != 0);
                        else umma_bf16(tmem_d, desc_a + 2 * k, desc_b + 2 * k, idesc, (kb | k) != 0);
                    }
                    if constexpr (CG == 2) {
                        umma_commit_cg2(&empty_bar[stage], 0b11);
                        if (kb == num_kb - 1) umma_commit_cg2(&tmem_full_bar[acc], 0b11);
                    } else {
                        umma_commit(&empty_bar[stage]);
                        if (kb == num_kb - 1) umma_commit(&tmem_full_bar[acc]);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue
        const int quarter = warp_idx & 3;  // TMEM lane quarter this warp may access
        const int row_in_tile = quarter * 32 + lane;
        // hand the accumulator buffer back to the MMA issuer (who lives in the leader CTA)
        auto release_acc = [&](uint64_t* bar) {
            if constexpr (CG == 2) mbar_arrive_cluster(bar, 0);
            else mbar_arrive(bar);
        };
        int stage_sel = 0;  // which staging tile the next chunk uses (persists across tiles)
        int mb, nb;
        for (int iter = 0; get_tile(iter, mb, nb); ++iter) {
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int m0 = mb * (kGemmBlockM * CG) + cta_rank * kGemmBlockM;
            const int row = m0 + row_in_tile;
            const bool row_ok = row < args.M;
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;

            if constexpr (TMA_EPI) {
                // ---- staged epilogue: registers -> swizzled smem tile -> TMA store / reduce-add
                const bool issuer = (threadIdx.x == 64);  // warp 2, lane 0
                uint8_t* my_row = nullptr;
                auto stage_open = [&]() {
                    // the store that last read this buffer was committed two groups ago
                    if (issuer) tma_store_wait_read<1>();
                    named_bar_sync(1, 128);
                    my_row = smem_stage + stage_sel * (kGemmBlockM * 128) + row_in_tile * 128;
                };
                auto put16 = [&](int chunk, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
                    st_shared_v4(my_row + ((chunk ^ (row_in_tile & 7)) << 4), a, b, c, d);
                };
                auto stage_close = [&](int out_col0) {
                    fence_proxy_async();
                    named_bar_sync(1, 128);
                    if (issuer) {
                        const void* src = smem_stage + stage_sel * (kGemmBlockM * 128);
                        if constexpr (kResid) tma_reduce_add_2d(&tmap_out, src, out_col0, m0);
                        else tma_store_2d(&tmap_out, src, out_col0, m0);
                        tma_store_commit();
                    }
                    stage_sel ^= 1;
                };
                if constexpr (EPI == EPI_BF16) {
#pragma unroll 1
                    for (int c = 0; c < BLOCK_N; c += 64) {
                        uint32_t r0[32], r1[32];
                        tmem_ld32(taddr + c, r0);
                        tmem_ld32(taddr + c + 32, r1);
                        tmem_ld_wait();
                        if (c + 64 == BLOCK_N) { tc_fence_before(); release_acc(&tmem_empty_bar[acc]); }
                        if (args.relu) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                r0[j] = __float_as_uint(fmaxf(__uint_as_float(r0[j]), 0.f));
                                r1[j] = __float_as_uint(fmaxf(__uint_as_float(r1[j]), 0.f));
                            }
                        }
                        stage_open();
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            put16(j, pack_bf16(__uint_as_float(r0[8 * j + 0]), __uint_as_float(r0[8 * j + 1])),
                                  pack_bf16(__uint_as_float(r0[8 * j + 2]), __uint_as_float(r0[8 * j + 3])),
                                  pack_bf16(__uint_as_float(r0[8 * j + 4]), __uint_as_float(r0[8 * j + 5])),
                                  pack_bf16(__uint_as_float(r0[8 * j + 6]), __uint_as_float(r0[8 * j + 7])));
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            put16(4 + j, pack_bf16(__uint_as_float(r1[8 * j + 0]), __uint_as_float(r1[8 * j + 1])),
                                  pack_bf16(__uint_as_float(r1[8 * j + 2]), __uint_as_float(r1[8 * j + 3])),
                                  pack_bf16(__uint_as_float(r1[8 * j + 4]), __uint_as_float(r1[8 * j + 5])),
                                  pack_bf16(__uint_as_float(r1[8 * j + 6]), __uint_as_float(r1[8 * j + 7])));
                        stage_close(nb * BLOCK_N + c);
                    }
                } else if constexpr (EPI == EPI_GATED_BF16) {
                    constexpr int HALF = BLOCK_N / 2;
#pragma unroll 1
                    for (int c = 0; c < HALF; c += 64) {
                        stage_open();
#pragma unroll 1
                        for (int hh = 0; hh < 2; ++hh) {
                            uint32_t g[32], l[32];
                            tmem_ld32(taddr + c + 32 * hh, g);
                            tmem_ld32(taddr + HALF + c + 32 * hh, l);
                            tmem_ld_wait();
                            if (c + 64 == HALF && hh == 1) { tc_fence_before(); release_acc(&tmem_empty_bar[acc]); }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                uint32_t v[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float a0 = gelu_new(__uint_as_float(g[8 * j + 2 * e])) * __uint_as_float(l[8 * j + 2 * e]);
                                    const float a1 = gelu_new(__uint_as_float(g[8 * j + 2 * e + 1])) * __uint_as_float(l[8 * j + 2 * e + 1]);
                                    v[e] = pack_bf16(a0, a1);
                                }
                                put16(4 * hh + j, v[0], v[1], v[2], v[3]);
                            }
                        }
                        stage_close(nb * HALF + c);
                    }
                } else {  // fp32 tiles of 32 columns: plain store (EPI_F32) or in-L2 add (EPI_RESID_F32)
#pragma unroll 1
                    for (int c = 0; c < BLOCK_N; c += 32) {
                        uint32_t r[32];
                        tmem_ld32(taddr + c, r);
                        tmem_ld_wait();
                        if (c + 32 == BLOCK_N) { tc_fence_before(); release_acc(&tmem_empty_bar[acc]); }
                        stage_open();
#pragma unroll
                        for (int j = 0; j < 8; ++j) put16(j, r[4 * j + 0], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                        stage_close(nb * BLOCK_N + c);
                    }
                }
            } else if constexpr (EPI == EPI_GATED_BF16) {
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
                        release_acc(&tmem_empty_bar[acc]);
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
                        release_acc(&tmem_empty_bar[acc]);
                    }
                    const int col0 = nb * BLOCK_N + c;
                    if (row_ok) {
                        if constexpr (EPI == EPI_BF16) {
                            __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(args.out) + static_cast<size_t>(row) * args.ldo;
                            if (args.relu) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaxf(__uint_as_float(r[j]), 0.f));
                            }
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
                        } else if constexpr (kResid) {
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

    if constexpr (TMA_EPI) {
        // shared memory must stay alive until the bulk stores have READ it; their global writes complete with the grid
        if (threadIdx.x == 64) tma_store_wait_read<0>();
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all();  // no CTA of the pair may exit while its peer can still touch its smem / TMEM
    else __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::kTmemCols);
        else tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

}  // namespace b200
