// Persistent tcgen05 encoder self-attention, round-2 design ("tc5"): 16 softmax warps, two threads per query row, ONE pass over
// TMEM with no shift, for documents of <= 192 tokens (modeling_t5.py:308-334: scores + position bias + mask -> fp32 softmax -> P.V;
// no 1/sqrt(d) scaling in T5).
//
// What the round-1 kernel (enc_attention_tc2_kernel, attention_tc.cuh) was bound by, measured on the B200
// (profiles/r02_attn_profile.txt): 8 softmax warps, one thread per 192-column row, two passes through TMEM. ncu: issue slots 31 %
// busy, tensor pipe 12 %, the six warps that own real rows at S = 184 execute ~1550 instructions per item at an IPC of 0.12 —
// a per-thread latency chain (TMEM load ~53 clk dependent, LDS 29 clk, MUFU, the FADD/FMNMX chains), not a throughput limit:
// experiments/tmem_probe.cu shows the TMEM read rate still growing at 16 warps (960 B/clk/SM vs 300-600 with 4-8 warps).
// So this kernel buys thread-level parallelism and sheds instructions:
//   * 18 warps: warp 0 TMA, warp 1 MMA issue + TMEM allocation, warps 2..17 softmax. A softmax warp is (quarter q = warp_idx & 3:
//     the 32 TMEM lanes it may touch, tile slot t, column half g): a query row's 192 score columns are shared by two threads
//     (96 columns = three 32-column tcgen05.ld chunks each), twelve of the sixteen warps own real rows at S = 184.
//   * ONE pass, NO shift: softmax is shift-invariant and fp32 / bf16 share one exponent range, so p = 2^v (v = log2(e) s + bias')
//     is as precise as 2^(v - max) as long as the row maximum lies within +-100 (powers of two — natural-log scores within +-69).
//     No row maximum, no write-back of v, no second TMEM read, no cross-thread dependency ahead of the P tile: per column
//     LDS (bias), FFMA, MUFU.EX2, FADD (row sum), half an F2FP. The two threads of a row exchange their partial sums through
//     shared memory for the deferred epilogue. A row whose total sum leaves [2^-100, 2^100) or is not finite — scores beyond
//     +-69, which trained T5 checkpoints do not produce but the contract must survive — is recomputed exactly by its two threads
//     on CUDA cores from Q/K/V in global memory (attn_slow_row: two-pass softmax with the true maximum), so the result is always
//     the exact softmax; tests/test_engine_gpu.py drives that path with scores in the hundreds.
//   * the second query tile of a 129..192-token document has at most 64 real rows: on odd items the MMA's A descriptor starts 64
//     rows earlier, so those rows land on TMEM lanes 64..127 (warps of quarters 2, 3) instead of lanes 0..63 — over two items
//     every SM sub-partition gets the same softmax work instead of quarters 0, 1 doing twice that of 2, 3.
// Roles, barriers, TMEM columns (S_t at t*192, O_t at 384 + 64 t) and the item pipeline (TMA one item ahead; per tile slot MMA-2 of
// item k then MMA-1 of item k+1; epilogue of item k deferred to the start of item k+1) are those of enc_attention_tc2_kernel.
// A document's result is a function of the document alone (lane placement does not enter the arithmetic): bit-identical across
// batch compositions, like every other kernel of the path.
#pragma once
#include <cuda.h>
#include "attention_tc.cuh"

namespace b200 {

constexpr int kAttn5Threads = 576;

template <int NKB>
struct AttnTc5Cfg {
    static constexpr int kRows = 64 * NKB;
    static constexpr int kQKVBytes = kRows * 128;
    static constexpr int kPBytes = NKB * 128 * 128;
    static constexpr int kWideBias = 512;                 // 511 used: index (j - i) + 255
    static constexpr int kRedBytes = 2 * 2 * 2 * 128 * 4; // partial row sums [use parity][tile slot][column half][lane]
    static constexpr int kFixedBytes = 3 * kQKVBytes + 2 * kPBytes + 16 * 8 + 16 + kRedBytes + 1024;
    static constexpr int kMaxResidentHeads = (227 * 1024 - kFixedBytes) / (kWideBias * 4);
    static constexpr int smem_bytes(int H) { return kFixedBytes + (H <= kMaxResidentHeads ? (H < 2 ? 2 : H) : 2) * kWideBias * 4; }
    static constexpr int kSCol = 64 * NKB;
    static constexpr int kOCol0 = 2 * 64 * NKB;
    static constexpr int kHalf = 32 * NKB;                // score columns per thread
    static_assert(kOCol0 + 128 <= 512, "S and O tiles must fit the 512 TMEM columns side by side");
};

// Exact softmax(q_i K^T + bias) V for ONE query row and 32 of its 64 output dims on CUDA cores: the rare-row fallback of the
// one-pass kernel (row sum outside [2^-100, 2^100) or not finite). sBq[j] = log2(e) * bias(j - qi).
__device__ __noinline__ void attn_slow_row(const __nv_bfloat16* __restrict__ qkv, int ld, int inner, int tok0, int len, int qi, int h, int g,
                                           const float* sBq, __nv_bfloat16* __restrict__ out, int ldo) {
    const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(qkv + static_cast<size_t>(tok0 + qi) * ld + h * 64);
    auto score = [&](int j) {
        const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(qkv + static_cast<size_t>(tok0 + j) * ld + inner + h * 64);
        float s = 0.f;
        for (int d = 0; d < 32; ++d) {
            const float2 a = __bfloat1622float2(q2[d]), b = __bfloat1622float2(k2[d]);
            s = fmaf(a.x, b.x, s);
            s = fmaf(a.y, b.y, s);
        }
        return fmaf(s, 1.4426950408889634f, sBq[j]);
    };
    float m = -INFINITY;
    for (int j = 0; j < len; ++j) m = fmaxf(m, score(j));
    float l = 0.f, o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    for (int j = 0; j < len; ++j) {
        const float p = ex2_approx(score(j) - m);
        l += p;
        const float pb = __bfloat162float(__float2bfloat16_rn(p));   // the tensor-core path multiplies the bf16-rounded P
        const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(qkv + static_cast<size_t>(tok0 + j) * ld + 2 * inner + h * 64 + g * 32);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float2 v = __bfloat1622float2(v2[i]);
            o[2 * i] = fmaf(pb, v.x, o[2 * i]);
            o[2 * i + 1] = fmaf(pb, v.y, o[2 * i + 1]);
        }
    }
    const float inv = 1.f / l;
    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(tok0 + qi) * ldo + h * 64 + g * 32);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 v;
        v.x = pack_bf16(o[8 * i + 0] * inv, o[8 * i + 1] * inv);
        v.y = pack_bf16(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
        v.z = pack_bf16(o[8 * i + 4] * inv, o[8 * i + 5] * inv);
        v.w = pack_bf16(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
        dst[i] = v;
    }
}

template <int NKB>
__global__ void __maxnreg__(112)
enc_attention_tc5_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __nv_bfloat16* __restrict__ qkv, int ld, int inner,
                         const int* __restrict__ cu, const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int ldo, int H,
                         int n_items, int len_limit) {
    using Cfg = AttnTc5Cfg<NKB>;
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
    uint64_t* qk_free = bars + 2;  // both MMA-1s of the item retired: Q/K smem reusable
    uint64_t* v_free = bars + 3;   // MMA-2s of the item retired: V smem reusable
    uint64_t* bar_s = bars + 4;    // [2] S_t ready in TMEM
    uint64_t* bar_p = bars + 6;    // [2] P_t written to smem, S_t drained (256 arrivals)
    uint64_t* bar_o = bars + 8;    // [2] O_t ready in TMEM
    uint64_t* o_free = bars + 10;  // [2] O_t drained by the epilogue (256 arrivals)
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 16);
    float* sL = reinterpret_cast<float*>(tmem_base_smem + 4);       // [2 parities][2 slots][2 halves][128 lanes]
    float* sBiasW = sL + Cfg::kRedBytes / 4;                        // [H or 2][kWideBias]
    const bool bias_resident = H <= Cfg::kMaxResidentHeads;

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp_idx == 1) tmem_alloc(tmem_base_smem, 512);
    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_qkv);
        mbar_init(bar_qk, 1);
        mbar_init(bar_v, 1);
        mbar_init(qk_free, 1);
        mbar_init(v_free, 1);
        for (int t = 0; t < 2; ++t) {
            mbar_init(&bar_s[t], 1);
            mbar_init(&bar_p[t], 256);
            mbar_init(&bar_o[t], 1);
            mbar_init(&o_free[t], 256);
        }
        fence_barrier_init();
    }
    if (bias_resident) {
        // the bias table is a weight (written once at load time, not by the preceding kernel): it may be read before pdl_wait
        for (int i = threadIdx.x; i < H * Cfg::kWideBias; i += kAttn5Threads) {
            const int h = i / Cfg::kWideBias, w = i - h * Cfg::kWideBias;
            const int rel = max(-kAttnRelClamp, min(kAttnRelClamp, w - 255));
            sBiasW[i] = bias[h * kAttnBiasLen + rel + kAttnRelClamp] * 1.4426950408889634f;
        }
    }
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    const int stride = gridDim.x;
    // documents longer than len_limit belong to the mma.sync tile kernel launched next to this one (see enc_attention_tc2_kernel)
    auto qualify = [&](int it) {
        while (it < n_items) {
            const int dd = it / H;
            if (cu[dd + 1] - cu[dd] <= len_limit) break;
            it += stride;
        }
        return it;
    };

    if (warp_idx == 0) {
        if (lane == 0) {
            int item = qualify(blockIdx.x);
            int doc = item < n_items ? item / H : 0;
            int tok0 = cu[doc], tok1 = cu[doc + 1];
            for (int k = 0; item < n_items; ++k) {
                const int h = item - doc * H;
                const int nitem = qualify(item + stride), ndoc = nitem < n_items ? nitem / H : 0;
                const int ntok0 = cu[ndoc], ntok1 = cu[ndoc + 1];        // next item's extent: in flight during this one
                const int nkb_used = (tok1 - tok0 + 63) >> 6;
                if (k > 0) mbar_wait(qk_free, (k - 1) & 1);
                mbar_arrive_expect_tx(bar_qk, 2 * nkb_used * 8192);
                for (int b = 0; b < nkb_used; ++b) {
                    tma_load_2d(sQ + b * 8192, &tmap_qkv, bar_qk, h * 64, tok0 + b * 64, kEvictFirst);
                    tma_load_2d(sK + b * 8192, &tmap_qkv, bar_qk, inner + h * 64, tok0 + b * 64, kEvictFirst);
                }
                if (k > 0) mbar_wait(v_free, (k - 1) & 1);
                mbar_arrive_expect_tx(bar_v, nkb_used * 8192);
                for (int b = 0; b < nkb_used; ++b)
                    tma_load_2d(sV + b * 8192, &tmap_qkv, bar_v, 2 * inner + h * 64, tok0 + b * 64, kEvictFirst);
                item = nitem; doc = ndoc; tok0 = ntok0; tok1 = ntok1;
            }
        }
    } else if (warp_idx == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(128, 64 * NKB);
            constexpr uint32_t idesc_o = make_idesc_bf16_bmn(128, 64);
            uint32_t use0 = 0, use1 = 0;     // completed uses of tile slots 0 / 1
            int nt_prev = 0, nkb_prev = 0;
            int item = qualify(blockIdx.x);
            int len = 0;
            if (item < n_items) { const int doc = item / H; len = cu[doc + 1] - cu[doc]; }
            // iteration k issues, per tile slot, MMA-2 of item k-1 and then MMA-1 of item k; one extra iteration drains the last item
            for (int k = 0; item < n_items || nt_prev > 0; ++k) {
                const bool have = item < n_items;
                const int nitem = qualify(item + stride);
                int nlen = 0;
                if (have && nitem < n_items) { const int ndoc = nitem / H; nlen = cu[ndoc + 1] - cu[ndoc]; }
                const int nt_cur = have ? (len + 127) >> 7 : 0, nkb_cur = (len + 63) >> 6;
                if (have) { mbar_wait(bar_qk, k & 1); tc_fence_after(); }
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    uint32_t& use = t == 0 ? use0 : use1;
                    if (t < nt_prev) {
                        if (t == 0) mbar_wait(bar_v, (k - 1) & 1);
                        mbar_wait(&bar_p[t], use & 1);
                        if (use > 0) mbar_wait(&o_free[t], (use - 1) & 1);
                        tc_fence_after();
                        for (int kb = 0; kb < nkb_prev; ++kb) {
                            const uint64_t da = make_sw128_kmajor_desc(smem_u32(sP + t * Cfg::kPBytes + kb * 16384));
                            // V block: rows = keys (the MMA K dimension), 128 B of head dims contiguous = MN-major B operand;
                            // 16 keys per MMA = 2048 B -> +128 in (addr >> 4)
                            const uint64_t db = make_sw128_kmajor_desc(smem_u32(sV + kb * 8192));
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                umma_bf16(tmem_base + Cfg::kOCol0 + t * 64, da + 2 * kk, db + 128 * kk, idesc_o, (kb | kk) != 0);
                        }
                        umma_commit(&bar_o[t]);
                        ++use;
                        if (t == nt_prev - 1) umma_commit(v_free);
                    }
                    if (t < nt_cur) {
                        // tile slot 1 of an odd item: the A tile starts 64 query rows earlier, so that rows 128..191 of the document land on
                        // TMEM lanes 64..127 (the 64 rows below them repeat rows 64..127 and are ignored)
                        const int row0 = t * 128 - ((t == 1 && (k & 1)) ? 64 : 0);
                        const uint64_t da = make_sw128_kmajor_desc(smem_u32(sQ + row0 * 128));
                        const uint64_t db = make_sw128_kmajor_desc(smem_u32(sK));
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16(tmem_base + t * Cfg::kSCol, da + 2 * kk, db + 2 * kk, idesc_s, kk != 0);
                        umma_commit(&bar_s[t]);
                        if (t == nt_cur - 1) umma_commit(qk_free);
                    }
                }
                nt_prev = nt_cur; nkb_prev = nkb_cur;
                item = nitem; len = nlen;
            }
        }
    } else {
        const int sw = warp_idx - 2;                  // 0..15
        const int quarter = warp_idx & 3;             // TMEM lane quarter this warp may access
        const int t = (sw >> 2) & 1;                  // tile slot
        const int g = sw >> 3;                        // column half: score columns [g*kHalf, (g+1)*kHalf), output dims [32 g, 32 g + 32)
        const int grp_tid = (((sw & 3) | (g << 2)) << 5) | lane;   // 0..255 within the eight warps of this tile slot
        const int lane_row = quarter * 32 + lane;     // TMEM lane = row of the 128-row MMA tile
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t taddr_s = tmem_base + lane_off + t * Cfg::kSCol;
        const uint32_t taddr_o = tmem_base + lane_off + Cfg::kOCol0 + t * 64 + g * 32;
        uint8_t* prow = sP + t * Cfg::kPBytes + lane_row * 128;
        uint32_t use = 0;
        int h_loaded = -1;
        int item = qualify(blockIdx.x);
        int doc = item < n_items ? item / H : 0;
        int tok0 = cu[doc], tok1 = cu[doc + 1];
        // ---- deferred epilogue of the previous use of this tile slot: O_t / l -> bf16 -> global
        bool pend = false, p_rows = false, p_row_ok = false;
        int p_tok0 = 0, p_len = 0, p_h = 0, p_qi = 0;
        uint32_t p_par = 0;
        const float* p_sBq = nullptr;
        auto epilogue = [&]() {
            mbar_wait(&bar_o[t], p_par);
            tc_fence_after();
            if (p_rows) {
                uint32_t o[32];
                tmem_ld32(taddr_o, o);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&o_free[t]);
                const float* lb = sL + ((p_par * 2 + t) * 2) * 128;
                const float l = lb[lane_row] + lb[128 + lane_row];
                if (p_row_ok) {
                    if (l >= 7.888609e-31f && l < 1.2676506e30f) {     // [2^-100, 2^100): also false for inf / NaN
                        const float inv = 1.f / l;
                        uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(p_tok0 + p_qi) * ldo + p_h * 64 + g * 32);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint4 v;
                            v.x = pack_bf16(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
                            v.y = pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
                            v.z = pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
                            v.w = pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
                            dst[i] = v;
                        }
                    } else {
                        attn_slow_row(qkv, ld, inner, p_tok0, p_len, p_qi, p_h, g, p_sBq, out, ldo);
                    }
                }
            } else {
                mbar_arrive(&o_free[t]);
            }
        };
        for (int k = 0; item < n_items; ++k) {
            const int h = item - doc * H;
            const int len = tok1 - tok0, my_tok0 = tok0;
            const int nitem = qualify(item + stride), ndoc = nitem < n_items ? nitem / H : 0;
            const int ntok0 = cu[ndoc], ntok1 = cu[ndoc + 1];            // next item's extent: in flight during this one
            item = nitem; doc = ndoc; tok0 = ntok0; tok1 = ntok1;
            if (t >= ((len + 127) >> 7)) continue;
            const int ncols = ((len + 63) >> 6) * 64;
            const int shift = (t == 1 && (k & 1)) ? 64 : 0;       // see the MMA issuer
            const int qi = t * 128 + lane_row - shift;            // query row of this thread (0..255 whatever the lane)
            const bool row_ok = lane_row >= shift && qi < len;
            const bool rows = quarter * 32 >= shift && (t * 128 + quarter * 32 - shift) < len;   // warp-uniform: some real query row
            float* sB = sBiasW + (bias_resident ? h : t) * Cfg::kWideBias;
            if (!bias_resident && h != h_loaded) {
                // the previous window may still be needed by a pending slow-row epilogue of this group: run the epilogues first
                if (pend) { epilogue(); pend = false; }
                named_bar_sync(1 + t, 256);               // every warp of the slot's group is past its reads of the old window
                for (int i = grp_tid; i < 511; i += 256) {
                    const int rel = max(-kAttnRelClamp, min(kAttnRelClamp, i - 255));
                    sB[i] = bias[h * kAttnBiasLen + rel + kAttnRelClamp] * 1.4426950408889634f;
                }
                named_bar_sync(1 + t, 256);
                h_loaded = h;
            }
            const uint32_t sBrow = smem_u32(sB) + (255 - qi) * 4;   // [sBrow + 4 j] = log2(e) * bias(j - qi)
            // the epilogue of the previous use comes BEFORE the wait on S_t: MMA-2(k-1, t) was issued ahead of MMA-1(k, t), so O_t is the
            // older result; draining it overlaps MMA-1 and licenses the writes into P_t below
            if (pend) { epilogue(); pend = false; }
            mbar_wait(&bar_s[t], use & 1);
            tc_fence_after();
            float l = 0.f;
            if (rows) {
                const int c0 = g * Cfg::kHalf;
                const int c_hi = min(ncols, c0 + Cfg::kHalf);               // columns this thread must fill in the P tile
                const int nch = len > c0 ? (min(len, c0 + Cfg::kHalf) - c0 + 15) >> 4 : 0;   // 16-column chunks that hold real keys
                float l4[4] = {0.f, 0.f, 0.f, 0.f};
                // 16 keys = 32 B = two 16 B chunks of the 128B-swizzled K-major P tile (k-block c / 64)
                auto store_p = [&](const uint32_t (&packed)[8], int c) {
                    uint8_t* kblk = prow + (c >> 6) * 16384;
                    const int chunk0 = (c & 63) >> 3;
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        st_shared_v4(kblk + (((chunk0 + i) ^ (lane_row & 7)) << 4), packed[4 * i], packed[4 * i + 1], packed[4 * i + 2],
                                     packed[4 * i + 3]);
                };
                auto chunk_p = [&](const uint32_t (&r)[16], int c) {
                    uint32_t packed[8];
                    if (c + 16 <= len) {
#pragma unroll
                        for (int e = 0; e < 16; e += 2) {
                            const float p0 = ex2_approx(fmaf(__uint_as_float(r[e]), 1.4426950408889634f, ld_shared_f32(sBrow + 4 * (c + e))));
                            const float p1 = ex2_approx(fmaf(__uint_as_float(r[e + 1]), 1.4426950408889634f, ld_shared_f32(sBrow + 4 * (c + e + 1))));
                            l4[(e >> 1) & 3] += p0 + p1;
                            packed[e >> 1] = pack_bf16(p0, p1);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; e += 2) {
                            float p0 = ex2_approx(fmaf(__uint_as_float(r[e]), 1.4426950408889634f, ld_shared_f32(sBrow + 4 * (c + e))));
                            float p1 = ex2_approx(fmaf(__uint_as_float(r[e + 1]), 1.4426950408889634f, ld_shared_f32(sBrow + 4 * (c + e + 1))));
                            p0 = (c + e < len) ? p0 : 0.f;
                            p1 = (c + e + 1 < len) ? p1 : 0.f;
                            l4[(e >> 1) & 3] += p0 + p1;
                            packed[e >> 1] = pack_bf16(p0, p1);
                        }
                    }
                    store_p(packed, c);
                };
                {
                    // TMEM loads double-buffered in registers: chunk i+1 is in flight while chunk i is turned into P
                    uint32_t ra[16], rb[16];
                    if (nch > 0) tmem_ld16(taddr_s + c0, ra);
#pragma unroll 1
                    for (int i = 0; i < nch; i += 2) {
                        tmem_ld_wait();
                        if (i + 1 < nch) tmem_ld16(taddr_s + c0 + 16 * (i + 1), rb);
                        chunk_p(ra, c0 + 16 * i);
                        if (i + 1 < nch) {
                            tmem_ld_wait();
                            if (i + 2 < nch) tmem_ld16(taddr_s + c0 + 16 * (i + 2), ra);
                            chunk_p(rb, c0 + 16 * (i + 1));
                        }
                    }
                }
                {
                    const uint32_t zeros[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                    for (int c = c0 + 16 * nch; c < c_hi; c += 16) store_p(zeros, c);
                }
                l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
                sL[(((use & 1) * 2 + t) * 2 + g) * 128 + lane_row] = l;
            }
            tc_fence_before();      // TMEM reads of S_t are complete before MMA-1 of the next use overwrites it
            fence_proxy_async();    // generic-proxy smem writes -> visible to the tensor core (async proxy)
            mbar_arrive(&bar_p[t]);
            pend = true; p_rows = rows; p_row_ok = row_ok; p_tok0 = my_tok0; p_len = len; p_h = h; p_qi = qi; p_par = use & 1;
            p_sBq = sB + (255 - qi);
            ++use;
        }
        if (pend) epilogue();
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200
