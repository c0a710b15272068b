// Round-1 persistent tcgen05 encoder self-attention (B200RANK_ATTN=tc2): one thread per query row, two passes over TMEM. Kept as the
// validated A/B partner of the round-2 kernel (attention_tc5.cuh, the default), which shares its TMA / MMA roles and TMEM layout;
// the exploratory variants of round 1 (unpipelined tc, row-split tc3, provisional-shift tc4, mma.sync regs / resident) were removed
// once tc5 had beaten all of them on the B200 (profiles/r02_bench_attn_ab.txt).
// No 1/sqrt(d) scaling (modeling_t5.py:308); bias + mask + fp32 softmax as modeling_t5.py:313-334; padded keys do not
// exist in the packed layout (rows past `len` belong to the next document: they are masked to -inf / never stored).
#pragma once
#include <cuda.h>
#include "attention_enc.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int kAttnTcThreads = 320;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: softmax tile 0, warps 6-9: softmax tile 1

// Instruction descriptor with an MN-major B operand (bit 16), otherwise as make_idesc_bf16.
__host__ __device__ constexpr uint32_t make_idesc_bf16_bmn(uint32_t m, uint32_t n) {
    return make_idesc_bf16(m, n) | (1u << 16);
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent variant (documents of at most 64*NKB <= 192 tokens): one CTA per SM walks (document, head) work items
// round-robin, and the three roles run one item apart from each other:
//   * the TMA thread refills Q/K as soon as both MMA-1s of the previous item have retired (qk_free) and V as soon as its
//     MMA-2s have (v_free), so the loads of item k+1 fly during the softmax of item k;
//   * S_t and O_t own separate TMEM columns (2*64*NKB + 2*64 <= 512) and the MMA thread issues, per tile slot t,
//     MMA-2(item k, t) immediately followed by MMA-1(item k+1, t): a warpgroup finds its next S tile ready when it comes
//     back from the epilogue, whatever the other warpgroup is doing;
//   * the softmax is two passes over TMEM with the bias added ONCE: pass 1 computes v = s*log2(e) + bias' (bias' pre-scaled,
//     laid out as a 511-entry window per head so that no clamp is needed), masks padded keys to -inf, keeps the row maximum
//     and writes v back into the S columns (tcgen05.st); pass 2 is ex2(v - m), row sum, bf16 pack, swizzled st.shared.
//     TMEM loads are double-buffered in registers and the max / sum chains are split four ways (the kernel is bound by
//     the serial latency of one thread walking its row, not by issue slots: ncu in profiles/r01_ncu_summary_final.txt).
// Barrier parities: bar_qk / bar_v / qk_free / v_free complete once per item; the per-tile barriers (bar_s, bar_p, bar_o,
// o_free) complete once per USE of tile slot t (an item shorter than 129 tokens does not touch slot 1), and every role
// derives the same use counts from cu[]. Warps whose 32 query rows lie past the document still walk the barrier sequence
// (without touching TMEM) so that no warp can arrive twice in one phase.
template <int NKB>
struct AttnTc2Cfg {
    static constexpr int kRows = 64 * NKB;
    static constexpr int kQKVBytes = kRows * 128;
    static constexpr int kPBytes = NKB * 128 * 128;
    static constexpr int kWideBias = 512;                 // 511 used: index (j - i) + 255
    static constexpr int kFixedBytes = 3 * kQKVBytes + 2 * kPBytes + 16 * 8 + 16 + 2048 + 1024;
    static constexpr int kMaxResidentHeads = (227 * 1024 - kFixedBytes) / (kWideBias * 4);
    // bias windows: one per head when they all fit next to the tiles (resident for the whole kernel), else one per warpgroup
    static constexpr int smem_bytes(int H) { return kFixedBytes + (H <= kMaxResidentHeads ? (H < 2 ? 2 : H) : 2) * kWideBias * 4; }
    static constexpr int kSCol = 64 * NKB;                // S_t at columns t * kSCol
    static constexpr int kOCol0 = 2 * 64 * NKB;           // O_t at columns kOCol0 + 64 t
    static_assert(kOCol0 + 128 <= 512, "S and O tiles must fit the 512 TMEM columns side by side");
};

template <int NKB>
__global__ void __launch_bounds__(kAttnTcThreads, 1)
enc_attention_tc2_kernel(const __grid_constant__ CUtensorMap tmap_qkv, int inner, const int* __restrict__ cu,
                         const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int ldo, int H, int n_items, int spin,
                         int len_limit) {
    using Cfg = AttnTc2Cfg<NKB>;
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
    uint64_t* v_free = bars + 3;   // MMA-2s of the item retired: V (and P) smem reusable
    uint64_t* bar_s = bars + 4;    // [2] S_t ready in TMEM
    uint64_t* bar_p = bars + 6;    // [2] P_t written to smem, S_t drained (128 arrivals)
    uint64_t* bar_o = bars + 8;    // [2] O_t ready in TMEM
    uint64_t* o_free = bars + 10;  // [2] O_t drained by the epilogue (128 arrivals)
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 16);
    float* sBiasW = reinterpret_cast<float*>(tmem_base_smem + 4) + 512;                                     // [H or 2][kWideBias]
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
            mbar_init(&bar_p[t], 128);
            mbar_init(&bar_o[t], 1);
            mbar_init(&o_free[t], 128);
        }
        fence_barrier_init();
    }
    if (bias_resident) {
        // the bias table is a weight (written once at load time, not by the preceding kernel): it may be read before pdl_wait
        for (int i = threadIdx.x; i < H * Cfg::kWideBias; i += kAttnTcThreads) {
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
    // Documents longer than len_limit (= 64 * NKB) belong to the mma.sync tile kernel launched next to this one: every role walks
    // only the qualifying items, so the barrier bookkeeping never sees the others. Which kernel a document gets is thereby a
    // property of the document, not of the batch it happens to be scored in (bit-identical results across batch compositions).
    auto qualify = [&](int it) {
        while (it < n_items) {
            const int dd = it / H;
            if (cu[dd + 1] - cu[dd] <= len_limit) break;
            it += stride;
        }
        return it;
    };
    auto wait = [&](uint64_t* bar, uint32_t parity) { if (spin) mbar_wait_spin(bar, parity); else mbar_wait(bar, parity); };

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
                if (k > 0) wait(qk_free, (k - 1) & 1);
                mbar_arrive_expect_tx(bar_qk, 2 * nkb_used * 8192);
                for (int b = 0; b < nkb_used; ++b) {
                    tma_load_2d(sQ + b * 8192, &tmap_qkv, bar_qk, h * 64, tok0 + b * 64, kEvictFirst);
                    tma_load_2d(sK + b * 8192, &tmap_qkv, bar_qk, inner + h * 64, tok0 + b * 64, kEvictFirst);
                }
                if (k > 0) wait(v_free, (k - 1) & 1);
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
                if (have) { wait(bar_qk, k & 1); tc_fence_after(); }
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    uint32_t& use = t == 0 ? use0 : use1;
                    if (t < nt_prev) {
                        if (t == 0) wait(bar_v, (k - 1) & 1);
                        wait(&bar_p[t], use & 1);
                        if (use > 0) wait(&o_free[t], (use - 1) & 1);
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
                        // S_t was drained by the softmax of its previous use: this thread waited on that bar_p before the MMA-2 above
                        const uint64_t da = make_sw128_kmajor_desc(smem_u32(sQ + t * 128 * 128));
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
        const int t = (warp_idx - 2) >> 2;            // query tile slot of this warpgroup
        const int quarter = warp_idx & 3;             // TMEM lane quarter of this warp
        const int wg_tid = threadIdx.x - 64 - t * 128;
        const int row_in_tile = quarter * 32 + lane;
        const int qi = t * 128 + row_in_tile;         // query index inside the document
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t taddr_s = tmem_base + lane_off + t * Cfg::kSCol;
        const uint32_t taddr_o = tmem_base + lane_off + Cfg::kOCol0 + t * 64;
        uint8_t* prow = sP + t * Cfg::kPBytes + row_in_tile * 128;
        uint32_t use = 0;
        int h_loaded = -1;
        int item = qualify(blockIdx.x);
        int doc = item < n_items ? item / H : 0;
        int tok0 = cu[doc], tok1 = cu[doc + 1];
        // ---- deferred epilogue of the previous item on this tile slot: O_t / l -> bf16 -> global
        bool pend = false, p_active = false;
        float p_l = 0.f;
        int p_tok0 = 0, p_len = 0, p_h = 0;
        uint32_t p_par = 0;
        auto epilogue = [&]() {
            wait(&bar_o[t], p_par);
            tc_fence_after();
            if (p_active) {
                uint32_t o0[32], o1[32];
                tmem_ld32(taddr_o, o0);
                tmem_ld32(taddr_o + 32, o1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&o_free[t]);
                if (qi < p_len) {
                    const float inv = 1.f / p_l;
                    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(p_tok0 + qi) * ldo + p_h * 64);
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
            } else {
                mbar_arrive(&o_free[t]);
            }
        };
        while (item < n_items) {
            const int h = item - doc * H;
            const int len = tok1 - tok0, my_tok0 = tok0;
            const int nitem = qualify(item + stride), ndoc = nitem < n_items ? nitem / H : 0;
            const int ntok0 = cu[ndoc], ntok1 = cu[ndoc + 1];            // next item's extent: in flight during this one
            item = nitem; doc = ndoc; tok0 = ntok0; tok1 = ntok1;
            if (t >= ((len + 127) >> 7)) continue;
            const int ncols = ((len + 63) >> 6) * 64;
            const bool active = (t * 128 + quarter * 32) < len;   // warp-uniform: at least one real query row
            float* sB = sBiasW + (bias_resident ? h : t) * Cfg::kWideBias;
            if (!bias_resident && h != h_loaded) {
                named_bar_sync(1 + t, 128);               // every warp of the group is past its reads of the old window
                for (int i = wg_tid; i < 511; i += 128) {
                    const int rel = max(-kAttnRelClamp, min(kAttnRelClamp, i - 255));
                    sB[i] = bias[h * kAttnBiasLen + rel + kAttnRelClamp] * 1.4426950408889634f;
                }
                named_bar_sync(1 + t, 128);
                h_loaded = h;
            }
            const uint32_t sBrow = smem_u32(sB) + (255 - qi) * 4;   // [sBrow + 4 j] = log2(e) * bias(j - qi)
            wait(&bar_s[t], use & 1);
            tc_fence_after();
            float m = -INFINITY, l = 0.f;
            if (active) {
                // ---- pass 1: v = s*log2(e) + bias', padded keys -> -inf, row max, v back into TMEM. Two register buffers:
                // the TMEM load of chunk c+1 is in flight while chunk c is processed.
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                auto pass1 = [&](uint32_t (&r)[32], int c) {
                    if (c + 32 <= len) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const float v = fmaf(__uint_as_float(r[e]), 1.4426950408889634f, ld_shared_f32(sBrow + 4 * (c + e)));
                            m4[e & 3] = fmaxf(m4[e & 3], v);
                            r[e] = __float_as_uint(v);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            float v = fmaf(__uint_as_float(r[e]), 1.4426950408889634f, ld_shared_f32(sBrow + 4 * (c + e)));
                            v = (c + e < len) ? v : -INFINITY;
                            m4[e & 3] = fmaxf(m4[e & 3], v);
                            r[e] = __float_as_uint(v);
                        }
                    }
                    tmem_st32(taddr_s + c, r);
                };
                {
                    uint32_t ra[32], rb[32];
                    tmem_ld32(taddr_s, ra);
#pragma unroll 1
                    for (int c = 0; c < len; c += 64) {
                        tmem_ld_wait();
                        if (c + 32 < len) tmem_ld32(taddr_s + c + 32, rb);
                        pass1(ra, c);
                        if (c + 32 < len) {
                            tmem_ld_wait();
                            if (c + 64 < len) tmem_ld32(taddr_s + c + 64, ra);
                            pass1(rb, c + 32);
                        }
                    }
                }
                m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            }
            // The previous item's epilogue runs HERE, between the passes: by now its MMA-2 has long retired (no wait on the
            // tensor core), and passing bar_o also licenses pass 2 to overwrite the P_t tile that MMA-2 was reading.
            if (pend) { epilogue(); pend = false; }
            if (active) {
                tmem_st_wait();
                // ---- pass 2: p = 2^(v - m), row sum, bf16 P into the swizzled A-operand tile (zeros past the document)
                float l4[4] = {0.f, 0.f, 0.f, 0.f};
                auto store_p = [&](const uint32_t (&packed)[16], int c) {
                    uint8_t* kblk = prow + (c >> 6) * 16384;
                    const int chunk0 = (c & 63) >> 3;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        st_shared_v4(kblk + (((chunk0 + i) ^ (row_in_tile & 7)) << 4), packed[4 * i], packed[4 * i + 1],
                                     packed[4 * i + 2], packed[4 * i + 3]);
                };
                auto pass2 = [&](const uint32_t (&r)[32], int c) {
                    uint32_t packed[16];
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                        const float p0 = ex2_approx(__uint_as_float(r[e]) - m);
                        const float p1 = ex2_approx(__uint_as_float(r[e + 1]) - m);
                        l4[(e >> 1) & 3] += p0 + p1;
                        packed[e >> 1] = pack_bf16(p0, p1);
                    }
                    store_p(packed, c);
                };
                {
                    uint32_t ra[32], rb[32];
                    tmem_ld32(taddr_s, ra);
#pragma unroll 1
                    for (int c = 0; c < len; c += 64) {
                        tmem_ld_wait();
                        if (c + 32 < len) tmem_ld32(taddr_s + c + 32, rb);
                        pass2(ra, c);
                        if (c + 32 < len) {
                            tmem_ld_wait();
                            if (c + 64 < len) tmem_ld32(taddr_s + c + 64, ra);
                            pass2(rb, c + 32);
                        }
                    }
                    const uint32_t zeros[16] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                    for (int c = (len + 31) & ~31; c < ncols; c += 32) store_p(zeros, c);
                }
                l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
            }
            tc_fence_before();      // TMEM reads/writes of S_t are complete before MMA-1 of the next use overwrites it
            fence_proxy_async();    // generic-proxy smem writes -> visible to the tensor core (async proxy)
            mbar_arrive(&bar_p[t]);
            pend = true; p_active = active; p_l = l; p_tok0 = my_tok0; p_len = len; p_h = h; p_par = use & 1;
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
