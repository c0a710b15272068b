// Persistent tcgen05 encoder self-attention, round-2 design ("tc5", the default): 16 softmax warps, two threads per query row, ONE
// pass over TMEM with no shift, P handed to the second MMA through tensor memory; for documents of <= 192 tokens
// (modeling_t5.py:308-334: scores + position bias + mask -> fp32 softmax -> P.V; no 1/sqrt(d) scaling in T5).
//
// What the round-1 kernel (enc_attention_tc2_kernel, attention_tc.cuh) was bound by, measured on the B200 (profiles/r02_attn_profile.txt):
// 8 softmax warps, one thread per 192-column row, two passes through TMEM; ncu: issue slots 31 % busy, tensor pipe 12 %, the six
// warps that own real rows at S = 184 run at an IPC of 0.12 — per-thread latency chains, spilled loop state that is an L2 round trip
// away (210 KB of the SM's 228 KB were shared memory, so local memory does not stay in L1), and cu[] look-ups in every role.
// experiments/tmem_probe.cu: the TMEM read rate keeps growing to 16 warps (960 B/clk/SM vs 300-600 with 4-8 warps). So:
//   * 18 warps: warp 0 TMA, warp 1 MMA issue + TMEM allocation, warps 2..17 softmax. A softmax warp is (quarter q = warp_idx & 3: the 32
//     TMEM lanes it may touch, tile slot t, column half g): a query row's 192 score columns are shared by two threads (96 columns =
//     six 16-column tcgen05.ld chunks each); twelve of the sixteen warps own real rows at S = 184.
//   * ONE pass, NO shift: softmax is shift-invariant and fp32 / bf16 share one exponent range, so p = 2^v (v = log2(e) s + bias') is as
//     precise as 2^(v - max) while the row maximum lies within +-100 (powers of two: natural-log scores within +-69). No row
//     maximum, no write-back of v, no second TMEM read, no dependency between the two threads of a row ahead of P: per column
//     LDS (bias), FFMA, MUFU.EX2, FADD (row sum), half an F2FP. The two threads leave their partial sums in shared memory for the
//     deferred epilogue. A row whose total sum leaves [2^-100, 2^100) or is not finite — scores beyond +-69, which trained T5
//     checkpoints do not produce but the contract must survive — is NOT stored: its byte in a global bad-row map is set, and after
//     the item loop the CTA walks its items again and recomputes exactly the marked rows on CUDA cores with the true maximum
//     (attn_slow_row), clearing the map. The result is always the exact softmax (tests drive this with scores in the hundreds),
//     and the hot loop contains no call (a call in the loop made ptxas keep the loop state in local memory).
//   * P through tensor memory: each thread writes the bf16 P of its 16-key chunk back into the S_t columns it has just read
//     (tcgen05.st; two bf16 per column, always behind its own read pointer) and MMA-2 reads its A operand from tensor memory
//     (tcgen05.mma [d], [a_tmem], b_desc). No P tile in shared memory (96 KB), no proxy fence, and the N = 64 MMA-2 no longer
//     competes with the softmax warps' bias loads for the 128 B/clk shared-memory port.
//   * the TMA thread is the only role that walks cu[]: it publishes {first token, length, head} of item k in a 4-entry shared-memory
//     ring before arming the item's load barrier; the other roles read the ring after a barrier they wait on anyway. Every barrier
//     completes exactly once per item (parity k & 1), for both tile slots — a short document simply gets no MMAs on slot 1 — and
//     an invalid descriptor after the last item ends all roles.
//   * the second query tile of a 129..192-token document has at most 64 real rows: on odd items the MMA's A descriptor starts 64
//     rows earlier, so those rows land on TMEM lanes 64..127 (warps of quarters 2, 3) instead of lanes 0..63 — over two items
//     every SM sub-partition gets the same softmax work instead of quarters 0, 1 doing twice that of 2, 3.
//   * nothing derived from the thread index lives across the item loop (re-derived from %tid.x each iteration; shared-memory
//     accesses are explicit ld/st.shared on 32-bit addresses): zero local-memory traffic inside the loop.
// TMEM columns: S_t at t*192 (P_t inside it), O_t at 384 + 64 t. Item pipeline: TMA one item ahead; per tile slot MMA-2 of item k
// then MMA-1 of item k+1; epilogue of item k deferred to the start of item k+1. B200, 100 documents x 16 heads, S = 184:
// 2.00 ms per step (tc2) -> 1.44 ms; ncu in profiles/r02_attn_profile.txt. Tried and dropped on the way: splitting S_t into
// independently pipelined key halves (more barriers, Q read twice: slower), busy-poll / plain try_wait loops (no change).
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
    static constexpr int kWideBias = 512;                 // 511 used: index (j - i) + 255
    static constexpr int kRedBytes = 2 * 2 * 2 * 128 * 4; // partial row sums [use parity][tile slot][column half][lane]
    static constexpr int kFixedBytes = 3 * kQKVBytes + 12 * 8 + 4 * 16 + 16 + kRedBytes + 1024;
    static constexpr int kMaxHeads = (227 * 1024 - kFixedBytes) / (kWideBias * 4);     // 71: every T5 size (xxl has 64 heads)
    static constexpr int smem_bytes(int H) { return kFixedBytes + H * kWideBias * 4; }
    static constexpr int kSCol = 64 * NKB;
    static constexpr int kOCol0 = 2 * 64 * NKB;
    static constexpr int kHalf = 32 * NKB;                // score columns per thread
    static_assert(kOCol0 + 128 <= 512, "S and O tiles must fit the 512 TMEM columns side by side");
};

// Exact softmax(q_i K^T + bias) V for ONE query row and 32 of its 64 output dims on CUDA cores: the rare-row fallback of the
// one-pass kernel (row sum outside [2^-100, 2^100) or not finite). sBq[j] = log2(e) * bias(j - qi) (a pointer into the global
// [H][512] table). Called only from the fix-up walk AFTER the item loop: a call inside the loop makes ptxas keep the loop state
// in local memory (ABI), and local memory is an L2 round trip in this kernel (profiles/r02_attn_profile.txt).
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
    const float inv = __fdividef(1.f, l);
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

// bias_wide: [H][512] fp32, bias_wide[h][w] = log2(e) * bias_h(clamp(w - 255, +-128)) (built once at weight-load time: the per-launch
// prologue is a plain 16 B-vector copy into shared memory instead of 8 k dependent global loads).
// All roles wait with try_wait + a 64 ns back-off (plain try_wait loops and busy polling measured the same on the B200; the back-off
// keeps single-lane producer warps from burning the issue slots of the sub-partition they share with softmax warps).
__device__ __forceinline__ void attn5_wait(uint64_t* bar, uint32_t parity) { mbar_wait_backoff(bar, parity, 64); }

// P_t never touches shared memory: every thread writes the bf16 P of its keys back into tensor memory with tcgen05.st, into the S_t
// columns it has just read (keys [96 g, 96 g + 96) -> columns 96 g .. 96 g + 47 of S_t, two bf16 per column), and MMA-2 takes its A
// operand from tensor memory (tcgen05.mma with [a_tmem]): no st.shared of P, no shared-memory read of P by the tensor core (the
// N = 64 MMA-2 was shared-memory bound on its 4 KB A operand), no proxy fence; 1.54 -> 1.44 ms per step against the shared-memory
// version (profiles/r02_bench_attn_ab.txt). The bias windows of all heads (H <= 71) are resident in the shared memory this frees.
template <int NKB>
__global__ void __launch_bounds__(kAttn5Threads, 1)
enc_attention_tc5_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __nv_bfloat16* __restrict__ qkv, int ld, int inner,
                         const int* __restrict__ cu, const float* __restrict__ bias_wide, __nv_bfloat16* __restrict__ out, int ldo, int H,
                         int n_items, int len_limit, uint8_t* __restrict__ bad_map) {
    using Cfg = AttnTc5Cfg<NKB>;
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + Cfg::kQKVBytes;
    uint8_t* sV = sK + Cfg::kQKVBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + Cfg::kQKVBytes);
    uint64_t* bar_qk = bars + 0;   // Q/K of item k landed (or: the end-of-work descriptor was published)
    uint64_t* bar_v = bars + 1;
    uint64_t* qk_free = bars + 2;  // both MMA-1s of the item retired: Q/K smem reusable
    uint64_t* v_free = bars + 3;   // MMA-2s of the item retired: V smem reusable
    uint64_t* bar_s = bars + 4;    // [2] S_t ready in TMEM
    uint64_t* bar_p = bars + 6;    // [2] P_t written to smem, S_t and O_t drained (256 arrivals)
    uint64_t* bar_o = bars + 8;    // [2] O_t ready in TMEM
    int4* sDesc = reinterpret_cast<int4*>(bars + 12);               // [4] item ring: {first token, length, head, valid}
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(sDesc + 4);   // [0] TMEM base, [2] "some row needs the fix-up walk" (as a float)
    float* sL = reinterpret_cast<float*>(tmem_base_smem + 4);       // [2 parities][2 slots][2 halves][128 lanes]
    float* sBiasW = sL + Cfg::kRedBytes / 4;                        // [H][kWideBias]

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp_idx == 1) tmem_alloc(tmem_base_smem, 512);
    if (warp_idx == 0 && lane == 0) {
        tmem_base_smem[2] = 0u;
        tma_prefetch_desc(&tmap_qkv);
        mbar_init(bar_qk, 1);
        mbar_init(bar_v, 1);
        mbar_init(qk_free, 1);
        mbar_init(v_free, 1);
        for (int t = 0; t < 2; ++t) {
            mbar_init(&bar_s[t], 1);
            mbar_init(&bar_p[t], 256);
            mbar_init(&bar_o[t], 1);
        }
        fence_barrier_init();
    }
    {
        // the bias table is a weight (written once at load time, not by the preceding kernel): it may be read before pdl_wait
        const float4* src = reinterpret_cast<const float4*>(bias_wide);
        float4* dst = reinterpret_cast<float4*>(sBiasW);
        for (int i = threadIdx.x; i < H * (Cfg::kWideBias / 4); i += kAttn5Threads) dst[i] = __ldg(src + i);
    }
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    // Every barrier above completes exactly once per item k (parity k & 1), for every item and for both tile slots: a document of at
    // most 128 tokens simply gets no MMAs on slot 1. The TMA thread is the only role that walks cu[]: it publishes {tok0, len, head}
    // of item k in sDesc[k & 3] before arming bar_qk(k); the MMA thread reads it after bar_qk(k), the softmax warps after bar_s(k).
    // After the last item it publishes an invalid descriptor, which flows down the same barriers and ends the other roles.
    if (warp_idx == 0) {
        if (lane == 0) {
            const int stride = gridDim.x;
            int k = 0;
            for (int item = blockIdx.x; item < n_items; item += stride) {
                const int doc = item / H, h = item - doc * H;
                const int tok0 = cu[doc], len = cu[doc + 1] - tok0;
                if (len > len_limit) continue;   // longer documents belong to the mma.sync tile kernel launched next to this one
                const int nkb_used = (len + 63) >> 6;
                if (k > 0) attn5_wait(qk_free, (k - 1) & 1);
                sDesc[k & 3] = make_int4(tok0, len, h, 1);
                mbar_arrive_expect_tx(bar_qk, 2 * nkb_used * 8192);
                for (int b = 0; b < nkb_used; ++b) {
                    tma_load_2d(sQ + b * 8192, &tmap_qkv, bar_qk, h * 64, tok0 + b * 64, kEvictFirst);
                    tma_load_2d(sK + b * 8192, &tmap_qkv, bar_qk, inner + h * 64, tok0 + b * 64, kEvictFirst);
                }
                if (k > 0) attn5_wait(v_free, (k - 1) & 1);
                mbar_arrive_expect_tx(bar_v, nkb_used * 8192);
                for (int b = 0; b < nkb_used; ++b)
                    tma_load_2d(sV + b * 8192, &tmap_qkv, bar_v, 2 * inner + h * 64, tok0 + b * 64, kEvictFirst);
                ++k;
            }
            if (k > 0) attn5_wait(qk_free, (k - 1) & 1);
            sDesc[k & 3] = make_int4(0, 0, 0, 0);
            mbar_arrive(bar_qk);
        }
    } else if (warp_idx == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(128, 64 * NKB);
            constexpr uint32_t idesc_o = make_idesc_bf16_bmn(128, 64);
            int len_prev = 0;
            for (int k = 0;; ++k) {
                attn5_wait(bar_qk, k & 1);
                tc_fence_after();
                const int4 dsc = sDesc[k & 3];
                const bool have = dsc.w != 0;
                const int len = dsc.y;
                const int nt = have ? (len + 127) >> 7 : 0;
                const int nt_prev = (len_prev + 127) >> 7, nkb_prev = (len_prev + 63) >> 6;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (k > 0) {
                        // MMA-2 of item k-1: O_t = P_t . V. bar_p also says that the slot's threads are done with O_t of item k-2
                        // (their deferred epilogue precedes their pass in program order) and with S_t of item k-1.
                        if (t == 0) attn5_wait(bar_v, (k - 1) & 1);
                        attn5_wait(&bar_p[t], (k - 1) & 1);
                        tc_fence_after();
                        if (t < nt_prev) {
                            for (int kb = 0; kb < nkb_prev; ++kb) {
                                // V block: rows = keys (the MMA K dimension), 128 B of head dims contiguous = MN-major B operand;
                                // 16 keys per MMA = 2048 B -> +128 in (addr >> 4)
                                const uint64_t db = make_sw128_kmajor_desc(smem_u32(sV + kb * 8192));
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    const int key0 = kb * 64 + kk * 16;          // P of keys [key0, key0 + 16): 8 columns inside S_t
                                    const int pcol = key0 < Cfg::kHalf ? (key0 >> 1) : Cfg::kHalf + ((key0 - Cfg::kHalf) >> 1);
                                    umma_bf16_ts(tmem_base + Cfg::kOCol0 + t * 64, tmem_base + t * Cfg::kSCol + pcol, db + 128 * kk, idesc_o, (kb | kk) != 0);
                                }
                            }
                        }
                        umma_commit(&bar_o[t]);
                        if (t == 1) umma_commit(v_free);
                    }
                    if (t < nt) {
                        // tile slot 1 of an odd item: the A tile starts 64 query rows earlier, so that rows 128..191 of the document land on
                        // TMEM lanes 64..127 (the 64 rows below them repeat rows 64..127 and are ignored)
                        const int row0 = t * 128 - ((t == 1 && (k & 1)) ? 64 : 0);
                        const uint64_t da = make_sw128_kmajor_desc(smem_u32(sQ + row0 * 128));
                        const uint64_t db = make_sw128_kmajor_desc(smem_u32(sK));
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16(tmem_base + t * Cfg::kSCol, da + 2 * kk, db + 2 * kk, idesc_s, kk != 0);
                    }
                    umma_commit(&bar_s[t]);
                    if (t == 1 && have) umma_commit(qk_free);
                }
                if (!have) break;
                len_prev = len;
            }
        }
    } else {
        // Everything this role keeps across the item loop is a handful of 32-bit shared-space addresses: values derived from the
        // thread index are cheap to recompute, a spilled one costs an L2 round trip here (210 KB of the SM's 228 KB are shared
        // memory, so local memory does not stay in L1) — profiles/r02_attn_profile.txt.
        // The thread's coordinates are re-derived from %tid.x (read through a volatile asm, so the compiler cannot hoist them out of
        // the loop and then spill them) at the two places that use them.
        auto tid_now = []() { uint32_t v; asm volatile("mov.u32 %0, %%tid.x;" : "=r"(v)); return static_cast<int>(v); };
        const uint32_t smem_s = smem_u32(smem);      // everything below is an offset from this one shared-space address
        constexpr uint32_t kBarOff = 3 * Cfg::kQKVBytes, kDescOff = kBarOff + 12 * 8,
                           kLOff = kDescOff + 4 * 16 + 16, kBiasOff = kLOff + Cfg::kRedBytes;
        constexpr int c0_of_g = Cfg::kHalf;
        int4 prev = make_int4(0, 0, 0, 0);            // descriptor of the item whose epilogue is pending
        for (int k = 0;; ++k) {
            const int wi = tid_now() >> 5, ln = tid_now() & 31;
            const int sw = wi - 2;                        // 0..15
            const int quarter = wi & 3;                   // TMEM lane quarter this warp may access
            const int t = (sw >> 2) & 1;                  // tile slot
            const int g = sw >> 3;                        // column half: score columns [g*kHalf, (g+1)*kHalf), output dims [32 g, 32 g + 32)
            const int lane_row = quarter * 32 + ln;       // TMEM lane = row of the 128-row MMA tile
            const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
            const uint32_t taddr_s = tmem_base + lane_off + t * Cfg::kSCol + g * Cfg::kHalf;
            const uint32_t taddr_o = tmem_base + lane_off + Cfg::kOCol0 + t * 64 + g * 32;
            const uint32_t sL_s = smem_s + kLOff + (t * 256 + lane_row) * 4;                // partial sums of this row: halves at +0 / +512 B, parity at +2048 B
            const uint32_t sDesc_s = smem_s + kDescOff;
            const uint32_t sBias_s = smem_s + kBiasOff;
            // ---- deferred epilogue of item k-1 on this slot: O_t / l -> bf16 -> global. It comes BEFORE the wait on S_t(k): MMA-2(k-1, t)
            // was issued ahead of MMA-1(k, t), and passing bar_o licenses the writes into P_t below.
            if (prev.w) {
                const int kp = k - 1;
                attn5_wait(&bar_o[t], kp & 1);
                tc_fence_after();
                const int shift = (t == 1 && (kp & 1)) ? 64 : 0;
                const int qi = t * 128 + lane_row - shift;
                const int len = prev.y;
                if (quarter * 32 >= shift && (t * 128 + quarter * 32 - shift) < len) {     // warp-uniform: this warp owns real query rows
                    uint32_t o[32];
                    tmem_ld32(taddr_o, o);
                    const uint32_t la = sL_s + (kp & 1) * 2048;
                    const float l = ld_shared_f32(la) + ld_shared_f32(la + 512);
                    tmem_ld_wait();
                    if (qi < len) {
                        if (l >= 7.888609e-31f && l < 1.2676506e30f) {     // [2^-100, 2^100): also false for inf / NaN
                            const float inv = __fdividef(1.f, l);   // l in [2^-100, 2^100): 2 ulp, and no slow-path call in the loop
                            uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(prev.x + qi) * ldo + prev.z * 64 + g * 32);
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
                            // rare: leave the row to the fix-up walk after the item loop (both half-row threads see the same l and mark the same byte)
                            bad_map[static_cast<size_t>(prev.x + qi) * H + prev.z] = 1;
                            st_shared_f32(smem_s + kDescOff + 4 * 16 + 8, 1.f);
                        }
                    }
                }
            }
            attn5_wait(&bar_s[t], k & 1);
            tc_fence_after();
            const int4 dsc = ld_shared_v4_s32(sDesc_s + (k & 3) * 16);
            if (!dsc.w) break;
            const int len = dsc.y, h = dsc.z;
            const int shift = (t == 1 && (k & 1)) ? 64 : 0;       // see the MMA issuer
            const int qi = t * 128 + lane_row - shift;            // query row of this thread (0..255 whatever the lane)
            const bool rows = quarter * 32 >= shift && (t * 128 + quarter * 32 - shift) < len;   // warp-uniform: some real query row
            if (rows) {
                const int c0 = g * c0_of_g;
                // [sBrow + 4 c] = log2(e) * bias(c0 + c - qi)
                const uint32_t sBrow = sBias_s + (h * Cfg::kWideBias + 255 - qi + c0) * 4;
                const int n_real = min(len - c0, Cfg::kHalf);                   // my columns that hold real keys (may be <= 0)
                const int n_fill = min(((len + 63) & ~63) - c0, Cfg::kHalf);    // my columns of the P tile the MMA will read
                float l0 = 0.f, l1 = 0.f;
                // keys c0 + c .. + 15 of this row -> 8 packed columns right behind the thread's read pointer in its own half of S_t
                auto store_p = [&](uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5, uint32_t a6, uint32_t a7, int c) {
                    tmem_st8(taddr_s + (c >> 1), a0, a1, a2, a3, a4, a5, a6, a7);
                };
                int c = 0;
                // full chunks: no masking in the inner loop (LDS, FFMA, MUFU.EX2, FADD, half an F2FP per column)
#pragma unroll 1
                for (; c + 16 <= n_real; c += 16) {
                    uint32_t r[16], pk[8];
                    tmem_ld16(taddr_s + c, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; e += 2) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(r[e]), 1.4426950408889634f, ld_shared_f32(sBrow + 4 * (c + e))));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(r[e + 1]), 1.4426950408889634f, ld_shared_f32(sBrow + 4 * (c + e + 1))));
                        if (e & 2) l1 += p0 + p1; else l0 += p0 + p1;
                        pk[e >> 1] = pack_bf16(p0, p1);
                    }
                    store_p(pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7], c);
                }
                if (c < n_real) {       // the chunk that holds the document's last keys: columns past them contribute p = 0
                    uint32_t r[16], pk[8];
                    tmem_ld16(taddr_s + c, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; e += 2) {
                        float p0 = ex2_approx(fmaf(__uint_as_float(r[e]), 1.4426950408889634f, ld_shared_f32(sBrow + 4 * (c + e))));
                        float p1 = ex2_approx(fmaf(__uint_as_float(r[e + 1]), 1.4426950408889634f, ld_shared_f32(sBrow + 4 * (c + e + 1))));
                        p0 = (c + e < n_real) ? p0 : 0.f;
                        p1 = (c + e + 1 < n_real) ? p1 : 0.f;
                        l0 += p0 + p1;
                        pk[e >> 1] = pack_bf16(p0, p1);
                    }
                    store_p(pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7], c);
                    c += 16;
                }
                for (c = max(c, 0); c < n_fill; c += 16) store_p(0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, c);
                st_shared_f32(sL_s + (k & 1) * 2048 + g * 512, l0 + l1);
            }
            tmem_st_wait();         // this thread's P columns have landed in tensor memory
            tc_fence_before();      // ... and its TMEM reads of S_t (and of O_t in the epilogue above) are complete before the MMAs touch them
            mbar_arrive(&bar_p[t]);
            prev = dsc;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    // ---- fix-up walk (cold): rows whose unshifted row sum left [2^-100, 2^100) were marked in bad_map instead of being stored. Walk this
    // CTA's items again and recompute exactly those rows on CUDA cores with the true row maximum; the marks are cleared on the way, so
    // the map is all-zero again when the kernel ends. Which rows take this path depends on the row alone.
    if (tmem_base_smem[2] != 0u && warp_idx >= 2) {
        const int stid = threadIdx.x - 64;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int doc = item / H, h = item - doc * H;
            const int tok0 = cu[doc], len = cu[doc + 1] - tok0;
            if (len > len_limit) continue;
            for (int task0 = 0; task0 < 2 * len; task0 += kAttn5Threads - 64) {
                const int task = task0 + stid, qi = task >> 1, g = task & 1;     // the two halves of a row sit in adjacent lanes
                const bool mine = task < 2 * len && bad_map[static_cast<size_t>(tok0 + qi) * H + h] != 0;
                if (mine) attn_slow_row(qkv, ld, inner, tok0, len, qi, h, g, bias_wide + h * Cfg::kWideBias + (255 - qi), out, ldo);
                __syncwarp();
                if (mine && g == 0) bad_map[static_cast<size_t>(tok0 + qi) * H + h] = 0;
            }
        }
    }
}

}  // namespace b200
