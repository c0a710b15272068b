// Encoder self-attention over packed variable-length documents (no padded keys exist, so the
// key-padding mask of modeling_t5.py:703-726 is implicit): fused  Q.K^T + relative-position bias
// -> fp32 online softmax -> P.V  per (document, head, 64-query tile). No 1/sqrt(d) scaling
// (modeling_t5.py:308). Scores never touch HBM.
//
// Round-1 implementation: bf16 mma.sync m16n8k16 with fp32 accumulation (4 warps x 16 query rows,
// 64-key blocks, cp.async double buffering, XOR-swizzled shared tiles + ldmatrix). Attention is
// 2.9 % of the encoder FLOPs at S = 184; the tcgen05 version is a later-round item.
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int kAttnRelClamp = 128;                     // |j - i| >= 128 all share one bucket (max_distance)
constexpr int kAttnBiasLen = 2 * kAttnRelClamp + 1;    // per-head table indexed by clamp(j - i) + 128

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t d = smem_u32(smem_dst);
    const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// element offset of (row, 8-element chunk) inside a [64][64] bf16 tile with 16 B chunks XOR-swizzled by row
__device__ __forceinline__ int sw_off(int row, int chunk) { return row * 64 + ((chunk ^ (row & 7)) << 3); }

// Loads a [64 rows][64 dims] bf16 tile (rows row0.. of this document, zero-filled past `len`).
__device__ __forceinline__ void load_tile_64x64(__nv_bfloat16* smem_tile, const __nv_bfloat16* gbase, size_t ld,
                                                int row0, int len, int tid) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int idx = tid + it * 128;
        const int r = idx >> 3, c = idx & 7;
        const bool ok = (row0 + r) < len;
        const __nv_bfloat16* src = gbase + static_cast<size_t>(ok ? (row0 + r) : 0) * ld + c * 8;
        cp_async16(smem_tile + sw_off(r, c), src, ok);
    }
}

// grid (q_tiles, H, n_docs), 128 threads.
// qkv: packed [tokens, ld] with q | k | v column blocks each `inner` wide; head h at columns h*64.
// bias: [H][kAttnBiasLen] fp32, index clamp(j - i, -128, 128) + 128  (bidirectional buckets).
__global__ void __launch_bounds__(128, 4)
enc_attention_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, int inner, const int* __restrict__ cu,
                     const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int ldo, int skip_upto) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(128) __nv_bfloat16 sQ[64 * 64];
    __shared__ __align__(128) __nv_bfloat16 sK[2][64 * 64];
    __shared__ __align__(128) __nv_bfloat16 sV[2][64 * 64];
    __shared__ float sBias[kAttnBiasLen];

    const int qt = blockIdx.x, h = blockIdx.y, doc = blockIdx.z;
    const int tok0 = cu[doc];
    const int len = cu[doc + 1] - tok0;
    const int q0 = qt * 64;
    if (q0 >= len || len <= skip_upto) return;   // skip_upto > 0: the short documents of a mixed batch go to the tcgen05 kernel
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;

    const __nv_bfloat16* gq = qkv + static_cast<size_t>(tok0) * ld + h * 64;
    const __nv_bfloat16* gk = gq + inner;
    const __nv_bfloat16* gv = gq + 2 * inner;

    for (int i = tid; i < kAttnBiasLen; i += 128) sBias[i] = bias[h * kAttnBiasLen + i];

    const int nkb = (len + 63) / 64;
    load_tile_64x64(sQ, gq, ld, q0, len, tid);
    load_tile_64x64(sK[0], gk, ld, 0, len, tid);
    load_tile_64x64(sV[0], gv, ld, 0, len, tid);
    cp_async_commit();

    uint32_t qf[4][4];
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    const int qi0 = q0 + warp * 16 + g;  // query index (within doc) of accumulator rows c0/c1; +8 for c2/c3

    for (int kb = 0; kb < nkb; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nkb) {
            load_tile_64x64(sK[buf ^ 1], gk, ld, (kb + 1) * 64, len, tid);
            load_tile_64x64(sV[buf ^ 1], gv, ld, (kb + 1) * 64, len, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        if (kb == 0) {
            // Q fragments (A operand): rows warp*16 + (lane % 16), 8-element chunk 2*ks + lane/16
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(qf[ks], sQ + sw_off(warp * 16 + (lane & 15), 2 * ks + (lane >> 4)));
        }

        // ---- S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                // lanes 0-7: keys np*16+0..7 @ d-chunk 2ks ; 8-15: same keys @ 2ks+1 ; 16-23: keys +8 @ 2ks ; 24-31: keys +8 @ 2ks+1
                uint32_t kf[4];
                const int krow = np * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int kch = 2 * ks + ((lane >> 3) & 1);
                ldmatrix_x4(kf, sK[buf] + sw_off(krow, kch));
                mma_bf16_16816(s[2 * np], qf[ks], kf[0], kf[1]);
                mma_bf16_16816(s[2 * np + 1], qf[ks], kf[2], kf[3]);
            }
        }

        // ---- bias, key-length mask, online softmax
        const int kbase = kb * 64;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = kbase + nt * 8 + 2 * t4 + (e & 1);
                const int qi = qi0 + ((e >> 1) << 3);
                int rel = j - qi;
                rel = max(-kAttnRelClamp, min(kAttnRelClamp, rel));
                float v = s[nt][e] + sBias[rel + kAttnRelClamp];
                v = (j < len) ? v : -INFINITY;
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
        for (int dt = 0; dt < 8; ++dt) {
            o[dt][0] *= scale[0];
            o[dt][1] *= scale[0];
            o[dt][2] *= scale[1];
            o[dt][3] *= scale[1];
        }

        // ---- O += P V   (V^T fragments via ldmatrix.trans)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {   // 16 keys per step
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {  // pairs of 8-wide d tiles
                // lanes 0-7: keys ks*16+0..7 @ d-chunk 2dp ; 8-15: keys +8 @ 2dp ; 16-23: keys 0..7 @ 2dp+1 ; 24-31: keys +8 @ 2dp+1
                uint32_t vf[4];
                const int vrow = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                const int vch = 2 * dp + (lane >> 4);
                ldmatrix_x4_trans(vf, sV[buf] + sw_off(vrow, vch));
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
    __nv_bfloat16* obase = out + static_cast<size_t>(tok0) * ldo + h * 64;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
        const int col = dt * 8 + 2 * t4;
        if (qi0 < len)
            *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(qi0) * ldo + col) = pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
        if (qi0 + 8 < len)
            *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(qi0 + 8) * ldo + col) = pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
    }
}

// Variant for short documents (len <= 256, the pointwise / pairwise regime): ONE CTA per (document, head) holds the whole
// Q, K and V of that head in shared memory (3 x S_pad x 128 B, loaded once with cp.async) and runs S_pad/16 warps, each
// owning 16 query rows and sweeping the resident keys in 64-key blocks with the same online softmax. Compared with the
// 64-query-tile kernel above: K/V are fetched once instead of once per query tile, there is no barrier inside the key
// loop, and key blocks / 8-key tiles beyond the document length are skipped.
// grid (H, n_docs), blockDim = (S_pad / 16) * 32 with S_pad = round_up(max_len, 64); dynamic smem = 3 * S_pad * 128 B.
__global__ void __launch_bounds__(512)
enc_attention_resident_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, int inner, const int* __restrict__ cu,
                              const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int ldo, int s_pad) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(128) uint8_t attn_smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(attn_smem);
    __nv_bfloat16* sK = sQ + s_pad * 64;
    __nv_bfloat16* sV = sK + s_pad * 64;
    __shared__ float sBias[kAttnBiasLen];

    const int h = blockIdx.x, doc = blockIdx.y;
    const int tok0 = cu[doc];
    const int len = cu[doc + 1] - tok0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const __nv_bfloat16* gq = qkv + static_cast<size_t>(tok0) * ld + h * 64;
    const __nv_bfloat16* gk = gq + inner;
    const __nv_bfloat16* gv = gq + 2 * inner;

    for (int i = tid; i < kAttnBiasLen; i += blockDim.x) sBias[i] = bias[h * kAttnBiasLen + i];
    const int rows_used = min(s_pad, (len + 63) & ~63);  // only the 64-row blocks that hold real tokens
    for (int idx = tid; idx < rows_used * 8; idx += blockDim.x) {
        const int r = idx >> 3, c = idx & 7;
        const bool ok = r < len;
        const size_t goff = static_cast<size_t>(ok ? r : 0) * ld + c * 8;
        cp_async16(sQ + sw_off(r, c), gq + goff, ok);
        cp_async16(sK + sw_off(r, c), gk + goff, ok);
        cp_async16(sV + sw_off(r, c), gv + goff, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (warp * 16 >= len) return;  // no barrier after this point

    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(qf[ks], sQ + sw_off(warp * 16 + (lane & 15), 2 * ks + (lane >> 4)));
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    const int qi0 = warp * 16 + g;
    const int nkb = (len + 63) / 64;

    for (int kb = 0; kb < nkb; ++kb) {
        const int kbase = kb * 64;
        const __nv_bfloat16* bK = sK + kbase * 64;
        const __nv_bfloat16* bV = sV + kbase * 64;
        const int npairs = min(4, (len - kbase + 15) / 16);  // 16-key pairs of n-tiles holding at least one real key
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                if (np < npairs) {
                    uint32_t kf[4];
                    const int krow = np * 16 + (lane & 7) + ((lane >> 4) << 3);
                    const int kch = 2 * ks + ((lane >> 3) & 1);
                    ldmatrix_x4(kf, bK + sw_off(krow, kch));
                    mma_bf16_16816(s[2 * np], qf[ks], kf[0], kf[1]);
                    mma_bf16_16816(s[2 * np + 1], qf[ks], kf[2], kf[3]);
                }
            }
        }
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = kbase + nt * 8 + 2 * t4 + (e & 1);
                const int qi = qi0 + ((e >> 1) << 3);
                int rel = j - qi;
                rel = max(-kAttnRelClamp, min(kAttnRelClamp, rel));
                float v = s[nt][e] + sBias[rel + kAttnRelClamp];
                v = (j < len) ? v : -INFINITY;
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
        uint32_t pf[4][4];
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
        for (int dt = 0; dt < 8; ++dt) {
            o[dt][0] *= scale[0];
            o[dt][1] *= scale[0];
            o[dt][2] *= scale[1];
            o[dt][3] *= scale[1];
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            if (ks < npairs) {  // P is exactly zero for the skipped 16-key groups
#pragma unroll
                for (int dp = 0; dp < 4; ++dp) {
                    uint32_t vf[4];
                    const int vrow = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                    const int vch = 2 * dp + (lane >> 4);
                    ldmatrix_x4_trans(vf, bV + sw_off(vrow, vch));
                    mma_bf16_16816(o[2 * dp], pf[ks], vf[0], vf[1]);
                    mma_bf16_16816(o[2 * dp + 1], pf[ks], vf[2], vf[3]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
    __nv_bfloat16* obase = out + static_cast<size_t>(tok0) * ldo + h * 64;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
        const int col = dt * 8 + 2 * t4;
        if (qi0 < len)
            *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(qi0) * ldo + col) = pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
        if (qi0 + 8 < len)
            *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(qi0 + 8) * ldo + col) = pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
    }
}

// Register-resident variant for short documents (len <= 64*NKB, NKB = 3 or 4): the 16 x len score strip of every warp stays in
// registers, so the softmax is exact and single-pass (no online rescaling of O, one max/sum reduction per row), the relative
// bias comes from an un-clamped 511-entry window table with compile-time offsets (one LDS per score, no index arithmetic), and
// exp() is one FFMA + one MUFU.EX2. Same tiling as enc_attention_kernel (64 query rows per CTA, 4 warps), but Q, all of K and
// all of V of the (document, head) are fetched once with cp.async and the kernel has only two barriers.
// ncu on the tiled kernel (profiles/r01_ncu_summary_v4.txt) showed it issue-bound (48 M warp instructions per launch, tensor
// pipe 34 % active); this variant needs ~2.5x fewer instructions per score.
// grid (ceil(maxlen/64), H, n_docs), 128 threads, dynamic smem = (64 + 2*64*NKB) * 128 B + 511 * 4 B.
constexpr int kAttnWideBias = 511;  // rel = j - i in [-255, 255] -> index rel + 255

template <int NKB>
__global__ void __launch_bounds__(128, 3)
enc_attention_regs_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, int inner, const int* __restrict__ cu,
                          const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int ldo) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(128) uint8_t attn_regs_smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(attn_regs_smem);
    __nv_bfloat16* sK = sQ + 64 * 64;
    __nv_bfloat16* sV = sK + 64 * NKB * 64;
    float* sBias = reinterpret_cast<float*>(sV + 64 * NKB * 64);

    const int qt = blockIdx.x, h = blockIdx.y, doc = blockIdx.z;
    const int tok0 = cu[doc];
    const int len = cu[doc + 1] - tok0;
    const int q0 = qt * 64;
    if (q0 >= len) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const __nv_bfloat16* gq = qkv + static_cast<size_t>(tok0) * ld + h * 64;
    const __nv_bfloat16* gk = gq + inner;
    const __nv_bfloat16* gv = gq + 2 * inner;
    const int nkb = (len + 63) >> 6;

    load_tile_64x64(sQ, gq, ld, q0, len, tid);
    for (int b = 0; b < nkb; ++b) load_tile_64x64(sK + b * 4096, gk, ld, b * 64, len, tid);
    cp_async_commit();
    for (int b = 0; b < nkb; ++b) load_tile_64x64(sV + b * 4096, gv, ld, b * 64, len, tid);
    cp_async_commit();
    for (int i = tid; i < kAttnWideBias; i += 128) {
        const int rel = max(-kAttnRelClamp, min(kAttnRelClamp, i - 255));
        sBias[i] = bias[h * kAttnBiasLen + rel + kAttnRelClamp] * 1.4426950408889634f;  // pre-scaled by log2(e)
    }
    cp_async_wait<1>();
    __syncthreads();

    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(qf[ks], sQ + sw_off(warp * 16 + (lane & 15), 2 * ks + (lane >> 4)));

    // ---- phase 1: S = Q K^T for the whole key range, in registers
    float s[NKB * 8][4];
#pragma unroll
    for (int i = 0; i < NKB * 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb) {
        if (kb < nkb) {
            const __nv_bfloat16* bK = sK + kb * 4096;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    if (kb * 64 + np * 16 < len) {
                        uint32_t kf[4];
                        const int krow = np * 16 + (lane & 7) + ((lane >> 4) << 3);
                        const int kch = 2 * ks + ((lane >> 3) & 1);
                        ldmatrix_x4(kf, bK + sw_off(krow, kch));
                        mma_bf16_16816(s[kb * 8 + 2 * np], qf[ks], kf[0], kf[1]);
                        mma_bf16_16816(s[kb * 8 + 2 * np + 1], qf[ks], kf[2], kf[3]);
                    }
                }
            }
        }
    }

    // ---- softmax (exact, single pass): v = s*log2e + bias*log2e; keys >= len -> -inf
    const float kLog2e = 1.4426950408889634f;
    const float* brow0 = sBias + (2 * t4 - (q0 + warp * 16 + g) + 255);   // row g      : + j_const
    const float* brow1 = brow0 - 8;                                          // row g + 8
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NKB * 8; ++nt) {
        const int jc = nt * 8;  // key index of element 0 of this tile, minus 2*t4 (folded into brow)
        if (jc < len) {
            float v0 = fmaf(s[nt][0], kLog2e, brow0[jc]);
            float v1 = fmaf(s[nt][1], kLog2e, brow0[jc + 1]);
            float v2 = fmaf(s[nt][2], kLog2e, brow1[jc]);
            float v3 = fmaf(s[nt][3], kLog2e, brow1[jc + 1]);
            if (jc + 8 > len) {  // tile straddles the document end (warp-uniform test, rare)
                const int j0 = jc + 2 * t4;
                if (j0 >= len) { v0 = -INFINITY; v2 = -INFINITY; }
                if (j0 + 1 >= len) { v1 = -INFINITY; v3 = -INFINITY; }
            }
            s[nt][0] = v0; s[nt][1] = v1; s[nt][2] = v2; s[nt][3] = v3;
            mx0 = fmaxf(mx0, fmaxf(v0, v1));
            mx1 = fmaxf(mx1, fmaxf(v2, v3));
        }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float l0 = 0.f, l1 = 0.f;
    uint32_t pf[NKB * 4][4];
#pragma unroll
    for (int nt = 0; nt < NKB * 8; ++nt) {
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
        if (nt * 8 < len) {
            p0 = ex2_approx(s[nt][0] - mx0);
            p1 = ex2_approx(s[nt][1] - mx0);
            p2 = ex2_approx(s[nt][2] - mx1);
            p3 = ex2_approx(s[nt][3] - mx1);
            l0 += p0 + p1;
            l1 += p2 + p3;
        }
        pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
        pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }

    // ---- phase 2: O = P V
    cp_async_wait<0>();
    __syncthreads();
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < NKB * 4; ++ks) {
        if (ks * 16 < len) {
            const __nv_bfloat16* bV = sV + (ks >> 2) * 4096;
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
                uint32_t vf[4];
                const int vrow = (ks & 3) * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                const int vch = 2 * dp + (lane >> 4);
                ldmatrix_x4_trans(vf, bV + sw_off(vrow, vch));
                mma_bf16_16816(o[2 * dp], pf[ks], vf[0], vf[1]);
                mma_bf16_16816(o[2 * dp + 1], pf[ks], vf[2], vf[3]);
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    const int qi0 = q0 + warp * 16 + g;
    __nv_bfloat16* obase = out + static_cast<size_t>(tok0) * ldo + h * 64;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
        const int col = dt * 8 + 2 * t4;
        if (qi0 < len)
            *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(qi0) * ldo + col) = pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
        if (qi0 + 8 < len)
            *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(qi0 + 8) * ldo + col) = pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
    }
}

}  // namespace b200
