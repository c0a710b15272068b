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


}  // namespace b200
