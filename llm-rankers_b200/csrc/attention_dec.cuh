// Tensor-core attention for the decoder side when there are many decoder positions (qlm: T = 33 labels; any prefix > 4):
// decoder self-attention (causal, unidirectional relative-position bias, modeling_t5.py:236-251, 308-334) and
// cross-attention over the packed encoder positions (no position bias :313-315, no padded keys in the packed layout).
// Same machinery as enc_attention_kernel (mma.sync m16n8k16, 64-query tiles, 64-key blocks double-buffered with cp.async,
// online softmax in fp32, no 1/sqrt(d) scaling), with queries and keys coming from different places:
//   queries : rows doc*T + t of `q` (leading dimension ldq), head h at columns h*64
//   keys    : cu == nullptr -> rows doc*T + j of `kv` (self-attention: kv = the fused qkv buffer, k_off = inner, v_off = 2*inner)
//             cu != nullptr -> rows cu[doc] + j of `kv` (cross-attention: the stacked cross-K|V buffer, k_off / v_off of the layer)
// It replaces the CUDA-core cross_attention_kernel<40,64> / dec_self_attention_kernel on that path: 4.7 + 1.3 ms of a 19.7 ms
// qlm step (100 documents, S = 144, T = 33; profiles/r01_bench_qlm_v1.json) were spent in those two.
#pragma once
#include "attention_enc.cuh"

namespace b200 {

template <bool CAUSAL>
__global__ void __launch_bounds__(128, 4)
dec_attention_mma_kernel(const __nv_bfloat16* __restrict__ q, int ldq, int T, const __nv_bfloat16* __restrict__ kv, size_t ldkv,
                         int k_off, int v_off, const int* __restrict__ cu, const float* __restrict__ bias, int bias_len,
                         __nv_bfloat16* __restrict__ out, int ldo) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(128) __nv_bfloat16 sQ[64 * 64];
    __shared__ __align__(128) __nv_bfloat16 sK[2][64 * 64];
    __shared__ __align__(128) __nv_bfloat16 sV[2][64 * 64];
    __shared__ float sBias[CAUSAL ? kAttnRelClamp + 1 : 1];

    const int qt = blockIdx.x, h = blockIdx.y, doc = blockIdx.z;
    const int q0 = qt * 64;
    if (q0 >= T) return;
    const int key_row0 = cu ? cu[doc] : doc * T;
    const int kv_len = cu ? cu[doc + 1] - key_row0 : T;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;

    const __nv_bfloat16* gq = q + static_cast<size_t>(doc) * T * ldq + h * 64;
    const __nv_bfloat16* gk = kv + static_cast<size_t>(key_row0) * ldkv + k_off + h * 64;
    const __nv_bfloat16* gv = kv + static_cast<size_t>(key_row0) * ldkv + v_off + h * 64;
    if (CAUSAL) {
        for (int i = tid; i <= kAttnRelClamp; i += 128) sBias[i] = bias[h * bias_len + min(i, bias_len - 1)];
    }
    // causal: keys beyond the last query of this tile are never attended
    const int k_end = CAUSAL ? min(kv_len, q0 + 64) : kv_len;
    const int nkb = (k_end + 63) / 64;
    load_tile_64x64(sQ, gq, ldq, q0, T, tid);
    load_tile_64x64(sK[0], gk, ldkv, 0, kv_len, tid);
    load_tile_64x64(sV[0], gv, ldkv, 0, kv_len, tid);
    cp_async_commit();

    uint32_t qf[4][4];
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    const int qi0 = q0 + warp * 16 + g;  // decoder position of accumulator rows c0/c1; +8 for c2/c3

    for (int kb = 0; kb < nkb; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nkb) {
            load_tile_64x64(sK[buf ^ 1], gk, ldkv, (kb + 1) * 64, kv_len, tid);
            load_tile_64x64(sV[buf ^ 1], gv, ldkv, (kb + 1) * 64, kv_len, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (kb == 0) {
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
                uint32_t kf[4];
                const int krow = np * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int kch = 2 * ks + ((lane >> 3) & 1);
                ldmatrix_x4(kf, sK[buf] + sw_off(krow, kch));
                mma_bf16_16816(s[2 * np], qf[ks], kf[0], kf[1]);
                mma_bf16_16816(s[2 * np + 1], qf[ks], kf[2], kf[3]);
            }
        }
        // ---- bias / masks, online softmax. Key 0 is visible to every query (causal: j = 0 <= i; cross: kv_len >= 1), so the
        // running maximum is finite after the first block and a fully masked later block contributes exp(-inf) = 0.
        const int kbase = kb * 64;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = kbase + nt * 8 + 2 * t4 + (e & 1);
                const int qi = qi0 + ((e >> 1) << 3);
                float v = s[nt][e];
                bool ok = j < kv_len;
                if (CAUSAL) {
                    ok = ok && j <= qi;
                    v += sBias[min(max(qi - j, 0), kAttnRelClamp)];
                }
                v = ok ? v : -INFINITY;
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
        // ---- O += P V
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
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
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
    __nv_bfloat16* obase = out + static_cast<size_t>(doc) * T * ldo + h * 64;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
        const int col = dt * 8 + 2 * t4;
        if (qi0 < T)
            *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(qi0) * ldo + col) = pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
        if (qi0 + 8 < T)
            *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(qi0 + 8) * ldo + col) = pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
    }
}

}  // namespace b200
