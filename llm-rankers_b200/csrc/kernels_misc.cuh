// HBM-bound and small kernels around the GEMMs: embedding gather, T5 RMSNorm,
// decoder self/cross attention for short decoder prefixes, restricted lm_head,
// score extraction. All fp32 statistics, bf16 GEMM operands, fp32 residual stream.
#pragma once
#include <cfloat>
#include "ptx.cuh"

namespace b200 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// x[t, :] = E[ids[t], :]   (fp32 table, fp32 residual stream; 16 B vector loads/stores)
// modeling_t5.py:682 (shared embedding, no sqrt(d) scaling)
__global__ void embed_kernel(const int* __restrict__ ids, const float* __restrict__ table, float* __restrict__ x,
                             int n_tokens, int d, int vocab) {
    pdl_trigger();
    pdl_wait();
    const int warps_per_block = blockDim.x >> 5;
    const int t = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (t >= n_tokens) return;
    int id = ids[t];
    if (id < 0 || id >= vocab) id = 0;
    const float4* src = reinterpret_cast<const float4*>(table + static_cast<size_t>(id) * d);
    float4* dst = reinterpret_cast<float4*>(x + static_cast<size_t>(t) * d);
    for (int i = threadIdx.x & 31; i < d / 4; i += 32) dst[i] = __ldg(src + i);
}

// T5LayerNorm (modeling_t5.py:55-68): h = bf16( x * rsqrt(mean(x^2) + eps) * w ), fp32 statistics.
// One warp per row; the row is read once (kept in registers for d <= 4096).
template <int MAX_VEC>  // MAX_VEC float4 per lane: d <= 128 * MAX_VEC
__global__ void rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, __nv_bfloat16* __restrict__ h,
                               int n_rows, int d, float eps, int reverse) {
    pdl_trigger();
    pdl_wait();
    const int warps_per_block = blockDim.x >> 5;
    int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    // reverse = 1: walk the rows from the end. The residual GEMM that precedes this kernel finished its LAST row blocks most
    // recently, so those rows are still in the 126 MB L2; and the GEMM that follows starts at row 0, which this kernel then
    // writes last (L2 ping-pong between memory-bound and compute-bound kernels).
    if (reverse) row = n_rows - 1 - row;
    const int lane = threadIdx.x & 31;
    const float4* src = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * d);
    const int nvec = d / 4;
    float4 v[MAX_VEC];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            v[i] = src[idx];
            ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
        }
    }
    ss = warp_sum(ss);
    const float r = rsqrtf(ss / static_cast<float>(d) + eps);
    const float4* wv = reinterpret_cast<const float4*>(w);
    uint2* dst = reinterpret_cast<uint2*>(h + static_cast<size_t>(row) * d);
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            const float4 g = __ldg(wv + idx);
            uint2 o;
            o.x = pack_bf16(v[i].x * r * g.x, v[i].y * r * g.y);
            o.y = pack_bf16(v[i].z * r * g.z, v[i].w * r * g.w);
            dst[idx] = o;
        }
    }
}

// Embedding gather fused with the first T5LayerNorm of the stack (SURVEY K1 + K2 of block 0): one warp per token reads the fp32 table
// row once, writes the residual stream x and h = bf16(x * rsqrt(mean(x^2) + eps) * w). Same per-lane summation order as
// rmsnorm_kernel, so the pair (embed_kernel, rmsnorm_kernel) and this kernel produce identical bits.
template <int MAX_VEC>
__global__ void embed_norm_kernel(const int* __restrict__ ids, const float* __restrict__ table, float* __restrict__ x,
                                  const float* __restrict__ w, __nv_bfloat16* __restrict__ h, int n_tokens, int d, int vocab, float eps) {
    pdl_trigger();
    pdl_wait();
    const int warps_per_block = blockDim.x >> 5;
    const int t = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (t >= n_tokens) return;
    const int lane = threadIdx.x & 31;
    int id = ids[t];
    if (id < 0 || id >= vocab) id = 0;
    const float4* src = reinterpret_cast<const float4*>(table + static_cast<size_t>(id) * d);
    float4* dx = reinterpret_cast<float4*>(x + static_cast<size_t>(t) * d);
    const int nvec = d / 4;
    float4 v[MAX_VEC];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            v[i] = __ldg(src + idx);
            dx[idx] = v[i];
            ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
        }
    }
    ss = warp_sum(ss);
    const float r = rsqrtf(ss / static_cast<float>(d) + eps);
    const float4* wv = reinterpret_cast<const float4*>(w);
    uint2* dst = reinterpret_cast<uint2*>(h + static_cast<size_t>(t) * d);
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            const float4 g = __ldg(wv + idx);
            uint2 o;
            o.x = pack_bf16(v[i].x * r * g.x, v[i].y * r * g.y);
            o.y = pack_bf16(v[i].z * r * g.z, v[i].w * r * g.w);
            dst[idx] = o;
        }
    }
}

// Decoder self-attention for a short prefix (<= 64 positions), causal, with the unidirectional
// relative-position bias (modeling_t5.py:236-251, 308-334; no 1/sqrt(d) scaling).
// grid (H, n_docs); rows are [q | k | v], each `inner` wide, head dim 64; row of (doc, position p) = doc * doc_rows + p.
// The Tq queries of a document sit at positions q_pos0 .. q_pos0 + Tq - 1 and attend to the keys at positions 0 .. own position:
//   whole prefix (qlm / likelihood / first greedy step):  doc_rows = Tq = T, q_pos0 = 0, rows in the qkv workspace;
//   KV-cached greedy step (generation/utils.py:2762-2804): doc_rows = row stride of the cache, Tq = 1, q_pos0 = number of cached
//   positions — the rows of the earlier positions were written by the earlier steps' projections (engine.cu, run_decoder_cached_step).
// A query's arithmetic is the same whichever way it is reached, so the cached step is bit-identical to re-running the prefix.
// Output rows are dense: doc * Tq + local query index. bias: [H][bias_len] indexed by (i - j) clamped to bias_len-1.
// kv_out != nullptr (first step of a cached greedy call): the k | v values read are also copied to the cache, whose rows have the same
// [q | k | v] layout and kv_out_doc_rows rows per document.
__global__ void dec_self_attention_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, int inner, int doc_rows, int q_pos0, int Tq,
                                          const float* __restrict__ bias, int bias_len,
                                          __nv_bfloat16* __restrict__ out, int ldo,
                                          __nv_bfloat16* __restrict__ kv_out = nullptr, int kv_out_doc_rows = 0) {
    pdl_trigger();
    pdl_wait();
    constexpr int D = 64;
    constexpr int MAXT = 64;
    __shared__ float sK[MAXT][D + 1];
    __shared__ float sV[MAXT][D + 1];
    __shared__ float sQ[4][D];
    __shared__ float sP[4][MAXT];
    const int h = blockIdx.x, doc = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = q_pos0 + Tq;   // keys in play
    const __nv_bfloat16* base = qkv + static_cast<size_t>(doc) * doc_rows * ld + h * D;
    __nv_bfloat16* cbase = kv_out ? kv_out + static_cast<size_t>(doc) * kv_out_doc_rows * ld + h * D : nullptr;
    for (int idx = threadIdx.x; idx < T * D; idx += blockDim.x) {
        const int j = idx / D, dd = idx % D;
        const __nv_bfloat16 kb = base[static_cast<size_t>(j) * ld + inner + dd], vb = base[static_cast<size_t>(j) * ld + 2 * inner + dd];
        sK[j][dd] = __bfloat162float(kb);
        sV[j][dd] = __bfloat162float(vb);
        if (cbase) {
            cbase[static_cast<size_t>(j) * ld + inner + dd] = kb;
            cbase[static_cast<size_t>(j) * ld + 2 * inner + dd] = vb;
        }
    }
    __syncthreads();
    const float* hb = bias + static_cast<size_t>(h) * bias_len;
    for (int i = q_pos0 + warp; i < T; i += 4) {
        sQ[warp][lane] = __bfloat162float(base[static_cast<size_t>(i) * ld + lane]);
        sQ[warp][lane + 32] = __bfloat162float(base[static_cast<size_t>(i) * ld + lane + 32]);
        __syncwarp();
        float s[2];
        float m = -INFINITY;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int j = lane + 32 * r;
            s[r] = -INFINITY;
            if (j <= i) {
                float acc = 0.f;
#pragma unroll 16
                for (int dd = 0; dd < D; ++dd) acc += sQ[warp][dd] * sK[j][dd];
                const int rel = min(i - j, bias_len - 1);
                s[r] = acc + hb[rel];
            }
            m = fmaxf(m, s[r]);
        }
        m = warp_max(m);
        float l = 0.f;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int j = lane + 32 * r;
            const float p = (j <= i) ? __expf(s[r] - m) : 0.f;
            sP[warp][j] = p;
            l += p;
        }
        l = warp_sum(l);
        __syncwarp();
        const float inv = 1.f / l;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int dd = lane + 32 * r;
            float acc = 0.f;
            for (int j = 0; j <= i; ++j) acc += sP[warp][j] * sV[j][dd];
            out[(static_cast<size_t>(doc) * Tq + (i - q_pos0)) * ldo + h * D + dd] = __float2bfloat16(acc * inv);
        }
        __syncwarp();
    }
}

// Cross-attention of T decoder positions over the S encoder positions of one document
// (modeling_t5.py:298-338 with key_value_states; zero position bias :313-315; the packed
// varlen layout holds no padded keys, so the key-padding mask :720-726 is implicit).
// grid (H, n_docs), 128 threads. q rows = doc*T + t ([*, inner]); K/V rows = packed encoder
// tokens cu[doc].., K at column k_off + h*64, V at v_off + h*64 of the cross-KV buffer.
// Keys are streamed in chunks of CH through shared memory with an online softmax per query
// (<MAXT=4, CH=128> for yes_no / generation prefixes, <40, 64> for qlm; both fit 48 KB static smem).
// Key split (flash-decoding): with gridDim.z = nsplit > 1 every CTA owns the key chunks c0 = (blockIdx.z + i*nsplit)*CH and writes
// its unnormalised partial result (m, l, o[64]) per query to `partial` ([n_docs][H][nsplit][T][66] fp32); cross_attention_combine_kernel
// merges the splits. One (document, head) pair otherwise streams all S keys through ONE CTA: 100 us per layer at S = 1.5 k
// (setwise compare prompts), i.e. 16 CTAs on a 148-SM GPU.
template <int MAXT, int CH>
__global__ void cross_attention_kernel(const __nv_bfloat16* __restrict__ q, int ldq, int T,
                                       const __nv_bfloat16* __restrict__ kv, size_t ldkv, int k_off, int v_off,
                                       const int* __restrict__ cu, __nv_bfloat16* __restrict__ out, int ldo,
                                       float* __restrict__ partial = nullptr) {
    pdl_trigger();
    pdl_wait();
    constexpr int D = 64;
    constexpr int LDS = D + 2;  // bf16 row stride 132 B: conflict-free for row-per-thread dots
    __shared__ __nv_bfloat16 sK[CH * LDS];
    __shared__ __nv_bfloat16 sV[CH * LDS];
    __shared__ float sQ[MAXT][D];
    __shared__ float sS[MAXT][CH];
    __shared__ float sO[MAXT][D];
    __shared__ float sM[MAXT], sL[MAXT];
    const int h = blockIdx.x, doc = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tok0 = cu[doc];
    const int S = cu[doc + 1] - tok0;
    for (int idx = tid; idx < T * D; idx += blockDim.x) {
        const int t = idx / D, dd = idx % D;
        sQ[t][dd] = __bfloat162float(q[(static_cast<size_t>(doc) * T + t) * ldq + h * D + dd]);
        sO[t][dd] = 0.f;
    }
    if (tid < T) { sM[tid] = -INFINITY; sL[tid] = 0.f; }
    __syncthreads();
    const int nsplit = gridDim.z;
    for (int c0 = blockIdx.z * CH; c0 < S; c0 += CH * nsplit) {
        const int n = min(CH, S - c0);
        // coalesced 16 B loads: 8 lanes per 128 B head row
        for (int idx = tid; idx < CH * 8; idx += blockDim.x) {
            const int j = idx >> 3, c = idx & 7;
            uint4 kk = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
            if (j < n) {
                const __nv_bfloat16* row = kv + static_cast<size_t>(tok0 + c0 + j) * ldkv + h * D + c * 8;
                kk = *reinterpret_cast<const uint4*>(row + k_off);
                vv = *reinterpret_cast<const uint4*>(row + v_off);
            }
            uint32_t* dk = reinterpret_cast<uint32_t*>(sK + j * LDS + c * 8);
            uint32_t* dv = reinterpret_cast<uint32_t*>(sV + j * LDS + c * 8);
            dk[0] = kk.x; dk[1] = kk.y; dk[2] = kk.z; dk[3] = kk.w;
            dv[0] = vv.x; dv[1] = vv.y; dv[2] = vv.z; dv[3] = vv.w;
        }
        __syncthreads();
        // scores: thread j owns key j of the chunk, loops over the T queries
        if (tid < CH) {
            const int j = tid;
            const __nv_bfloat162* krow = reinterpret_cast<const __nv_bfloat162*>(sK + j * LDS);
            for (int t = 0; t < T; ++t) {
                float acc = 0.f;
#pragma unroll
                for (int dd = 0; dd < D / 2; ++dd) {
                    const float2 kf = __bfloat1622float2(krow[dd]);
                    acc += sQ[t][2 * dd] * kf.x + sQ[t][2 * dd + 1] * kf.y;
                }
                sS[t][j] = (j < n) ? acc : -INFINITY;
            }
        }
        __syncthreads();
        // online softmax + PV: one warp per query, lanes own 2 output dims
        for (int t = warp; t < T; t += 4) {
            float m = -INFINITY;
            for (int j = lane; j < CH; j += 32) m = fmaxf(m, sS[t][j]);
            m = warp_max(m);
            const float m_old = sM[t];
            const float m_new = fmaxf(m_old, m);
            const float scale = __expf(m_old - m_new);
            float lsum = 0.f;
            for (int j = lane; j < CH; j += 32) {
                const float p = __expf(sS[t][j] - m_new);
                sS[t][j] = p;
                lsum += p;
            }
            lsum = warp_sum(lsum);
            __syncwarp();
            float o0 = sO[t][2 * lane] * scale, o1 = sO[t][2 * lane + 1] * scale;
            for (int j = 0; j < n; ++j) {
                const float p = sS[t][j];
                const float2 vf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sV + j * LDS + 2 * lane));
                o0 += p * vf.x;
                o1 += p * vf.y;
            }
            sO[t][2 * lane] = o0;
            sO[t][2 * lane + 1] = o1;
            if (lane == 0) { sM[t] = m_new; sL[t] = sL[t] * scale + lsum; }
        }
        __syncthreads();
    }
    if (nsplit > 1) {
        float* dst = partial + ((static_cast<size_t>(doc) * gridDim.x + h) * nsplit + blockIdx.z) * T * 66;
        for (int idx = tid; idx < T * 66; idx += blockDim.x) {
            const int t = idx / 66, dd = idx % 66;
            dst[idx] = dd < D ? sO[t][dd] : (dd == D ? sM[t] : sL[t]);   // a split with no keys leaves m = -inf, l = 0, o = 0
        }
        return;
    }
    for (int idx = tid; idx < T * D; idx += blockDim.x) {
        const int t = idx / D, dd = idx % D;
        out[(static_cast<size_t>(doc) * T + t) * ldo + h * D + dd] = __float2bfloat16(sO[t][dd] / sL[t]);
    }
}

// Merge of the key splits: grid (H, n_docs), 64 threads (one per head dim); exact log-sum-exp combination in fp32.
__global__ void cross_attention_combine_kernel(const float* __restrict__ partial, int nsplit, int T, __nv_bfloat16* __restrict__ out, int ldo) {
    pdl_trigger();
    pdl_wait();
    const int h = blockIdx.x, doc = blockIdx.y, dd = threadIdx.x;
    const float* src = partial + (static_cast<size_t>(doc) * gridDim.x + h) * nsplit * T * 66;
    for (int t = 0; t < T; ++t) {
        float m = -INFINITY;
        for (int sp = 0; sp < nsplit; ++sp) m = fmaxf(m, src[(sp * T + t) * 66 + 64]);
        float l = 0.f, o = 0.f;
        for (int sp = 0; sp < nsplit; ++sp) {
            const float* p = src + (sp * T + t) * 66;
            const float w = __expf(p[64] - m);   // exp(-inf - m) = 0 for an empty split
            l += p[65] * w;
            o += p[dd] * w;
        }
        out[(static_cast<size_t>(doc) * T + t) * ldo + h * 64 + dd] = __float2bfloat16(o / l);
    }
}

// Cross-attention for a single decoder position (T = 1: pointwise yes_no), the HBM-bound case: one warp per
// (document, head) streams that head's K rows (one 128 B row per lane: lane j owns keys j, j+32, ...), does the softmax
// with warp shuffles, then streams the V rows coalesced (lane owns 2 of the 64 dims). grid (H/4, n_docs), 128 threads:
// the 4 warps of a block take 4 adjacent heads, i.e. 512 contiguous bytes of every K/V row.
// MAX_ROUNDS * 32 bounds the encoder length handled by this kernel (longer documents take the chunked kernel above).
template <int MAX_ROUNDS>
__global__ void __launch_bounds__(128)
cross_attention_t1_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ kv, size_t ldkv,
                          int k_off, int v_off, const int* __restrict__ cu, __nv_bfloat16* __restrict__ out, int ldo) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sQ[4][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x * 4 + warp, doc = blockIdx.y;
    const int tok0 = cu[doc];
    const int S = cu[doc + 1] - tok0;
    {
        const __nv_bfloat162 qv = *reinterpret_cast<const __nv_bfloat162*>(q + static_cast<size_t>(doc) * ldq + h * 64 + 2 * lane);
        const float2 qf = __bfloat1622float2(qv);
        sQ[warp][2 * lane] = qf.x;
        sQ[warp][2 * lane + 1] = qf.y;
    }
    __syncwarp();
    const __nv_bfloat16* kbase = kv + static_cast<size_t>(tok0) * ldkv + k_off + h * 64;
    const __nv_bfloat16* vbase = kv + static_cast<size_t>(tok0) * ldkv + v_off + h * 64;
    float s[MAX_ROUNDS];
    float m = -INFINITY;
#pragma unroll
    for (int r = 0; r < MAX_ROUNDS; ++r) {
        const int j = r * 32 + lane;
        s[r] = -INFINITY;
        if (j < S) {
            const uint4* row = reinterpret_cast<const uint4*>(kbase + static_cast<size_t>(j) * ldkv);
            uint4 c[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) c[i] = __ldg(row + i);
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&c[i]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 kf = __bfloat1622float2(p2[e]);
                    acc += sQ[warp][8 * i + 2 * e] * kf.x + sQ[warp][8 * i + 2 * e + 1] * kf.y;
                }
            }
            s[r] = acc;
        }
        m = fmaxf(m, s[r]);
    }
    m = warp_max(m);
    float l = 0.f;
#pragma unroll
    for (int r = 0; r < MAX_ROUNDS; ++r) {
        s[r] = (r * 32 + lane < S) ? __expf(s[r] - m) : 0.f;
        l += s[r];
    }
    l = warp_sum(l);
    // V phase: 16 B per lane, 8 lanes per 128 B row, 4 rows per warp-wide load (lane>>3 picks the row, lane&7 the 8-dim chunk)
    // -> 4x the bytes in flight of a 4 B/lane row walk; partial sums over the 4 row groups are folded with two shuffles.
    const int chunk = lane & 7, rsub = lane >> 3;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
    for (int r = 0; r < MAX_ROUNDS; ++r) {
        const int base = r * 32;
        if (base < S) {  // warp-uniform
#pragma unroll
            for (int jj = 0; jj < 32; jj += 4) {
                const int j = base + jj + rsub;
                const float p = __shfl_sync(0xffffffffu, s[r], jj + rsub);  // p == 0 for j >= S
                if (j < S) {
                    const uint4 c = __ldg(reinterpret_cast<const uint4*>(vbase + static_cast<size_t>(j) * ldkv + chunk * 8));
                    const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 vf = __bfloat1622float2(p2[e]);
                        o[2 * e] += p * vf.x;
                        o[2 * e + 1] += p * vf.y;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
        o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
    }
    if (rsub == 0) {
        const float inv = 1.f / l;
        uint4 v;
        v.x = pack_bf16(o[0] * inv, o[1] * inv);
        v.y = pack_bf16(o[2] * inv, o[3] * inv);
        v.z = pack_bf16(o[4] * inv, o[5] * inv);
        v.w = pack_bf16(o[6] * inv, o[7] * inv);
        *reinterpret_cast<uint4*>(out + static_cast<size_t>(doc) * ldo + h * 64 + chunk * 8) = v;
    }
}

// dst[c, r] = src[r, c] for a bf16 matrix (load-time helper: builds W_v^T for the fused decoder W_o.W_v product)
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, int rows, int cols, int ld_src,
                                      __nv_bfloat16* __restrict__ dst, int ld_dst) {
    pdl_trigger();
    pdl_wait();
    __shared__ __nv_bfloat16 tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = src[static_cast<size_t>(r) * ld_src + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[static_cast<size_t>(c) * ld_dst + r] = tile[threadIdx.x][i];
    }
}

// logits[r, c] = h[row_of(r), :] . lm_head[cols[c], :]  for a short list of vocabulary ids
// (modeling_t5.py:1110 restricted to the columns pointwise.py:120-121 / setwise.py:186 read).
// grid = n_rows, one warp per column (strided). row_of(r) = r * row_stride + row_offset.
__global__ void lm_head_cols_kernel(const __nv_bfloat16* __restrict__ h, int d, int row_stride, int row_offset,
                                    const __nv_bfloat16* __restrict__ lm_head, const int* __restrict__ cols, int ncols,
                                    float scale, float* __restrict__ logits) {
    pdl_trigger();
    pdl_wait();
    const int r = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(h + static_cast<size_t>(r * row_stride + row_offset) * d);
    for (int c = warp; c < ncols; c += nwarps) {
        const __nv_bfloat162* wv = reinterpret_cast<const __nv_bfloat162*>(lm_head + static_cast<size_t>(cols[c]) * d);
        float acc = 0.f;
        for (int i = lane; i < d / 2; i += 32) {
            const float2 a = __bfloat1622float2(hv[i]);
            const float2 b = __bfloat1622float2(wv[i]);
            acc += a.x * b.x + a.y * b.y;
        }
        acc = warp_sum(acc);
        if (lane == 0) logits[static_cast<size_t>(r) * ncols + c] = acc * scale;
    }
}

// pointwise.py:120-124: softmax over (yes, no) -> P(yes); fp32 like the reference's CPU path.
__global__ void yes_no_score_kernel(const float* __restrict__ logits2, float* __restrict__ score, int n) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float y = logits2[2 * i], no = logits2[2 * i + 1];
    const float m = fmaxf(y, no);
    const float ey = expf(y - m), en = expf(no - m);
    score[i] = ey / (ey + en);
}

// Gather rows r*row_stride + row_offset of a bf16 matrix into a dense [n, d] matrix
// (last decoder position of every document, for the full-vocabulary lm_head GEMM).
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, int d, int row_stride, int row_offset,
                                   __nv_bfloat16* __restrict__ dst, int n) {
    pdl_trigger();
    pdl_wait();
    const int r = blockIdx.x;
    if (r >= n) return;
    const uint4* s = reinterpret_cast<const uint4*>(src + static_cast<size_t>(r * row_stride + row_offset) * d);
    uint4* o = reinterpret_cast<uint4*>(dst + static_cast<size_t>(r) * d);
    for (int i = threadIdx.x; i < d / 8; i += blockDim.x) o[i] = s[i];
}

// Per-row reductions over a full-vocabulary fp32 logits row. One block per row.
// Logits are multiplied by `scale` on read (d_model^-0.5 when embeddings are tied, else 1).
//   mode 0: out_f[r] = logit[label[r]] - logsumexp(row)      (qlm: pointwise.py:77-79, -CE per position)
//   mode 1: out_i[r] = argmax(row) (first index on ties, like torch.argmax)
//   mode 2: out_f[r*ncols + c] = softmax(row)[cols[c]]        (setwise.py:184-186 likelihood scoring)
__global__ void vocab_row_kernel(const float* __restrict__ logits, int V, size_t ld, int mode, float scale,
                                 const int* __restrict__ labels, const int* __restrict__ cols, int ncols,
                                 float* __restrict__ out_f, int* __restrict__ out_i) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    const int r = blockIdx.x;
    const float* row = logits + static_cast<size_t>(r) * ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    float m = -INFINITY;
    int mi = 0x7fffffff;
    for (int i = tid; i < V; i += blockDim.x) {
        const float v = row[i] * scale;
        if (v > m) { m = v; mi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
    }
    if (lane == 0) { s_val[warp] = m; s_idx[warp] = mi; }
    __syncthreads();
    if (warp == 0) {
        m = lane < nwarps ? s_val[lane] : -INFINITY;
        mi = lane < nwarps ? s_idx[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, m, o);
            const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
            if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
        }
        if (lane == 0) { s_val[0] = m; s_idx[0] = mi; }
    }
    __syncthreads();
    m = s_val[0];
    mi = s_idx[0];
    if (mode == 1) {
        if (tid == 0) out_i[r] = mi;
        return;
    }
    __syncthreads();
    float sum = 0.f;
    for (int i = tid; i < V; i += blockDim.x) sum += expf(row[i] * scale - m);
    sum = warp_sum(sum);
    if (lane == 0) s_val[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        sum = lane < nwarps ? s_val[lane] : 0.f;
        sum = warp_sum(sum);
        if (lane == 0) s_val[0] = sum;
    }
    __syncthreads();
    sum = s_val[0];
    if (mode == 0) {
        if (tid == 0) out_f[r] = row[labels[r]] * scale - m - logf(sum);
    } else {
        for (int c = tid; c < ncols; c += blockDim.x) out_f[static_cast<size_t>(r) * ncols + c] = expf(row[cols[c]] * scale - m) / sum;
    }
}

// The same reductions with the row read from memory ONCE: 1024 threads hold the scaled row in registers (MAXV float4 each, i.e.
// V <= 4096 * MAXV), so the maximum, the sum of exponentials and the gathers are one streaming pass over the 128 KB row instead of
// two or three (qlm reads 5 k such rows per device pass). Argmax ties resolve to the first index exactly as above.
template <int MAXV>
__global__ void __launch_bounds__(1024)
vocab_row_regs_kernel(const float* __restrict__ logits, int V, size_t ld, int mode, float scale,
                      const int* __restrict__ labels, const int* __restrict__ cols, int ncols,
                      float* __restrict__ out_f, int* __restrict__ out_i) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    const int r = blockIdx.x;
    const float* row = logits + static_cast<size_t>(r) * ld;
    const float4* row4 = reinterpret_cast<const float4*>(row);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int nvec = V >> 2;
    float4 v[MAXV];
    float m = -INFINITY;
    int mi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int idx = tid + i * 1024;
        if (idx < nvec) {
            float4 t = __ldcs(row4 + idx);
            t.x *= scale; t.y *= scale; t.z *= scale; t.w *= scale;
            v[i] = t;
            if (t.x > m) { m = t.x; mi = 4 * idx; }
            if (t.y > m) { m = t.y; mi = 4 * idx + 1; }
            if (t.z > m) { m = t.z; mi = 4 * idx + 2; }
            if (t.w > m) { m = t.w; mi = 4 * idx + 3; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
    }
    if (lane == 0) { s_val[warp] = m; s_idx[warp] = mi; }
    __syncthreads();
    if (warp == 0) {
        m = lane < nwarps ? s_val[lane] : -INFINITY;
        mi = lane < nwarps ? s_idx[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, m, o);
            const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
            if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
        }
        if (lane == 0) { s_val[0] = m; s_idx[0] = mi; }
    }
    __syncthreads();
    m = s_val[0];
    mi = s_idx[0];
    if (mode == 1) {
        if (tid == 0) out_i[r] = mi;
        return;
    }
    __syncthreads();
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int idx = tid + i * 1024;
        if (idx < nvec) sum += expf(v[i].x - m) + expf(v[i].y - m) + expf(v[i].z - m) + expf(v[i].w - m);
    }
    sum = warp_sum(sum);
    if (lane == 0) s_val[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        sum = lane < nwarps ? s_val[lane] : 0.f;
        sum = warp_sum(sum);
        if (lane == 0) s_val[0] = sum;
    }
    __syncthreads();
    sum = s_val[0];
    if (mode == 0) {
        if (tid == 0) out_f[r] = row[labels[r]] * scale - m - logf(sum);
    } else {
        for (int c = tid; c < ncols; c += blockDim.x) out_f[static_cast<size_t>(r) * ncols + c] = expf(row[cols[c]] * scale - m) / sum;
    }
}

// qlm: score[doc] = sum_t logprob[doc*T + t]   (pointwise.py:79)
__global__ void sum_rows_kernel(const float* __restrict__ v, int T, float* __restrict__ out, int n) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += v[static_cast<size_t>(i) * T + t];
    out[i] = s;
}

// Debug-only reference GEMM on CUDA cores (enabled with B200RANK_DEBUG_SIMT_GEMM=1) used by the
// GPU tests to validate the tcgen05 kernel and to bisect failures; never used on the product path.
__global__ void gemm_simt_debug_kernel(const __nv_bfloat16* __restrict__ a, int lda, const __nv_bfloat16* __restrict__ w,
                                       int ldw, int M, int N, int K, int epi, int block_n, void* out, int ldo) {
    const int m = blockIdx.y * blockDim.y + threadIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_out = (epi == 2) ? N / 2 : N;
    if (m >= M || c >= n_out) return;
    auto dot = [&](int wrow) {
        float acc = 0.f;
        for (int k = 0; k < K; ++k)
            acc += __bfloat162float(a[static_cast<size_t>(m) * lda + k]) * __bfloat162float(w[static_cast<size_t>(wrow) * ldw + k]);
        return acc;
    };
    if (epi == 2) {
        const int half = block_n / 2;
        const int nb = c / half, j = c % half;
        const float g = dot(nb * block_n + j), l = dot(nb * block_n + half + j);
        reinterpret_cast<__nv_bfloat16*>(out)[static_cast<size_t>(m) * ldo + c] = __float2bfloat16(gelu_new(g) * l);
    } else {
        const float acc = dot(c);
        if (epi == 0) reinterpret_cast<__nv_bfloat16*>(out)[static_cast<size_t>(m) * ldo + c] = __float2bfloat16(acc);
        else if (epi == 5) reinterpret_cast<__nv_bfloat16*>(out)[static_cast<size_t>(m) * ldo + c] = __float2bfloat16(fmaxf(acc, 0.f));
        else if (epi == 1) reinterpret_cast<float*>(out)[static_cast<size_t>(m) * ldo + c] += acc;
        else reinterpret_cast<float*>(out)[static_cast<size_t>(m) * ldo + c] = acc;
    }
}

}  // namespace b200
