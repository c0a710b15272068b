// Stand-alone micro-benchmark (NOT part of libb200rank.so, not built by __graft_entry__.build()): what bounds the fp32 residual
// epilogue of the O-proj / FFN-out GEMMs (gemm_tcgen05.cuh, EPI_RESID_F32)?
//
// ncu shows the O-proj GEMM (M 18400, N 1024, K 1024) at 52 % tensor-active: its 128 x 256 fp32 tile leaves through eight
// 128 x 32 staging tiles + cp.reduce.async.bulk.tensor .add (x += acc performed in L2), 75 MB of reductions per launch in 37-45 us.
// This probe replays ONLY that epilogue traffic from one persistent 128-thread "epilogue warpgroup" per SM, in the GEMM's tile order,
// and times alternatives, so that one 30-second GPU call answers whether the path is bound by L2 reduction throughput, by the
// per-chunk latency of the 2-deep staging ring, or by neither:
//   mode 0  reduce-add, 2 staging buffers           (what ships)
//   mode 1  plain TMA store, 2 staging buffers      (same bytes, no read-modify-write in L2: the reduction's own cost)
//   mode 2  reduce-add, NBUF = 4 staging buffers    (deeper ring: latency- or throughput-bound?)
//   mode 3  reduce-add, NBUF = 8
//   mode 4  TMA load x tile -> add in registers -> TMA store (the epilogue a norm-folding design needs), NBUF 4
//   mode 5  as 4 plus a bf16 copy of the tile through a second staging ring (bf16(x_new) for the next GEMM's A operand)
// Each mode runs `reps` launches back to back over a [M, N] fp32 matrix; prints us per launch and GB/s of epilogue payload.
//
//   nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -o /tmp/epi_probe experiments/epi_probe.cu && /tmp/epi_probe
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../llm-rankers_b200/csrc/ptx.cuh"

using namespace b200;

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); }     \
    } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(PFN_encodeTiled enc, void* ptr, uint64_t rows, uint64_t cols, bool f32) {
    CUtensorMap m;
    const size_t esz = f32 ? 4 : 2;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * esz};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
    return m;
}

constexpr int kTileBytes = 128 * 128;  // 128 rows x 128 B (32 fp32 or 64 bf16 columns), 128B swizzle

// MODE: 0 reduce, 1 store, 4 load+add+store, 5 = 4 + bf16 copy.  NBUF staging tiles in the ring.
template <int MODE, int NBUF>
__global__ void __launch_bounds__(128, 1)
epi_probe_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_h, int M, int N, int block_n) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage = smem;                                   // [NBUF][kTileBytes] fp32 staging (store / reduce source)
    uint8_t* ldbuf = stage + NBUF * kTileBytes;              // [NBUF][kTileBytes] x_old tiles (modes 4, 5)
    uint8_t* hbuf = ldbuf + (MODE >= 4 ? NBUF : 0) * kTileBytes;  // [4][kTileBytes] bf16 staging (mode 5): reused every 8 chunks > NBUF groups
    uint64_t* ld_bar = reinterpret_cast<uint64_t*>(hbuf + (MODE == 5 ? 4 : 0) * kTileBytes);   // [NBUF]
    const int tid = threadIdx.x;
    const bool issuer = tid == 0;
    if (issuer) {
        tma_prefetch_desc(&tmap_x);
        for (int b = 0; b < NBUF; ++b) mbar_init(&ld_bar[b], 1);
        fence_barrier_init();
    }
    __syncthreads();
    const int tiles_m = (M + 127) / 128, tiles_n = N / block_n, num_tiles = tiles_m * tiles_n;
    const int chunks = block_n / 32;                         // 32-column fp32 chunks per tile, as the GEMM epilogue walks them
    // total chunk sequence of this CTA: tile (blockIdx.x + it * gridDim.x), chunk c
    const int my_tiles = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const long total = (long)my_tiles * chunks;
    auto coords = [&](long q, int& m0, int& c0) {
        const int tile = blockIdx.x + (int)(q / chunks) * gridDim.x;
        m0 = (tile / tiles_n) * 128;
        c0 = (tile % tiles_n) * block_n + (int)(q % chunks) * 32;
    };
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 1e-3f * (float)(tid + j);   // stands for the tcgen05.ld of the accumulator chunk
    if (MODE >= 4 && issuer) {                                // prefetch the first NBUF x_old tiles
        for (long q = 0; q < NBUF && q < total; ++q) {
            int m0, c0; coords(q, m0, c0);
            mbar_arrive_expect_tx(&ld_bar[q], kTileBytes);
            tma_load_2d(ldbuf + q * kTileBytes, &tmap_x, &ld_bar[q], c0, m0, kEvictNormal);
        }
    }
    int hsel = 0;
    for (long q = 0; q < total; ++q) {
        const int b = (int)(q % NBUF);
        int m0, c0; coords(q, m0, c0);
        // the bulk op that last read staging buffer b was committed NBUF groups ago
        if (issuer) tma_store_wait_read<NBUF - 1>();
        named_bar_sync(1, 128);
        uint8_t* my_row = stage + b * kTileBytes + tid * 128;
        if (MODE >= 4) {
            mbar_wait(&ld_bar[b], (uint32_t)((q / NBUF) & 1));
            const uint8_t* xr = ldbuf + b * kTileBytes + tid * 128;
            uint32_t hp[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 x = *reinterpret_cast<const float4*>(xr + ((j ^ (tid & 7)) << 4));
                const float v0 = x.x + acc[4 * j], v1 = x.y + acc[4 * j + 1], v2 = x.z + acc[4 * j + 2], v3 = x.w + acc[4 * j + 3];
                st_shared_v4(my_row + ((j ^ (tid & 7)) << 4), __float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2), __float_as_uint(v3));
                hp[2 * j] = pack_bf16(v0, v1);
                hp[2 * j + 1] = pack_bf16(v2, v3);
            }
            if (MODE == 5) {
                // 32 bf16 columns = 64 B = half a staging row: two consecutive chunks fill one 128 x 64 bf16 tile
                uint8_t* hr = hbuf + hsel * kTileBytes + tid * 128;
                const int half = (int)(q & 1);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    st_shared_v4(hr + (((4 * half + j) ^ (tid & 7)) << 4), hp[4 * j], hp[4 * j + 1], hp[4 * j + 2], hp[4 * j + 3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                st_shared_v4(my_row + ((j ^ (tid & 7)) << 4), __float_as_uint(acc[4 * j]), __float_as_uint(acc[4 * j + 1]),
                             __float_as_uint(acc[4 * j + 2]), __float_as_uint(acc[4 * j + 3]));
        }
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (issuer) {
            if (MODE == 0) tma_reduce_add_2d(&tmap_x, stage + b * kTileBytes, c0, m0);
            else tma_store_2d(&tmap_x, stage + b * kTileBytes, c0, m0);
            if (MODE == 5 && (q & 1)) {
                tma_store_2d(&tmap_h, hbuf + hsel * kTileBytes, c0 - 32, m0);
            }
            tma_store_commit();
            if (MODE >= 4 && q + NBUF < total) {          // refill this load buffer for chunk q + NBUF (all threads are past their reads)
                int m1, c1; coords(q + NBUF, m1, c1);
                mbar_arrive_expect_tx(&ld_bar[b], kTileBytes);
                tma_load_2d(ldbuf + b * kTileBytes, &tmap_x, &ld_bar[b], c1, m1, kEvictNormal);
            }
        }
        if (MODE == 5 && (q & 1)) hsel = (hsel + 1) & 3;   // the bf16 stores share the commit groups of the fp32 ring; 4 tiles = 8 chunks of reuse distance
    }
    if (issuer) tma_store_wait_read<0>();
    __syncthreads();
}

template <int MODE, int NBUF>
static void run(const char* name, PFN_encodeTiled enc, float* x, void* h, int M, int N, int block_n, int grid, int reps) {
    CUtensorMap tx = make_map(enc, x, M, N, true);
    CUtensorMap th = make_map(enc, h, M, N, false);
    const size_t smem = (size_t)(NBUF + (MODE >= 4 ? NBUF : 0) + (MODE == 5 ? 4 : 0)) * kTileBytes + NBUF * 8 + 1024 + 64;
    CK(cudaFuncSetAttribute(epi_probe_kernel<MODE, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) epi_probe_kernel<MODE, NBUF><<<grid, 128, smem>>>(tx, th, M, N, block_n);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) epi_probe_kernel<MODE, NBUF><<<grid, 128, smem>>>(tx, th, M, N, block_n);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / reps;
    const double bytes = (double)M * N * 4.0;
    printf("%-58s %8.1f us/launch  %7.1f GB/s of fp32 tile payload\n", name, us, bytes / us * 1e-3);
}

int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 18400, N = argc > 2 ? atoi(argv[2]) : 1024, reps = argc > 3 ? atoi(argv[3]) : 50;
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(fn);
    float* x = nullptr;
    void* h = nullptr;
    const size_t rows = ((size_t)M + 127) / 128 * 128;
    CK(cudaMalloc(&x, rows * N * 4));
    CK(cudaMalloc(&h, rows * N * 2));
    CK(cudaMemset(x, 0, rows * N * 4));
    printf("epilogue probe: M %d N %d, %d SMs, one 128-thread CTA per SM, tile order of the GEMM (n fastest, round-robin), %d launches each\n", M, N, sms, reps);
    run<0, 2>("0 reduce-add, 2 staging buffers (ships)", enc, x, h, M, N, 256, sms, reps);
    run<1, 2>("1 plain store, 2 staging buffers", enc, x, h, M, N, 256, sms, reps);
    run<0, 4>("2 reduce-add, 4 staging buffers", enc, x, h, M, N, 256, sms, reps);
    run<0, 8>("3 reduce-add, 8 staging buffers", enc, x, h, M, N, 256, sms, reps);
    run<1, 8>("  plain store, 8 staging buffers", enc, x, h, M, N, 256, sms, reps);
    run<4, 4>("4 TMA load + register add + TMA store, 4+4 buffers", enc, x, h, M, N, 256, sms, reps);
    run<5, 4>("5 as 4 plus bf16(x_new) tile store", enc, x, h, M, N, 256, sms, reps);
    CK(cudaFree(x)); CK(cudaFree(h));
    return 0;
}
