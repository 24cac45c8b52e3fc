// ES-MoE dispatch, persistent TMA + tcgen05 version (BatchedExpertComputation.compute_sparse_experts_batched with 1x1-conv
// experts, moe/utils.py:119-209):   out[b] = clamp( sum_j fp16( fp16(x[b] W[e_bj]^T) * w_bj ), +-clamp ),  w <= w_min dropped.
//
// Routing is per image, so the "gather" is a row range and the "scatter" is the output tile itself: the kernel reads every
// 128-token tile of x ONCE (TMA, 4 x 16 KB swizzled k-chunks kept resident), streams the <= 2 routed experts' weight tiles
// through a 5-deep TMA ring, accumulates each expert in its own TMEM accumulator (tcgen05.mma, M=128, N=128, K=256) and
// combines them with the routing weights in the epilogue, which leaves through a swizzled staging tile and TMA stores.
// Algorithmic HBM traffic = (k+1)*d*2 bytes per token (x in, out) + the expert weights once (they live in L2).
//   warp 0: TMA producer | warp 1: MMA issuer | warps 2..9: epilogue (TMEM lane quarter = warp%4, column half = (warp-2)/4)
// TMEM: 2 (double buffer over N halves) x 2 (experts) x 128 columns = 512.
#include <cuda.h>

#include "tc_common.cuh"

namespace ym {

constexpr int DP_THREADS = 320, DP_BM = 128, DP_BN = 128, DP_KC = 64, DP_BSTAGES = 5;

struct DispatchParams {
    const int* route_idx; const float* route_w;
    int topk, HW, tiles_per_img, total_tiles, K, N;   // K = d (multiple of 64, <= 256), N = outputs (multiple of 128, <= 256)
    float w_min, clamp;
};

__device__ __forceinline__ void dp_tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void dp_tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void dp_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// live routes of image b (weights above the eval threshold), in route order
__device__ __forceinline__ int dp_routes(const DispatchParams& p, int b, int (&e)[2], float (&w)[2]) {
    int n = 0;
    for (int j = 0; j < p.topk && j < 2; ++j) {
        const float wj = p.route_w[b * p.topk + j];
        if (wj > p.w_min) { e[n] = p.route_idx[b * p.topk + j]; w[n] = wj; ++n; }
    }
    return n;
}

__global__ void __launch_bounds__(DP_THREADS, 1) tc_dispatch_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                    const __grid_constant__ CUtensorMap map_w,
                                                                    const __grid_constant__ CUtensorMap map_o, const DispatchParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    constexpr int A_CHUNK = DP_BM * 128, B_TILE = DP_BN * 128, STG_HALF = DP_BM * 128;
    const int kchunks = p.K / DP_KC;                       // <= 4
    unsigned char* sA = smem;                              // [4][A_CHUNK]       64 KB
    unsigned char* sB = sA + 4 * A_CHUNK;                  // [DP_BSTAGES][B_TILE] 80 KB
    unsigned char* stg = sB + DP_BSTAGES * B_TILE;         // [2 buffers][2 halves][STG_HALF] 64 KB
    __shared__ uint64_t a_full, a_empty, b_full[DP_BSTAGES], b_empty[DP_BSTAGES], t_full[2], t_empty[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::mbar_init(&a_full, 1);
        tc::mbar_init(&a_empty, 1);
        for (int s = 0; s < DP_BSTAGES; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&t_full[a], 1); tc::mbar_init(&t_empty[a], 8); }
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const int nhalves = p.N / DP_BN;                        // 1 or 2

    if (warp == 0) {
        if (lane == 0) {
            uint32_t titer = 0, bidx = 0;                   // titer counts PROCESSED tiles only (phase of a_full / a_empty)
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int b = tile / p.tiles_per_img, m0 = b * p.HW + (tile - b * p.tiles_per_img) * DP_BM;
                int e[2]; float w[2];
                const int nr = dp_routes(p, b, e, w);
                if (nr == 0) continue;                      // epilogue writes zeros without the tensor core
                tc::mbar_wait(&a_empty, (titer & 1) ^ 1);
                dp_expect_tx(&a_full, (uint32_t)(kchunks * A_CHUNK));
                for (int kc = 0; kc < kchunks; ++kc) dp_tma_load_2d(sA + kc * A_CHUNK, &map_x, kc * DP_KC, m0, &a_full);
                for (int nh = 0; nh < nhalves; ++nh)
                    for (int j = 0; j < nr; ++j)
                        for (int kc = 0; kc < kchunks; ++kc, ++bidx) {
                            const int s = bidx % DP_BSTAGES;
                            tc::mbar_wait(&b_empty[s], ((bidx / DP_BSTAGES) & 1) ^ 1);
                            dp_expect_tx(&b_full[s], (uint32_t)B_TILE);
                            dp_tma_load_2d(sB + s * B_TILE, &map_w, kc * DP_KC, e[j] * p.N + nh * DP_BN, &b_full[s]);
                        }
                ++titer;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_f16(DP_BM, DP_BN);
            uint32_t titer = 0, bidx = 0, it2 = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int b = tile / p.tiles_per_img;
                int e[2]; float w[2];
                const int nr = dp_routes(p, b, e, w);
                if (nr == 0) continue;
                tc::mbar_wait(&a_full, titer & 1);
                tc::fence_after_sync();
                for (int nh = 0; nh < nhalves; ++nh, ++it2) {
                    const uint32_t buf = it2 & 1;
                    tc::mbar_wait(&t_empty[buf], ((it2 >> 1) & 1) ^ 1);
                    tc::fence_after_sync();
                    for (int j = 0; j < nr; ++j) {
                        const uint32_t tacc = tmem_base + buf * 256 + j * 128;
                        for (int kc = 0; kc < kchunks; ++kc, ++bidx) {
                            const int s = bidx % DP_BSTAGES;
                            tc::mbar_wait(&b_full[s], (bidx / DP_BSTAGES) & 1);
                            tc::fence_after_sync();
                            const uint64_t adesc = tc::make_desc(smem_u32(sA + kc * A_CHUNK), 1024, 2);
                            const uint64_t bdesc = tc::make_desc(smem_u32(sB + s * B_TILE), 1024, 2);
#pragma unroll
                            for (int k = 0; k < DP_KC / 16; ++k) tc::mma_f16_ss(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) ? 1u : 0u);
                            tc::mma_commit(&b_empty[s]);
                        }
                    }
                    tc::mma_commit(&t_full[buf]);
                }
                tc::mma_commit(&a_empty);                   // all MMAs reading this tile's x chunks are done
                ++titer;
            }
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;     // 64 columns per thread
        const int r = q * 32 + lane;
        const bool elected = (warp == 2 && lane == 0);
        uint32_t it2 = 0, sidx = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int b = tile / p.tiles_per_img, mloc = (tile - b * p.tiles_per_img) * DP_BM, m0 = b * p.HW + mloc;
            int e[2]; float w[2];
            const int nr = dp_routes(p, b, e, w);
            for (int nh = 0; nh < nhalves; ++nh, ++sidx) {
                unsigned char* sb = stg + (sidx & 1) * (2 * STG_HALF) + half * STG_HALF;   // this thread's 64-column half
                uint32_t buf = 0;
                if (nr > 0) {
                    buf = it2 & 1;
                    tc::mbar_wait(&t_full[buf], (it2 >> 1) & 1);
                    tc::fence_after_sync();
                }
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                    for (int j = 0; j < nr; ++j) {
                        uint32_t rr[32];
                        tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256 + j * 128 + half * 64 + c0, rr);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i)   // expert output rounded to fp16, weighted in fp32, rounded to fp16 (utils.py:200-203)
                            v[i] += __half2float(__float2half_rn(__half2float(__float2half_rn(__uint_as_float(rr[i]))) * w[j]));
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 o;
                        uint32_t pk[4];
#pragma unroll
                        for (int h2 = 0; h2 < 4; ++h2) {
                            const float a0 = fminf(fmaxf(v[c * 8 + 2 * h2], -p.clamp), p.clamp);
                            const float a1 = fminf(fmaxf(v[c * 8 + 2 * h2 + 1], -p.clamp), p.clamp);
                            pk[h2] = pack_half2(a0, a1);
                        }
                        o = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(sb + tc::sw128_offset(r, c0 / 8 + c)) = o;
                    }
                }
                if (nr > 0) {
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&t_empty[buf]);
                    ++it2;
                }
                tc::fence_proxy_async();
                if (elected) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // previous stores released the other buffer
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
                if (elected) {
                    unsigned char* s0 = stg + (sidx & 1) * (2 * STG_HALF);
                    // rows beyond this image's HW are clipped by giving the store an image-local tensor map row bound
                    dp_tma_store_2d(&map_o, s0, nh * DP_BN, m0);
                    dp_tma_store_2d(&map_o, s0 + STG_HALF, nh * DP_BN + 64, m0);
                    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                }
                (void)mloc;
            }
        }
        if (elected) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*DpEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static DpEncodeFn dp_encode() {
    static DpEncodeFn fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<DpEncodeFn>(q);
    }
    return fn;
}
static bool dp_map2d(CUtensorMap* m, const void* base, cuuint64_t cols, cuuint64_t rows, cuuint64_t pitch_bytes, cuuint32_t box_c,
                     cuuint32_t box_r) {
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_c, box_r};
    cuuint32_t est[2] = {1, 1};
    return dp_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace ym

using namespace ym;

// Returns 1 when the persistent TMA kernel supports the shape (otherwise ym_moe_dispatch_tc's one-tile-per-CTA kernel is used).
extern "C" int ym_moe_dispatch_v2_supported(int HW, int C, int N, int topk, int ldx, int ldw, int ldo) {
    return dp_encode() != nullptr && C % 64 == 0 && C <= 256 && N % 128 == 0 && N <= 256 && topk >= 1 && topk <= 2 && HW % 128 == 0 &&
           ldx % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0;
}

extern "C" int ym_moe_dispatch_v2(const void* x, int ldx, int B, int HW, int C, const void* w_all, int ldw, int E, const int* route_idx,
                                  const float* route_w, int topk, int N, float w_min, float clamp, void* out, int ldo, void* stream) {
    YM_CHECK_ARG(x && w_all && route_idx && route_w && out, "ym_moe_dispatch_v2: null pointer");
    YM_CHECK_ARG(ym_moe_dispatch_v2_supported(HW, C, N, topk, ldx, ldw, ldo), "ym_moe_dispatch_v2: unsupported shape (HW %% 128, C %% 64 <= 256, "
                 "N %% 128 <= 256, top_k <= 2)");
    YM_CHECK_ARG((((uintptr_t)x | (uintptr_t)w_all | (uintptr_t)out) & 15) == 0, "ym_moe_dispatch_v2: 16-byte alignment");
    if (B == 0) return YM_OK;
    CUtensorMap mx, mw, mo;
    if (!dp_map2d(&mx, x, (cuuint64_t)C, (cuuint64_t)B * HW, (cuuint64_t)ldx * 2, 64, DP_BM) ||
        !dp_map2d(&mw, w_all, (cuuint64_t)C, (cuuint64_t)E * N, (cuuint64_t)ldw * 2, 64, DP_BN) ||
        !dp_map2d(&mo, out, (cuuint64_t)N, (cuuint64_t)B * HW, (cuuint64_t)ldo * 2, 64, DP_BM)) {
        ym_set_error("ym_moe_dispatch_v2: cuTensorMapEncodeTiled failed");
        return YM_ERR_CUDA;
    }
    DispatchParams p;
    p.route_idx = route_idx; p.route_w = route_w; p.topk = topk; p.HW = HW; p.tiles_per_img = HW / DP_BM;
    p.total_tiles = B * p.tiles_per_img; p.K = C; p.N = N; p.w_min = w_min; p.clamp = clamp;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const size_t smem = (size_t)4 * DP_BM * 128 + DP_BSTAGES * DP_BN * 128 + 4 * DP_BM * 128 + 1024;
    cudaError_t e = cudaFuncSetAttribute(tc_dispatch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ym_set_error("ym_moe_dispatch_v2: smem attr %zu: %s", smem, cudaGetErrorString(e)); return YM_ERR_CUDA; }
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    tc_dispatch_kernel<<<grid, DP_THREADS, smem, (cudaStream_t)stream>>>(mx, mw, mo, p);
    YM_CHECK_LAUNCH("tc_dispatch");
    return YM_OK;
}
