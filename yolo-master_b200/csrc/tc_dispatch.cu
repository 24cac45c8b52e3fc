// ES-MoE dispatch, persistent TMA + tcgen05 version (BatchedExpertComputation.compute_sparse_experts_batched with 1x1-conv
// experts, moe/utils.py:119-209):   out[b] = clamp( sum_j fp16( fp16(x[b] W[e_bj]^T) * w_bj ), +-clamp ),  w <= w_min dropped.
//
// Routing is per image, so the "gather" is a row range and the "scatter" is the output tile itself: the kernel reads every
// 128-token tile of x ONCE (TMA, 4 x 16 KB swizzled k-chunks kept resident, double-buffered so the next tile's HBM read
// overlaps this tile's MMAs), streams the <= 2 routed experts' weight tiles through a 4-deep TMA ring, accumulates each expert in its own TMEM accumulator (tcgen05.mma, M=128, N=128, K=256) and
// combines them with the routing weights in the epilogue, which leaves through a swizzled staging tile and TMA stores.
// Algorithmic HBM traffic = (k+1)*d*2 bytes per token (x in, out) + the expert weights once (they live in L2).
//   warp 0: TMA producer | warp 1: MMA issuer | warps 2..9: epilogue (TMEM lane quarter = warp%4, column half = (warp-2)/4)
// TMEM: 2 (double buffer over N halves) x 2 (experts) x 128 columns = 512.
#include <cuda.h>

#include "tc_common.cuh"

namespace ym {

constexpr int DP_THREADS = 320, DP_BM = 128, DP_BN = 128, DP_KC = 64, DP_BSTAGES = 4;

struct DispatchParams {
    const int* route_idx; const float* route_w;
    int topk, HW, tiles_per_img, total_tiles, K, N;   // K = d (multiple of 64, <= 256), N = outputs (multiple of 128, <= 256)
    float w_min, clamp;
    int debug;   // profiling only (ym_set_dispatch_debug): 1 skip weight loads, 2 skip x loads, 4 skip stores, 8 skip epilogue math, 16 skip MMAs
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
// first tile >= t (stepping by the grid) whose image has at least one live route; total_tiles when there is none
__device__ __forceinline__ int dp_next_live(const DispatchParams& p, int t, int step) {
    int e[2]; float w[2];
    for (; t < p.total_tiles; t += step)
        if (dp_routes(p, t / p.tiles_per_img, e, w) > 0) return t;
    return p.total_tiles;
}

__global__ void __launch_bounds__(DP_THREADS, 1) tc_dispatch_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                    const __grid_constant__ CUtensorMap map_w,
                                                                    const __grid_constant__ CUtensorMap map_o, const DispatchParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (STS, not generic ST)
    constexpr int A_CHUNK = DP_BM * 128, A_TILE = 4 * A_CHUNK, B_TILE = DP_BN * 128, STG_HALF = DP_BM * 128;
    const int kchunks = p.K / DP_KC;                       // <= 4
    unsigned char* sA = smem;                              // [2 tiles][4][A_CHUNK]  128 KB: the next x tile lands while this one is multiplied
    unsigned char* sB = sA + 2 * A_TILE;                   // [DP_BSTAGES][B_TILE]    64 KB
    unsigned char* stg = sB + DP_BSTAGES * B_TILE;         // [2 halves][STG_HALF]    32 KB
    __shared__ uint64_t a_full[2], a_empty[2], b_full[DP_BSTAGES], b_empty[DP_BSTAGES], t_full[2], t_empty[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&a_full[a], 1); tc::mbar_init(&a_empty[a], 1); }
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
            // titer counts PROCESSED tiles only: x tile buffer = titer & 1, phase of a_full / a_empty = (titer >> 1) & 1
            uint32_t titer = 0, bidx = 0;
            auto load_x = [&](int tile, uint32_t it) {
                const int b = tile / p.tiles_per_img, m0 = b * p.HW + (tile - b * p.tiles_per_img) * DP_BM;
                const uint32_t ab = it & 1;
                if (p.debug & 2) return;
                tc::mbar_wait(&a_empty[ab], ((it >> 1) & 1) ^ 1);
                dp_expect_tx(&a_full[ab], (uint32_t)(kchunks * A_CHUNK));
                for (int kc = 0; kc < kchunks; ++kc) dp_tma_load_2d(sA + ab * A_TILE + kc * A_CHUNK, &map_x, kc * DP_KC, m0, &a_full[ab]);
            };
            int tile = dp_next_live(p, blockIdx.x, gridDim.x);
            if (tile < p.total_tiles) load_x(tile, 0);
            while (tile < p.total_tiles) {
                const int b = tile / p.tiles_per_img;
                int e[2]; float w[2];
                const int nr = dp_routes(p, b, e, w);
                const int next = dp_next_live(p, tile + gridDim.x, gridDim.x);
                bool next_issued = next >= p.total_tiles;
                int local = 0;
                for (int nh = 0; nh < nhalves; ++nh)
                    for (int j = 0; j < nr; ++j)
                        for (int kc = 0; kc < kchunks; ++kc, ++bidx, ++local) {
                            // once a ring-full of this tile's weights is in flight the previous tile's MMAs have retired
                            // (a b_empty wait below could only pass after them), so its x buffer is free: prefetch the next tile
                            if (!next_issued && local == DP_BSTAGES) { load_x(next, titer + 1); next_issued = true; }
                            if (p.debug & 1) continue;
                            const int s = bidx % DP_BSTAGES;
                            tc::mbar_wait(&b_empty[s], ((bidx / DP_BSTAGES) & 1) ^ 1);
                            dp_expect_tx(&b_full[s], (uint32_t)B_TILE);
                            dp_tma_load_2d(sB + s * B_TILE, &map_w, kc * DP_KC, e[j] * p.N + nh * DP_BN, &b_full[s]);
                        }
                if (!next_issued) load_x(next, titer + 1);
                ++titer;
                tile = next;
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
                const uint32_t ab = titer & 1;
                if (!(p.debug & 2)) tc::mbar_wait(&a_full[ab], (titer >> 1) & 1);
                tc::fence_after_sync();
                for (int nh = 0; nh < nhalves; ++nh, ++it2) {
                    const uint32_t buf = it2 & 1;
                    tc::mbar_wait(&t_empty[buf], ((it2 >> 1) & 1) ^ 1);
                    tc::fence_after_sync();
                    for (int j = 0; j < nr; ++j) {
                        const uint32_t tacc = tmem_base + buf * 256 + j * 128;
                        for (int kc = 0; kc < kchunks; ++kc, ++bidx) {
                            const int s = bidx % DP_BSTAGES;
                            if (!(p.debug & 1)) tc::mbar_wait(&b_full[s], (bidx / DP_BSTAGES) & 1);
                            tc::fence_after_sync();
                            const uint64_t adesc = tc::make_desc(smem_u32(sA + ab * A_TILE + kc * A_CHUNK), 1024, 2);
                            const uint64_t bdesc = tc::make_desc(smem_u32(sB + s * B_TILE), 1024, 2);
                            if (!(p.debug & 16)) {
#pragma unroll
                                for (int k = 0; k < DP_KC / 16; ++k) tc::mma_f16_ss(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) ? 1u : 0u);
                            }
                            tc::mma_commit(&b_empty[s]);
                        }
                    }
                    tc::mma_commit(&t_full[buf]);
                }
                tc::mma_commit(&a_empty[ab]);               // all MMAs reading this tile's x chunks are done
                ++titer;
            }
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;     // 64 columns per thread
        const int r = q * 32 + lane;
        const bool elected = (warp == 2 && lane == 0);
        unsigned char* sb = stg + half * STG_HALF;          // this thread's 64-column half of the staging tile
        const __half2 hi2 = __float2half2_rn(p.clamp), lo2 = __float2half2_rn(-p.clamp);
        uint32_t it2 = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int b = tile / p.tiles_per_img, m0 = b * p.HW + (tile - b * p.tiles_per_img) * DP_BM;
            int e[2]; float w[2];
            const int nr = dp_routes(p, b, e, w);
            for (int nh = 0; nh < nhalves; ++nh) {
                uint32_t buf = 0;
                if (nr > 0) {
                    buf = it2 & 1;
                    tc::mbar_wait(&t_full[buf], (it2 >> 1) & 1);
                    tc::fence_after_sync();
                }
                // out = sum_j fp16( fp32(fp16(x W_j^T)) * w_j ) accumulated in fp16 like index_add_ on an fp16 tensor (utils.py:200-203)
                __half2 acc[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] = __float2half2_rn(0.f);
                for (int j = 0; j < ((p.debug & 8) ? 0 : nr); ++j) {
#pragma unroll
                    for (int c0 = 0; c0 < 64; c0 += 32) {
                        uint32_t rr[32];
                        tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256 + j * 128 + half * 64 + c0, rr);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float2 f = __half22float2(__floats2half2_rn(__uint_as_float(rr[2 * i]), __uint_as_float(rr[2 * i + 1])));
                            f.x *= w[j];
                            f.y *= w[j];
                            acc[c0 / 2 + i] = __hadd2(acc[c0 / 2 + i], __floats2half2_rn(f.x, f.y));
                        }
                    }
                }
                if (nr > 0) {                               // accumulators are in registers: hand the TMEM buffer back to the MMA warp
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&t_empty[buf]);
                    ++it2;
                }
                if (elected) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // previous stores have read the staging tile
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint32_t pk[4];
#pragma unroll
                    for (int h2 = 0; h2 < 4; ++h2) {
                        const __half2 v = __hmin2(__hmax2(acc[c * 4 + h2], lo2), hi2);
                        pk[h2] = *reinterpret_cast<const uint32_t*>(&v);
                    }
                    *reinterpret_cast<uint4*>(sb + tc::sw128_offset(r, c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
                tc::fence_proxy_async();
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
                if (elected && !(p.debug & 4)) {
                    dp_tma_store_2d(&map_o, stg, nh * DP_BN, m0);
                    dp_tma_store_2d(&map_o, stg + STG_HALF, nh * DP_BN + 64, m0);
                    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                }
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

static int g_dispatch_debug = 0;
// Profiling aid (tools/profile_dispatch.py): disables pipeline stages of tc_dispatch_kernel so their cost can be read off the
// kernel time.  Results are WRONG for any non-zero mask; the default is 0 and nothing in the package sets it.
extern "C" void ym_set_dispatch_debug(int mask) { g_dispatch_debug = mask; }
extern "C" int ym_dispatch_debug_mask() { return g_dispatch_debug; }

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
    p.total_tiles = B * p.tiles_per_img; p.K = C; p.N = N; p.w_min = w_min; p.clamp = clamp; p.debug = g_dispatch_debug;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const size_t smem = (size_t)8 * DP_BM * 128 + DP_BSTAGES * DP_BN * 128 + 2 * DP_BM * 128 + 1024;
    cudaError_t e = cudaFuncSetAttribute(tc_dispatch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ym_set_error("ym_moe_dispatch_v2: smem attr %zu: %s", smem, cudaGetErrorString(e)); return YM_ERR_CUDA; }
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    tc_dispatch_kernel<<<grid, DP_THREADS, smem, (cudaStream_t)stream>>>(mx, mw, mo, p);
    YM_CHECK_LAUNCH("tc_dispatch");
    return YM_OK;
}
