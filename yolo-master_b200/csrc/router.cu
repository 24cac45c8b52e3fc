// EfficientSpatialRouter on device: avg-pool -> conv3x3 -> BN -> SiLU -> (spatial mean) -> conv1x1 -> BN -> softmax
// -> top-k -> renormalise, writing (idx int32[B,k], w fp32[B,k], probs fp32[B,E]).
// Reference: moe/routers.py:283-304 (forward), :185-265 (_process_logits, eval branch).
//
// The reference applies conv1x1+BN per pooled pixel and then takes the fp32 spatial mean (routers.py:300); both are
// affine, so mean(BN(conv1x1(h))) == BN(conv1x1(mean(h))) and the [B,E,H',W'] map is never materialised.  All router
// arithmetic here is fp32 on the fp16 activations (the reference's own fp32 contract covers mean/softmax/top-k).
// Replaces ~10 launches + 3 host syncs (routers.py:51,295,301) with 2 small launches and no sync: routing stays on
// the device as an index table consumed by ym_moe_expert_gemm.
#include "ym_common.cuh"
#include "router_core.cuh"

namespace ym {

// Fused avg-pool + conv3x3 + BN + SiLU + per-tile sums.  (Round-1 history: separate pool / hidden kernels ran at 41 GB/s,
// L1-instruction bound - two 16-byte loads per four FMAs - with the pooled map making a round trip through global memory;
// profiles/r01_launch_roofline.txt.)  w1 layout [tap][c/4][r][4] fp32: the Cr lanes of a quad read consecutive float4.
// One CTA = a 4 x 16 tile of POOLED pixels of one image (a 2 x 16 tile was measured: more CTAs but 2.25x halo re-reads, 36 -> 54 us at P3):
//   phase 1: the haloed 6 x 18 pooled tile is averaged straight from the fp16 activation into shared memory (fp32);
//   phase 2: each thread owns one reduced channel r and FOUR horizontally adjacent pixels: per (ky, c4) it loads 6 input
//            float4 (shared memory, broadcast across the r lanes) and 3 weight float4 (L1) for 48 FMAs (was 2 loads per 4).
// partial[b, tile, r] = sum over the tile's pixels of SiLU(scale1*conv + shift1); router_finish_kernel is unchanged.
constexpr int RT_TY = 4, RT_TX = 16, RT_HY = RT_TY + 2, RT_HX = RT_TX + 2, RT_QUADS = RT_TY * (RT_TX / 4);

// The input-channel reduction of an item is split over S = blockDim / (RT_QUADS * Cr) slices (threads of the same (quad, r), different
// channel ranges; up to 512 threads per CTA): the per-thread chain of weight loads (L2-latency bound: 288 dependent iterations at
// C = 128) shrinks S-fold, and at Cr = 8 the half of the CTA that had no item gets one.  Slice partials are summed in slice order.
__global__ void __launch_bounds__(512, 2) router_fused_kernel(const __half* __restrict__ x, int ldx, int H, int W, int C, int ps,
                                                           int Hp, int Wp, int Cr, const float* __restrict__ w1,
                                                           const float* __restrict__ scale1, const float* __restrict__ shift1,
                                                           float* __restrict__ partial, int tiles_x, int nblk) {
    pdl_prologue();
    extern __shared__ float rsm[];
    float* sp = rsm;                                   // [RT_HY][RT_HX][C] pooled tile with halo (zero outside the map)
    float* red = rsm + RT_HY * RT_HX * C;              // [RT_QUADS][Cr]
    const int b = blockIdx.y;
    const int ty0 = (blockIdx.x / tiles_x) * RT_TY, tx0 = (blockIdx.x % tiles_x) * RT_TX;
    const int C8 = C >> 3;
    const float inv = 1.f / (float)(ps * ps);
    const __half* xb = x + (long long)b * H * W * ldx;
    for (int i = threadIdx.x; i < RT_HY * RT_HX * C8; i += blockDim.x) {
        const int c8 = i % C8, pp = i / C8;
        const int hy = pp / RT_HX, hx = pp - hy * RT_HX;
        const int py = ty0 + hy - 1, px = tx0 + hx - 1;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        if (py >= 0 && py < Hp && px >= 0 && px < Wp) {
            for (int dy = 0; dy < ps; ++dy)
                for (int dx = 0; dx < ps; ++dx) {
                    const Half8 h = *reinterpret_cast<const Half8*>(xb + ((long long)(py * ps + dy) * W + (px * ps + dx)) * ldx + c8 * 8);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __half22float2(h.v[j]);
                        acc[2 * j] += f.x;
                        acc[2 * j + 1] += f.y;
                    }
                }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] *= inv;
        }
        float4* dst = reinterpret_cast<float4*>(sp + (size_t)pp * C + c8 * 8);
        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    __syncthreads();
    const int C4 = C >> 2;
    const int nitems = RT_QUADS * Cr;                   // (quad, r)
    const int S = max(1, (int)blockDim.x / nitems);     // channel slices per item
    float* red4 = red + RT_QUADS * Cr;                  // [S][nitems][4] per-slice conv partials of the four pixels
    for (int wi = threadIdx.x; wi < nitems * S; wi += blockDim.x) {
        const int it = wi % nitems, slice = wi / nitems;
        {
            const int r = it % Cr, quad = it / Cr;
            const int qy = quad >> 2, qx = quad & 3;        // pixels (ty0 + qy, tx0 + 4*qx + 0..3)
            const int c4_lo = (int)((long long)C4 * slice / S), c4_hi = (int)((long long)C4 * (slice + 1) / S);
            float a[4][4];
#pragma unroll
            for (int pxl = 0; pxl < 4; ++pxl)
#pragma unroll
                for (int l = 0; l < 4; ++l) a[pxl][l] = 0.f;
            const float4* wbase = reinterpret_cast<const float4*>(w1) + r;
            for (int ky = 0; ky < 3; ++ky) {
                const float* srow = sp + (size_t)((qy + ky) * RT_HX + 4 * qx) * C;
#pragma unroll 2
                for (int c4 = c4_lo; c4 < c4_hi; ++c4) {
                    float4 wv[3];
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) wv[kx] = __ldg(wbase + ((long long)(ky * 3 + kx) * C4 + c4) * Cr);
                    float4 xv[6];
#pragma unroll
                    for (int j = 0; j < 6; ++j) xv[j] = *reinterpret_cast<const float4*>(srow + (size_t)j * C + c4 * 4);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                        for (int pxl = 0; pxl < 4; ++pxl) {
                            a[pxl][0] = fmaf(xv[pxl + kx].x, wv[kx].x, a[pxl][0]);
                            a[pxl][1] = fmaf(xv[pxl + kx].y, wv[kx].y, a[pxl][1]);
                            a[pxl][2] = fmaf(xv[pxl + kx].z, wv[kx].z, a[pxl][2]);
                            a[pxl][3] = fmaf(xv[pxl + kx].w, wv[kx].w, a[pxl][3]);
                        }
                    }
                }
            }
            float4 part;
            part.x = (a[0][0] + a[0][1]) + (a[0][2] + a[0][3]);
            part.y = (a[1][0] + a[1][1]) + (a[1][2] + a[1][3]);
            part.z = (a[2][0] + a[2][1]) + (a[2][2] + a[2][3]);
            part.w = (a[3][0] + a[3][1]) + (a[3][2] + a[3][3]);
            reinterpret_cast<float4*>(red4)[slice * nitems + it] = part;
        }
    }
    __syncthreads();
    for (int it = threadIdx.x; it < nitems; it += blockDim.x) {
        const int r = it % Cr, quad = it / Cr;
        const int qy = quad >> 2, qx = quad & 3;
        float4 acc = reinterpret_cast<const float4*>(red4)[it];
        for (int sl = 1; sl < S; ++sl) {                  // fixed slice order: deterministic
            const float4 t = reinterpret_cast<const float4*>(red4)[sl * nitems + it];
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        const float conv[4] = {acc.x, acc.y, acc.z, acc.w};
        const float sc = scale1[r], sh = shift1[r];
        float hsum = 0.f;
#pragma unroll
        for (int pxl = 0; pxl < 4; ++pxl) {
            const int py = ty0 + qy, px = tx0 + 4 * qx + pxl;
            if (py < Hp && px < Wp) {
                const float v = fmaf(conv[pxl], sc, sh);
                hsum += v / (1.f + expf(-v));
            }
        }
        red[quad * Cr + r] = hsum;
    }
    __syncthreads();
    if ((int)threadIdx.x < Cr) {   // fixed-order (deterministic) reduction over the tile's quads
        float s2 = 0.f;
        for (int q = 0; q < RT_QUADS; ++q) s2 += red[q * Cr + threadIdx.x];
        partial[((long long)b * nblk + blockIdx.x) * Cr + threadIdx.x] = s2;
    }
}

// One warp per image: mean hidden -> logits -> softmax -> top-k (lowest index wins ties) -> renormalise (router_core.cuh).
__global__ void __launch_bounds__(32) router_finish_kernel(const RouterFin r) {
    pdl_prologue();
    __shared__ float hm[64];
    __shared__ float pr[64];
    int ids[8];
    float vals[8];
    router_finish_warp(r, blockIdx.x, threadIdx.x, hm, pr, ids, vals, true);
}

}  // namespace ym

using namespace ym;

// scratch: fp32, at least B*Hp*Wp*C + B*nblk*Cr floats (query with ym_router_scratch_floats).
extern "C" long long ym_router_scratch_floats(int B, int H, int W, int C, int Cr, int pool) {
    const bool do_pool = H > pool && W > pool;
    const int Hp = do_pool ? H / pool : H, Wp = do_pool ? W / pool : W;
    const int ppb = 256 / Cr;
    const int nblk = (Hp * Wp + ppb - 1) / ppb;
    return (long long)B * Hp * Wp * C + (long long)B * nblk * Cr;
}

// Tile grid of router_fused_kernel over the pooled map (also the number of partial-sum blocks per image); *npix = pooled pixels.
extern "C" int ym_router_blocks(int H, int W, int pool, int* npix) {
    const bool do_pool = H > pool && W > pool;  // routers.py:289-292
    const int ps = do_pool ? pool : 1;
    const int Hp = H / ps, Wp = W / ps;
    if (npix) *npix = Hp * Wp;
    return ((Wp + RT_TX - 1) / RT_TX) * ((Hp + RT_TY - 1) / RT_TY);
}

// First half of ym_router_topk: the fused pool + conv3x3 + BN + SiLU pass, leaving per-tile partial sums [B][nblk][Cr] at the start of
// `scratch`.  The second half (router_finish_warp) runs either as its own launch (ym_router_topk) or inside the prologue of the
// consumer (ym_moe_ffn_routed).
extern "C" int ym_router_partial(const void* x, int ldx, int B, int H, int W, int C, int pool, const float* w1, int Cr, const float* scale1,
                                 const float* shift1, float* scratch, void* stream) {
    YM_CHECK_ARG(x && w1 && scale1 && shift1 && scratch, "ym_router_partial: null pointer");
    YM_CHECK_ARG(Cr == 8 || Cr == 16 || Cr == 32 || Cr == 64, "ym_router_topk: reduced channels must be 8/16/32/64 (got %d)", Cr);
    YM_CHECK_ARG(pool >= 1, "ym_router_topk: pool");
    if (B == 0) return YM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool do_pool = H > pool && W > pool;  // routers.py:289-292
    const int ps = do_pool ? pool : 1;
    const int Hp = H / ps, Wp = W / ps;
    YM_CHECK_ARG(C % 8 == 0 && ldx % 8 == 0 && (((uintptr_t)x) & 15) == 0, "ym_router_topk: C / pitch must be multiples of 8 halves");
    const int tiles_x = (Wp + RT_TX - 1) / RT_TX, tiles_y = (Hp + RT_TY - 1) / RT_TY;
    const int nblk = tiles_x * tiles_y;
    float* partial = scratch;
    const int nitems = RT_QUADS * Cr;
    int threads = nitems;                                    // S channel slices per item, at most 4 and at most 512 threads
    while (threads * 2 <= 512 && threads * 2 <= 4 * nitems && (C / 4) >= 2 * (threads / nitems) * 2) threads *= 2;
    if (threads < 256) threads = 256;                        // the pooling phase wants a full CTA (extra threads carry no conv item)
    const int S = threads / nitems > 0 ? threads / nitems : 1;
    const size_t smem = ((size_t)RT_HY * RT_HX * C + RT_QUADS * (size_t)Cr + (size_t)S * nitems * 4) * sizeof(float);
    YM_CHECK_ARG(smem <= 200 * 1024, "ym_router_topk: C=%d too wide for the fused router tile", C);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(router_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ym_set_error("ym_router_topk: smem attr: %s", cudaGetErrorString(e)); return YM_ERR_CUDA; }
        smem_set = smem;
    }
    launch_pdl(router_fused_kernel, dim3(nblk, B), threads, smem, st, (const __half*)x, ldx, H, W, C, ps, Hp, Wp, Cr, w1, scale1, shift1, partial,
                                                         tiles_x, nblk);
    YM_CHECK_LAUNCH("router_fused");
    return YM_OK;
}

extern "C" int ym_router_topk(const void* x, int ldx, int B, int H, int W, int C, int pool, const float* w1, int Cr,
                              const float* scale1, const float* shift1, const float* w2, const float* scale2,
                              const float* shift2, int E, int topk, float* scratch, int* idx_out, float* w_out,
                              float* probs_out, void* stream) {
    YM_CHECK_ARG(x && w1 && scale1 && shift1 && w2 && scale2 && shift2 && scratch && idx_out && w_out,
                 "ym_router_topk: null pointer");
    YM_CHECK_ARG(E >= 1 && E <= 64 && topk >= 1 && topk <= 8 && topk <= E, "ym_router_topk: need 1<=topk<=min(8,E), E<=64");
    if (B == 0) return YM_OK;
    const int rc = ym_router_partial(x, ldx, B, H, W, C, pool, w1, Cr, scale1, shift1, scratch, stream);
    if (rc) return rc;
    RouterFin r;
    r.partial = scratch; r.nblk = ym_router_blocks(H, W, pool, &r.npix); r.Cr = Cr; r.w2 = w2; r.scale2 = scale2; r.shift2 = shift2;
    r.E = E; r.topk = topk; r.idx_out = idx_out; r.w_out = w_out; r.probs_out = probs_out;
    launch_pdl(router_finish_kernel, B, 32, 0, (cudaStream_t)stream, r);
    YM_CHECK_LAUNCH("router_finish");
    return YM_OK;
}
