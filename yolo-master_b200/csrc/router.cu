// EfficientSpatialRouter on device: avg-pool -> conv3x3 -> BN -> SiLU -> (spatial mean) -> conv1x1 -> BN -> softmax
// -> top-k -> renormalise, writing (idx int32[B,k], w fp32[B,k], probs fp32[B,E]).
// Reference: moe/routers.py:283-304 (forward), :185-265 (_process_logits, eval branch).
//
// The reference applies conv1x1+BN per pooled pixel and then takes the fp32 spatial mean (routers.py:300); both are
// affine, so mean(BN(conv1x1(h))) == BN(conv1x1(mean(h))) and the [B,E,H',W'] map is never materialised.  All router
// arithmetic here is fp32 on the fp16 activations (the reference's own fp32 contract covers mean/softmax/top-k).
// Replaces ~10 launches + 3 host syncs (routers.py:51,295,301) with 3 small launches and no sync: routing stays on
// the device as an index table consumed by ym_moe_expert_gemm.
#include "ym_common.cuh"

namespace ym {

// pooled[b, py, px, c] = mean over the ps x ps block (F.avg_pool2d, floor semantics), fp32
__global__ void __launch_bounds__(256) router_pool_kernel(const __half* __restrict__ x, int ldx, int B, int H, int W, int C,
                                                          int ps, int Hp, int Wp, float* __restrict__ pooled) {
    const long long total = (long long)B * Hp * Wp * C;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % C);
    const long long pp = idx / C;
    const int px = (int)(pp % Wp);
    const int py = (int)((pp / Wp) % Hp);
    const int b = (int)(pp / ((long long)Wp * Hp));
    const __half* xb = x + ((long long)b * H * W) * ldx + c;
    float s = 0.f;
    for (int dy = 0; dy < ps; ++dy)
        for (int dx = 0; dx < ps; ++dx) s += __half2float(xb[((long long)(py * ps + dy) * W + (px * ps + dx)) * ldx]);
    pooled[idx] = s / (float)(ps * ps);
}

// hidden[b, pix, r] = SiLU(scale1[r] * conv3x3(pooled)[pix, r] + shift1[r]); partial[b, blk, r] = sum over the block's pixels.
// grid (nblk, B), block 256 = PIX_PER_BLOCK pixels x Cr lanes (Cr in {8,16,32,64}).  w1 layout [tap][c/4][r][4] fp32 so that
// the Cr lanes of a pixel read consecutive float4 (coalesced) while the pooled input float4 is a broadcast.
__global__ void __launch_bounds__(256) router_hidden_kernel(const float* __restrict__ pooled, int Hp, int Wp, int C, int Cr,
                                                            const float* __restrict__ w1, const float* __restrict__ scale1,
                                                            const float* __restrict__ shift1, float* __restrict__ partial,
                                                            int nblk) {
    __shared__ float red[256];
    const int b = blockIdx.y;
    const int ppb = 256 / Cr;
    const int r = threadIdx.x % Cr;
    const int pl = threadIdx.x / Cr;
    const int pix = blockIdx.x * ppb + pl;
    const int C4 = C >> 2;
    float hval = 0.f;
    if (pix < Hp * Wp) {
        const int py = pix / Wp, px = pix % Wp;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = py + ky - 1;
            if (iy < 0 || iy >= Hp) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = px + kx - 1;
                if (ix < 0 || ix >= Wp) continue;
                const float4* pin = reinterpret_cast<const float4*>(pooled + (((long long)b * Hp + iy) * Wp + ix) * C);
                const float4* pw = reinterpret_cast<const float4*>(w1) + (long long)(ky * 3 + kx) * C4 * Cr + r;
#pragma unroll 4
                for (int c4 = 0; c4 < C4; ++c4) {
                    const float4 xv = __ldg(pin + c4);
                    const float4 wv = __ldg(pw + (long long)c4 * Cr);
                    a0 = fmaf(xv.x, wv.x, a0);
                    a1 = fmaf(xv.y, wv.y, a1);
                    a2 = fmaf(xv.z, wv.z, a2);
                    a3 = fmaf(xv.w, wv.w, a3);
                }
            }
        }
        const float v = fmaf((a0 + a1) + (a2 + a3), scale1[r], shift1[r]);
        hval = v / (1.f + expf(-v));
    }
    red[threadIdx.x] = hval;
    __syncthreads();
    if (threadIdx.x < Cr) {  // fixed-order (deterministic) reduction over the block's pixels
        float s = 0.f;
        for (int i = 0; i < ppb; ++i) s += red[i * Cr + threadIdx.x];
        partial[((long long)b * nblk + blockIdx.x) * Cr + threadIdx.x] = s;
    }
}

// One warp per image: mean hidden -> logits -> softmax -> top-k (lowest index wins ties) -> renormalise.
__global__ void __launch_bounds__(32) router_finish_kernel(const float* __restrict__ partial, int nblk, int Cr, int npix,
                                                           const float* __restrict__ w2,      // [E][Cr]
                                                           const float* __restrict__ scale2,  // [E]
                                                           const float* __restrict__ shift2, int E, int topk,
                                                           int* __restrict__ idx_out, float* __restrict__ w_out,
                                                           float* __restrict__ probs_out) {
    __shared__ float hm[64];
    __shared__ float pr[64];
    const int b = blockIdx.x, lane = threadIdx.x;
    for (int r = lane; r < Cr; r += 32) {
        float s = 0.f;
        for (int i = 0; i < nblk; ++i) s += partial[((long long)b * nblk + i) * Cr + r];
        hm[r] = s / (float)npix;
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int e = lane; e < E; e += 32) {
        float acc = 0.f;
        for (int r = 0; r < Cr; ++r) acc = fmaf(w2[e * Cr + r], hm[r], acc);
        const float lg = fmaf(acc, scale2[e], shift2[e]);
        pr[e] = lg;
        mx = fmaxf(mx, lg);
    }
    mx = warp_max(mx);
    __syncwarp();
    float sum = 0.f;
    for (int e = lane; e < E; e += 32) {
        const float v = expf(pr[e] - mx);
        pr[e] = v;
        sum += v;
    }
    sum = warp_sum(sum);
    __syncwarp();
    for (int e = lane; e < E; e += 32) {
        pr[e] = pr[e] / sum;
        if (probs_out) probs_out[(long long)b * E + e] = pr[e];
    }
    __syncwarp();
    if (lane == 0) {
        float vals[8];
        int ids[8];
        unsigned long long taken = 0ull;
        float tot = 0.f;
        for (int j = 0; j < topk; ++j) {
            int best = -1;
            float bv = -INFINITY;
            for (int e = 0; e < E; ++e) {
                if ((taken >> e) & 1ull) continue;
                if (pr[e] > bv) { bv = pr[e]; best = e; }
            }
            taken |= 1ull << best;
            vals[j] = bv;
            ids[j] = best;
            tot += bv;
        }
        tot = fmaxf(tot, 1e-6f);
        for (int j = 0; j < topk; ++j) {
            idx_out[b * topk + j] = ids[j];
            w_out[b * topk + j] = vals[j] / tot;
        }
    }
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

extern "C" int ym_router_topk(const void* x, int ldx, int B, int H, int W, int C, int pool, const float* w1, int Cr,
                              const float* scale1, const float* shift1, const float* w2, const float* scale2,
                              const float* shift2, int E, int topk, float* scratch, int* idx_out, float* w_out,
                              float* probs_out, void* stream) {
    YM_CHECK_ARG(x && w1 && scale1 && shift1 && w2 && scale2 && shift2 && scratch && idx_out && w_out,
                 "ym_router_topk: null pointer");
    YM_CHECK_ARG(Cr == 8 || Cr == 16 || Cr == 32 || Cr == 64, "ym_router_topk: reduced channels must be 8/16/32/64 (got %d)", Cr);
    YM_CHECK_ARG(E >= 1 && E <= 64 && topk >= 1 && topk <= 8 && topk <= E, "ym_router_topk: need 1<=topk<=min(8,E), E<=64");
    YM_CHECK_ARG(pool >= 1, "ym_router_topk: pool");
    YM_CHECK_ARG(C % 4 == 0, "ym_router_topk: C must be a multiple of 4");
    if (B == 0) return YM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool do_pool = H > pool && W > pool;  // routers.py:289-292
    const int ps = do_pool ? pool : 1;
    const int Hp = H / ps, Wp = W / ps;
    float* pooled = scratch;
    const long long npool = (long long)B * Hp * Wp * C;
    float* partial = scratch + npool;
    router_pool_kernel<<<(int)((npool + 255) / 256), 256, 0, st>>>((const __half*)x, ldx, B, H, W, C, ps, Hp, Wp, pooled);
    YM_CHECK_LAUNCH("router_pool");
    const int ppb = 256 / Cr;
    const int nblk = (Hp * Wp + ppb - 1) / ppb;
    router_hidden_kernel<<<dim3(nblk, B), 256, 0, st>>>(pooled, Hp, Wp, C, Cr, w1, scale1, shift1, partial, nblk);
    YM_CHECK_LAUNCH("router_hidden");
    router_finish_kernel<<<B, 32, 0, st>>>(partial, nblk, Cr, Hp * Wp, w2, scale2, shift2, E, topk, idx_out, w_out, probs_out);
    YM_CHECK_LAUNCH("router_finish");
    return YM_OK;
}
