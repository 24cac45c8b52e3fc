// ES_MOE router and combine kernels (sm_100a).
//   DynamicRoutingLayer.forward + _hard_top_k   moe/routers.py:458-496,519-527
//   ES_MOE._sparse_forward routing part          moe/modules.py:659-684 (rank>=1 experts below dynamic_threshold dropped,
//                                                 retained weights renormalised)
//   ES_MOE final norm                            moe/modules.py:496,581 (BatchNorm + SiLU on the weighted expert sum)
// The router writes idx int32 [B][k] (-1 = dropped) and w fp32 [B][k]; no (B,E,H,W) weight map is materialised
// (routers.py:496) and no host sync happens (modules.py: `torch.where` per expert).
#include "ym_common.cuh"

namespace ym {

// partial[b][chunk][c] = sum over the chunk's pixels of x[b][pix][c]
__global__ void __launch_bounds__(256) esmoe_gap_kernel(const __half* __restrict__ x, int ldx, int HW, int C, int pix_per_chunk,
                                                        float* __restrict__ partial, int nchunk) {
    const int b = blockIdx.y, ch = blockIdx.x;
    const int p0 = ch * pix_per_chunk, p1 = min(HW, p0 + pix_per_chunk);
    const int cg = C >> 3;                       // 8-channel groups
    const int rows = 256 / cg > 0 ? 256 / cg : 1;
    __shared__ float red[256 * 8];
    const int g = threadIdx.x % cg, r = threadIdx.x / cg;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (r < rows) {
        const __half* xb = x + (long long)b * HW * ldx + g * 8;
        for (int p = p0 + r; p < p1; p += rows) {
            const Half8 v = *reinterpret_cast<const Half8*>(xb + (long long)p * ldx);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 f = __half22float2(v.v[q]);
                acc[2 * q] += f.x;
                acc[2 * q + 1] += f.y;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) red[threadIdx.x * 8 + q] = acc[q];
    __syncthreads();
    if (threadIdx.x < cg) {                      // fixed-order reduction over the row slices
        float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int rr = 0; rr < rows; ++rr)
#pragma unroll
            for (int q = 0; q < 8; ++q) s[q] += red[(rr * cg + threadIdx.x) * 8 + q];
#pragma unroll
        for (int q = 0; q < 8; ++q) partial[((long long)b * nchunk + ch) * C + threadIdx.x * 8 + q] = s[q];
    }
}

// one CTA (128 threads) per image
__global__ void __launch_bounds__(128) esmoe_route_kernel(const float* __restrict__ partial, int nchunk, int HW, int C, int Cr,
                                                          const float* __restrict__ w1, const float* __restrict__ b1,   // [Cr][C], [Cr]
                                                          const float* __restrict__ w2, const float* __restrict__ b2,   // [E][Cr], [E]
                                                          int E, int topk, float dyn_thr, int* __restrict__ idx_out,
                                                          float* __restrict__ w_out, float* __restrict__ probs_out) {
    __shared__ float mean[1024];
    __shared__ float hid[256];
    __shared__ float pr[64];
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int c = tid; c < C; c += 128) {
        float s = 0.f;
        for (int k = 0; k < nchunk; ++k) s += partial[((long long)b * nchunk + k) * C + c];
        mean[c] = s / (float)HW;
    }
    __syncthreads();
    for (int r = tid; r < Cr; r += 128) {
        float a = b1[r];
        for (int c = 0; c < C; ++c) a = fmaf(w1[r * C + c], mean[c], a);
        hid[r] = a / (1.f + expf(-a));
    }
    __syncthreads();
    if (tid < E) {
        float a = b2[tid];
        for (int r = 0; r < Cr; ++r) a = fmaf(w2[tid * Cr + r], hid[r], a);
        pr[tid] = fminf(fmaxf(a, -30.f), 30.f);
    }
    __syncthreads();
    if (tid == 0) {
        float mx = -INFINITY, sum = 0.f;
        for (int e = 0; e < E; ++e) mx = fmaxf(mx, pr[e]);
        for (int e = 0; e < E; ++e) { pr[e] = expf(pr[e] - mx); sum += pr[e]; }
        for (int e = 0; e < E; ++e) { pr[e] /= sum; if (probs_out) probs_out[b * E + e] = pr[e]; }
        float vals[8];
        int ids[8];
        unsigned long long taken = 0ull;
        float tot = 0.f;
        for (int j = 0; j < topk; ++j) {
            int best = 0;
            float bv = -INFINITY;
            for (int e = 0; e < E; ++e)
                if (!((taken >> e) & 1ull) && pr[e] > bv) { bv = pr[e]; best = e; }
            taken |= 1ull << best;
            vals[j] = bv; ids[j] = best; tot += bv;
        }
        tot = fmaxf(tot, 1e-6f);                                   // stable_normalize
        float kept = 0.f;
        for (int j = 0; j < topk; ++j) {
            vals[j] /= tot;
            if (j > 0 && dyn_thr > 0.f && vals[j] < dyn_thr) ids[j] = -1;   // modules.py:674-679
            else kept += vals[j];
        }
        kept = fmaxf(kept, 1.1920929e-7f);                         // finfo(float32).eps
        for (int j = 0; j < topk; ++j) {
            idx_out[b * topk + j] = ids[j];
            w_out[b * topk + j] = ids[j] < 0 ? 0.f : vals[j] / kept;
        }
    }
}

// out[b][pix][c] = SiLU(scale[c] * sum_{j: idx[b][j] >= 0} y[b*topk+j][pix][c] + shift[c])
__global__ void __launch_bounds__(256) esmoe_combine_kernel(const __half* __restrict__ y, int ldy, const int* __restrict__ idx, int topk,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            __half* __restrict__ out, int ldo, int B, int HW, int C) {
    const int cg = C >> 3;
    const long long total = (long long)B * HW * cg;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int g = (int)(i % cg);
    const long long pix = i / cg;
    const int b = (int)(pix / HW), r = (int)(pix % HW);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < topk; ++j) {
        if (idx[b * topk + j] < 0) continue;
        const Half8 v = *reinterpret_cast<const Half8*>(y + ((long long)(b * topk + j) * HW + r) * ldy + g * 8);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 f = __half22float2(v.v[q]);
            acc[2 * q] += f.x;
            acc[2 * q + 1] += f.y;
        }
    }
    Half8 o;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float a0 = silu_f(fmaf(acc[2 * q], scale[g * 8 + 2 * q], shift[g * 8 + 2 * q]));
        const float a1 = silu_f(fmaf(acc[2 * q + 1], scale[g * 8 + 2 * q + 1], shift[g * 8 + 2 * q + 1]));
        o.v[q] = __floats2half2_rn(a0, a1);
    }
    *reinterpret_cast<Half8*>(out + pix * ldo + g * 8) = o;
}

}  // namespace ym

using namespace ym;

extern "C" long long ym_esmoe_scratch_floats(int B, int HW, int C) { return (long long)B * ((HW + 255) / 256) * C; }

extern "C" int ym_esmoe_route(const void* x, int ldx, int B, int HW, int C, const float* w1, const float* b1, int Cr, const float* w2,
                              const float* b2, int E, int topk, float dyn_thr, float* scratch, int* idx_out, float* w_out,
                              float* probs_out, void* stream) {
    YM_CHECK_ARG(x && w1 && b1 && w2 && b2 && scratch && idx_out && w_out, "ym_esmoe_route: null pointer");
    YM_CHECK_ARG(C % 8 == 0 && C <= 1024 && Cr <= 256 && E <= 64 && topk >= 1 && topk <= 8 && topk <= E && ldx % 8 == 0,
                 "ym_esmoe_route: C<=1024 (mult of 8), Cr<=256, E<=64, topk<=8");
    if (B == 0) return YM_OK;
    const int nchunk = (HW + 255) / 256;
    esmoe_gap_kernel<<<dim3(nchunk, B), 256, 0, (cudaStream_t)stream>>>((const __half*)x, ldx, HW, C, 256, scratch, nchunk);
    YM_CHECK_LAUNCH("esmoe_gap");
    esmoe_route_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(scratch, nchunk, HW, C, Cr, w1, b1, w2, b2, E, topk, dyn_thr, idx_out,
                                                           w_out, probs_out);
    YM_CHECK_LAUNCH("esmoe_route");
    return YM_OK;
}

extern "C" int ym_esmoe_combine(const void* y, int ldy, const int* route_idx, int topk, const float* scale, const float* shift,
                                void* out, int ldo, int B, int HW, int C, void* stream) {
    YM_CHECK_ARG(y && route_idx && scale && shift && out, "ym_esmoe_combine: null pointer");
    YM_CHECK_ARG(C % 8 == 0 && ldy % 8 == 0 && ldo % 8 == 0, "ym_esmoe_combine: multiples of 8");
    if (B == 0) return YM_OK;
    const long long total = (long long)B * HW * (C / 8);
    esmoe_combine_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __half*)y, ldy, route_idx, topk, scale,
                                                                                      shift, (__half*)out, ldo, B, HW, C);
    YM_CHECK_LAUNCH("esmoe_combine");
    return YM_OK;
}
