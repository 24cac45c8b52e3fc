// Kernels of the Mixture-of-Transformers / Mixture-of-Attention neck blocks (C2fMoT, C2fMoA), NHWC fp16 activations.
// Everything here is HBM / latency bound (head_dim 8..32, 49-token windows, 3-expert routers): SIMT kernels, one query row
// per thread (shuffle-free online softmax), K/V staged in shared memory, 16-byte global accesses, deterministic reductions.
#include <math.h>

#include "ym_common.cuh"

namespace ym {

// ---------------------------------------------------------------------------------------------------------------------
// block-wide deterministic sum (fixed tree order)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}

// GroupNorm statistics -> per (image, channel) affine: scale = rstd*gamma, shift = beta - mean*rstd*gamma.
// One CTA per (group, image); VEC consecutive channels of a pixel are read with one 2*VEC- or 4*VEC-byte load (the scalar
// version ran at 158 GB/s: profiles/r01_launch_roofline_moa_mot_n.txt).  Two passes (mean, then centred squares) for accuracy.
template <typename T, int VEC>
__device__ __forceinline__ void gn_load(const T* __restrict__ p, float (&v)[VEC]) {
    if constexpr (sizeof(T) == 2 && VEC == 8) {
        const Half8 h = *reinterpret_cast<const Half8*>(p);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h.v[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
    } else if constexpr (sizeof(T) == 2 && VEC == 4) {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        const float2 f0 = unpack_half2(u.x), f1 = unpack_half2(u.y);
        v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
    } else if constexpr (sizeof(T) == 4 && VEC == 4) {
        const float4 f = *reinterpret_cast<const float4*>(p);
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[j] = (float)p[j];
    }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ x, int ld, int HW, int C, int G, float eps,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float* __restrict__ scale, float* __restrict__ shift) {
    __shared__ float red[8];
    const int g = blockIdx.x, b = blockIdx.y, cpg = C / G;
    const T* xb = x + (size_t)b * HW * ld + g * cpg;
    const int vpr = cpg / VEC;                  // vectors per pixel
    const int nv = HW * vpr;
    const float n = (float)HW * (float)cpg;
    float s = 0.f;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        const int r = i / vpr, c = (i - r * vpr) * VEC;
        float v[VEC];
        gn_load<T, VEC>(xb + (size_t)r * ld + c, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) s += v[j];
    }
    const float mean = block_sum(s, red) / n;
    float q = 0.f;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        const int r = i / vpr, c = (i - r * vpr) * VEC;
        float v[VEC];
        gn_load<T, VEC>(xb + (size_t)r * ld + c, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) { const float d = v[j] - mean; q += d * d; }
    }
    const float var = block_sum(q, red) / n;
    const float rstd = rsqrtf(var + eps);
    if (threadIdx.x < cpg) {
        const int c = g * cpg + threadIdx.x;
        const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
        scale[(size_t)b * C + c] = rstd * ga;
        shift[(size_t)b * C + c] = be - mean * rstd * ga;
    }
}

template <typename T>
static void gn_stats_launch(const T* x, int ld, int B, int HW, int C, int G, float eps, const float* gamma, const float* beta,
                            float* scale, float* shift, cudaStream_t st) {
    const int cpg = C / G;
    const dim3 grid(G, B);
    const bool al = (((uintptr_t)x) % 16 == 0) && (ld * (int)sizeof(T)) % 16 == 0 && (cpg * (int)sizeof(T)) % 8 == 0;
    if (sizeof(T) == 2 && al && cpg % 8 == 0 && (cpg * 2) % 16 == 0)
        gn_stats_kernel<T, 8><<<grid, 256, 0, st>>>(x, ld, HW, C, G, eps, gamma, beta, scale, shift);
    else if (al && cpg % 4 == 0 && (sizeof(T) == 2 || (cpg * 4) % 16 == 0))
        gn_stats_kernel<T, 4><<<grid, 256, 0, st>>>(x, ld, HW, C, G, eps, gamma, beta, scale, shift);
    else
        gn_stats_kernel<T, 1><<<grid, 256, 0, st>>>(x, ld, HW, C, G, eps, gamma, beta, scale, shift);
}

// LayerNorm over C per token row: one warp per row.
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, int ldx, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, __half* __restrict__ out,
                                                        int ldo, long long rows, int C) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const __half* xr = x + row * ldx;
    float s = 0.f;
    for (int c = lane * 8; c < C; c += 256) {
        const Half8 h = *reinterpret_cast<const Half8*>(xr + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h.v[j]); s += f.x + f.y; }
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
    for (int c = lane * 8; c < C; c += 256) {
        const Half8 h = *reinterpret_cast<const Half8*>(xr + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h.v[j]);
            q += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    for (int c = lane * 8; c < C; c += 256) {
        const Half8 h = *reinterpret_cast<const Half8*>(xr + c);
        Half8 o;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h.v[j]);
            o.v[j] = __floats2half2_rn((f.x - mean) * rstd * gamma[c + 2 * j] + beta[c + 2 * j],
                                       (f.y - mean) * rstd * gamma[c + 2 * j + 1] + beta[c + 2 * j + 1]);
        }
        *reinterpret_cast<Half8*>(out + row * ldo + c) = o;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// small-head attention, one query row per thread
// ---------------------------------------------------------------------------------------------------------------------
template <int HDP>
struct RowState {
    float q[HDP], acc[HDP], m, l;
};

template <int HDP>
__device__ __forceinline__ void load_row(float (&dst)[HDP], const __half* __restrict__ p, float mul) {
#pragma unroll
    for (int c = 0; c < HDP; c += 8) {
        const Half8 h = *reinterpret_cast<const Half8*>(p + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h.v[j]);
            dst[c + 2 * j] = f.x * mul;
            dst[c + 2 * j + 1] = f.y * mul;
        }
    }
}

// K / V tiles live in shared memory as fp32 [keys][HDP] (converted once per tile): the inner loops then read whole rows with
// 16-byte broadcast loads and contain no conversions.  (First version: fp16 tiles read as half2 - 8 LDS + 8 cvt per 16 FMA at
// head_dim 8 - ran at 0.6 T scores/s, LSU / conversion bound; profiles/r01_launch_roofline_moa_mot_n.txt.)
// consume `nk` keys (multiple of 8 slots; keys >= nvalid are masked) for QPT query rows held by this thread
template <int HDP, int QPT>
__device__ __forceinline__ void attend_tile(RowState<HDP> (&st)[QPT], const float* __restrict__ sk, const float* __restrict__ sv,
                                            int nk, int nvalid) {
    for (int j0 = 0; j0 < nk && j0 < nvalid; j0 += 8) {
        float s[QPT][8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            float kr[HDP];
#pragma unroll
            for (int d = 0; d < HDP; d += 4) {
                const float4 f = *reinterpret_cast<const float4*>(sk + (j0 + jj) * HDP + d);
                kr[d] = f.x; kr[d + 1] = f.y; kr[d + 2] = f.z; kr[d + 3] = f.w;
            }
            const bool ok = j0 + jj < nvalid;
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) {
                float a = 0.f;
#pragma unroll
                for (int d = 0; d < HDP; ++d) a = fmaf(st[qq].q[d], kr[d], a);
                s[qq][jj] = ok ? a : -INFINITY;
            }
        }
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) {
            float mx = s[qq][0];
#pragma unroll
            for (int jj = 1; jj < 8; ++jj) mx = fmaxf(mx, s[qq][jj]);
            const float mn = fmaxf(st[qq].m, mx);
            const float corr = exp2f(st[qq].m - mn);   // m == -inf on the first group: exp2(-inf) = 0
            st[qq].m = mn;
            st[qq].l *= corr;
#pragma unroll
            for (int d = 0; d < HDP; ++d) st[qq].acc[d] *= corr;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                s[qq][jj] = exp2f(s[qq][jj] - mn);
                st[qq].l += s[qq][jj];
            }
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            float vr[HDP];
#pragma unroll
            for (int d = 0; d < HDP; d += 4) {
                const float4 f = *reinterpret_cast<const float4*>(sv + (j0 + jj) * HDP + d);
                vr[d] = f.x; vr[d + 1] = f.y; vr[d + 2] = f.z; vr[d + 3] = f.w;
            }
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq)
#pragma unroll
                for (int d = 0; d < HDP; ++d) st[qq].acc[d] = fmaf(s[qq][jj], vr[d], st[qq].acc[d]);
        }
    }
}

template <int HDP>
__device__ __forceinline__ void init_row(RowState<HDP>& st) {
    st.m = -INFINITY;
    st.l = 0.f;
#pragma unroll
    for (int d = 0; d < HDP; ++d) { st.acc[d] = 0.f; st.q[d] = 0.f; }
}

template <int HDP>
__device__ __forceinline__ void store_row(const RowState<HDP>& st, __half* __restrict__ o) {
    const float inv = 1.f / st.l;
#pragma unroll
    for (int c = 0; c < HDP; c += 8) {
        Half8 h;
#pragma unroll
        for (int j = 0; j < 4; ++j) h.v[j] = __floats2half2_rn(st.acc[c + 2 * j] * inv, st.acc[c + 2 * j + 1] * inv);
        *reinterpret_cast<Half8*>(o + c) = h;
    }
}

// 8 fp16 channels -> 8 fp32 in a shared-memory row
__device__ __forceinline__ void stage8(float* __restrict__ dst, const Half8& h) {
    const float2 f0 = __half22float2(h.v[0]), f1 = __half22float2(h.v[1]), f2 = __half22float2(h.v[2]), f3 = __half22float2(h.v[3]);
    *reinterpret_cast<float4*>(dst) = make_float4(f0.x, f0.y, f1.x, f1.y);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(f2.x, f2.y, f3.x, f3.y);
}

constexpr int AT_TK = 64;        // keys per shared-memory tile
constexpr int AT_THREADS = 128;

// Global attention: q rows [batch*Nq], k/v rows [batch*Nkv]; head h uses channels [h*HDP, (h+1)*HDP) of each pointer.
// Each thread owns QPT query rows (qi, qi + 128, ...): a K / V row read from shared memory is used QPT times.
template <int HDP, int QPT>
__global__ void __launch_bounds__(AT_THREADS) attn_small_kernel(const __half* __restrict__ q, int ldq, const __half* __restrict__ k,
                                                                int ldk, const __half* __restrict__ v, int ldv, int Nq, int Nkv,
                                                                float scale_log2, __half* __restrict__ out, int ldo) {
    __shared__ __align__(16) float sk[AT_TK * HDP];
    __shared__ __align__(16) float sv[AT_TK * HDP];
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * (AT_THREADS * QPT) + threadIdx.x;
    RowState<HDP> st[QPT];
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) {
        init_row<HDP>(st[qq]);
        const int qi = q0 + qq * AT_THREADS;
        load_row<HDP>(st[qq].q, q + ((size_t)b * Nq + (qi < Nq ? qi : 0)) * ldq + h * HDP, scale_log2);
    }
    constexpr int VPR = HDP / 8;   // 16-byte vectors per row
    for (int k0 = 0; k0 < Nkv; k0 += AT_TK) {
        __syncthreads();
        for (int i = threadIdx.x; i < AT_TK * VPR; i += blockDim.x) {
            const int r = i / VPR, part = i - r * VPR;
            const int kr = k0 + r;
            Half8 hk, hv;
            if (kr < Nkv) {
                hk = *reinterpret_cast<const Half8*>(k + ((size_t)b * Nkv + kr) * ldk + h * HDP + part * 8);
                hv = *reinterpret_cast<const Half8*>(v + ((size_t)b * Nkv + kr) * ldv + h * HDP + part * 8);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) hk.v[j] = hv.v[j] = __floats2half2_rn(0.f, 0.f);
            }
            stage8(sk + r * HDP + part * 8, hk);
            stage8(sv + r * HDP + part * 8, hv);
        }
        __syncthreads();
        attend_tile<HDP, QPT>(st, sk, sv, AT_TK, Nkv - k0);
    }
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) {
        const int qi = q0 + qq * AT_THREADS;
        if (qi < Nq) store_row<HDP>(st[qq], out + ((size_t)b * Nq + qi) * ldo + h * HDP);
    }
}

// Window attention (win*win <= 64 tokens per window) with the reference's pad-then-roll token mapping:
// window (wy,wx) token (iy,ix) -> padded-rolled position (py,px) -> source ((py+shift)%Hp, (px+shift)%Wp); sources outside
// HxW are padding tokens whose k/v are `padk` / `padv` (per head layout) or zeros.
template <int HDP>
__global__ void __launch_bounds__(64) attn_window_kernel(const __half* __restrict__ q, int ldq, const __half* __restrict__ k,
                                                         int ldk, const __half* __restrict__ v, int ldv, int H, int W, int win,
                                                         int shift, const __half* __restrict__ padq,
                                                         const __half* __restrict__ padk, const __half* __restrict__ padv,
                                                         float scale_log2, __half* __restrict__ out, int ldo) {
    __shared__ __align__(16) float sk[64 * HDP];
    __shared__ __align__(16) float sv[64 * HDP];
    const int Hp = (H + win - 1) / win * win, Wp = (W + win - 1) / win * win;
    const int nwx = Wp / win;
    const int wy = blockIdx.x / nwx, wx = blockIdx.x - wy * nwx;
    const int h = blockIdx.y, b = blockIdx.z;
    const int t = threadIdx.x, WT = win * win;
    const int iy = t / win, ix = t - iy * win;
    int sy = wy * win + iy + shift, sx = wx * win + ix + shift;
    if (sy >= Hp) sy -= Hp;
    if (sx >= Wp) sx -= Wp;
    const bool tok = t < WT;
    const bool real = tok && sy < H && sx < W;
    const size_t row = (size_t)b * H * W + (size_t)sy * W + sx;
    RowState<HDP> st[1];
    init_row<HDP>(st[0]);
    constexpr int VPR = HDP / 8;
    {
        const __half* kp = real ? k + row * ldk + h * HDP : (padk ? padk + h * HDP : nullptr);
        const __half* vp = real ? v + row * ldv + h * HDP : (padv ? padv + h * HDP : nullptr);
#pragma unroll
        for (int part = 0; part < VPR; ++part) {
            Half8 hk, hv;
#pragma unroll
            for (int j = 0; j < 4; ++j) hk.v[j] = hv.v[j] = __floats2half2_rn(0.f, 0.f);
            if (tok && kp) hk = *reinterpret_cast<const Half8*>(kp + part * 8);
            if (tok && vp) hv = *reinterpret_cast<const Half8*>(vp + part * 8);
            stage8(sk + t * HDP + part * 8, hk);
            stage8(sv + t * HDP + part * 8, hv);
        }
        if (real) load_row<HDP>(st[0].q, q + row * ldq + h * HDP, scale_log2);
    }
    (void)padq;   // padding queries produce rows that the reference crops away: never computed
    __syncthreads();
    if (!real) return;
    attend_tile<HDP, 1>(st, sk, sv, 64, WT);
    store_row<HDP>(st[0], out + row * ldo + h * HDP);
}

// ---------------------------------------------------------------------------------------------------------------------
// Deformable sampling (mot/experts.py:416-475): per (token, head): 4 points, offsets tanh*0.25 around the token's own
// normalised position, softmax over points, bilinear zero-padded align_corners sampling of V.
// oa: fp32 [rows, nh*np*3] = [offsets (h,p,2) | attention logits (h,p)].
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) deform_sample_kernel(const float* __restrict__ oa, int ldoa, const __half* __restrict__ v,
                                                            int ldv, int B, int H, int W, int nh, int hd, int np,
                                                            int align_corners, __half* __restrict__ out, int ldo) {
    const int cpt = hd >> 3;   // 8-channel chunks per head
    const long long total = (long long)B * H * W * nh * cpt;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int chunk = (int)(i % cpt);
    const int h = (int)((i / cpt) % nh);
    const long long row = i / ((long long)cpt * nh);
    const int n = (int)(row % ((long long)H * W));
    const long long b = row / ((long long)H * W);
    const int ty = n / W, tx = n - ty * W;
    const float refx = (float)tx / (float)max(W - 1, 1) * 2.f - 1.f;
    const float refy = (float)ty / (float)max(H - 1, 1) * 2.f - 1.f;
    const float* offp = oa + row * ldoa + (size_t)h * np * 2;
    const float* awp = oa + row * ldoa + (size_t)nh * np * 2 + (size_t)h * np;
    float mx = -INFINITY;
    for (int p = 0; p < np; ++p) mx = fmaxf(mx, awp[p]);
    float den = 0.f;
    for (int p = 0; p < np; ++p) den += expf(awp[p] - mx);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const __half* vb = v + (size_t)b * H * W * ldv + h * hd + chunk * 8;
    for (int p = 0; p < np; ++p) {
        const float wgt = expf(awp[p] - mx) / den;
        const float gx = fminf(fmaxf(refx + tanhf(offp[2 * p]) * 0.25f, -1.f), 1.f);
        const float gy = fminf(fmaxf(refy + tanhf(offp[2 * p + 1]) * 0.25f, -1.f), 1.f);
        float fx, fy;
        if (align_corners) {
            fx = (gx + 1.f) * 0.5f * (float)(W - 1);
            fy = (gy + 1.f) * 0.5f * (float)(H - 1);
        } else {
            fx = ((gx + 1.f) * (float)W - 1.f) * 0.5f;
            fy = ((gy + 1.f) * (float)H - 1.f) * 0.5f;
        }
        const float x0f = floorf(fx), y0f = floorf(fy);
        const int x0 = (int)x0f, y0 = (int)y0f;
        const float ax = fx - x0f, ay = fy - y0f;
#pragma unroll
        for (int cy = 0; cy < 2; ++cy) {
#pragma unroll
            for (int cx = 0; cx < 2; ++cx) {
                const int xx = x0 + cx, yy = y0 + cy;
                if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
                const float cw = wgt * (cx ? ax : 1.f - ax) * (cy ? ay : 1.f - ay);
                const Half8 hv = *reinterpret_cast<const Half8*>(vb + ((size_t)yy * W + xx) * ldv);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(hv.v[j]);
                    acc[2 * j] = fmaf(cw, f.x, acc[2 * j]);
                    acc[2 * j + 1] = fmaf(cw, f.y, acc[2 * j + 1]);
                }
            }
        }
    }
    Half8 o;
#pragma unroll
    for (int j = 0; j < 4; ++j) o.v[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
    *reinterpret_cast<Half8*>(out + row * ldo + h * hd + chunk * 8) = o;
}

// ---------------------------------------------------------------------------------------------------------------------
// Per-token routers (mot/router.py:211-291, moa/router.py:50-62), fp32 throughout.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int RT_MAXH = 32;

// hidden[row, j] = sum_c w1[j][c] * x[row, c]          (1x1 conv, no bias)
__global__ void __launch_bounds__(128) token_router_hidden_kernel(const __half* __restrict__ x, int ldx, const float* __restrict__ w1,
                                                                  int C, int HID, float* __restrict__ hidden, long long rows) {
    extern __shared__ float sw[];   // [HID][C]
    for (int i = threadIdx.x; i < HID * C; i += blockDim.x) sw[i] = w1[i];
    __syncthreads();
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    float acc[RT_MAXH];
#pragma unroll
    for (int j = 0; j < RT_MAXH; ++j) acc[j] = 0.f;
    const __half* xr = x + row * ldx;
    for (int c = 0; c < C; c += 8) {
        const Half8 h = *reinterpret_cast<const Half8*>(xr + c);
        float xv[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h.v[j]); xv[2 * j] = f.x; xv[2 * j + 1] = f.y; }
#pragma unroll
        for (int j = 0; j < RT_MAXH; ++j) {
            if (j < HID) {
                const float* wr = sw + j * C + c;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[j] = fmaf(xv[e], wr[e], acc[j]);
            }
        }
    }
    for (int j = 0; j < HID; ++j) hidden[row * HID + j] = acc[j];
}

// logits = w2 . silu(hidden*sc + sh) + b2;  probs = softmax(logits / T);  top-k (k < E) + renormalise, scattered dense.
__global__ void __launch_bounds__(128) token_router_finish_kernel(const float* __restrict__ hidden, const float* __restrict__ sc,
                                                                  const float* __restrict__ sh, const float* __restrict__ w2,
                                                                  const float* __restrict__ b2, int HID, int E, int topk,
                                                                  const float* __restrict__ temp_ptr, float temp, int HW,
                                                                  float* __restrict__ weights, int* __restrict__ idx_out,
                                                                  long long rows) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const long long b = row / HW;
    float T = temp_ptr ? *temp_ptr : temp;
    float lg[4] = {0.f, 0.f, 0.f, 0.f};
    for (int e = 0; e < E; ++e) lg[e] = b2[e];
    for (int j = 0; j < HID; ++j) {
        const float t = fmaf(hidden[row * HID + j], sc[b * HID + j], sh[b * HID + j]);
        const float a = t / (1.f + expf(-t));
        for (int e = 0; e < E; ++e) lg[e] = fmaf(w2[e * HID + j], a, lg[e]);
    }
    float mx = -INFINITY;
    for (int e = 0; e < E; ++e) { lg[e] = lg[e] / T; mx = fmaxf(mx, lg[e]); }
    float den = 0.f, p[4] = {0.f, 0.f, 0.f, 0.f};
    for (int e = 0; e < E; ++e) { p[e] = expf(lg[e] - mx); den += p[e]; }
    for (int e = 0; e < E; ++e) p[e] /= den;
    if (topk < E) {
        float outw[4] = {0.f, 0.f, 0.f, 0.f};
        bool used[4] = {false, false, false, false};
        int sel[4];
        float sum = 0.f;
        for (int r = 0; r < topk; ++r) {
            int best = -1;
            for (int e = 0; e < E; ++e)
                if (!used[e] && (best < 0 || p[e] > p[best])) best = e;   // ties -> lower index
            used[best] = true;
            sel[r] = best;
            sum += p[best];
        }
        sum = fmaxf(sum, 1e-6f);
        for (int r = 0; r < topk; ++r) {
            outw[sel[r]] = p[sel[r]] / sum;
            if (idx_out) idx_out[row * topk + r] = sel[r];
        }
        for (int e = 0; e < E; ++e) weights[row * E + e] = outw[e];
    } else {
        for (int e = 0; e < E; ++e) {
            weights[row * E + e] = p[e];
            if (idx_out) idx_out[row * E + e] = e;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Performer-style linear attention of the MoA global head (moa/heads.py:318-352), fp32.
//   phase 1: per (image, head, token chunk): partial kv[nb][hd] = sum_n kf[n][i]*v[n][d], ksum[nb] = sum_n kf[n][i]
//   phase 2: per token: out = clamp(qf.kv) / max(qf.ksum, eps), partials summed in fixed order
// ---------------------------------------------------------------------------------------------------------------------
constexpr int LA_CHUNK = 128;

template <int HDP>
__global__ void __launch_bounds__(LA_CHUNK) linattn_reduce_kernel(const __half* __restrict__ k, int ldk, const __half* __restrict__ v,
                                                                  int ldv, const float* __restrict__ rf, int nb, int hd, int N,
                                                                  float feat_scale, float eps, float limit,
                                                                  float* __restrict__ partial) {
    __shared__ float skf[LA_CHUNK][HDP + 1];
    __shared__ float svv[LA_CHUNK][HDP + 1];
    __shared__ float srf[HDP * HDP];
    const int chunk = blockIdx.x, h = blockIdx.y, b = blockIdx.z, nchunks = gridDim.x;
    for (int i = threadIdx.x; i < nb * hd; i += blockDim.x) srf[i] = rf[i];
    __syncthreads();
    const int n = chunk * LA_CHUNK + threadIdx.x;
    float kr[HDP], vr[HDP];
    if (n < N) {
        load_row<HDP>(kr, k + ((size_t)b * N + n) * ldk + h * HDP, 1.f);
        load_row<HDP>(vr, v + ((size_t)b * N + n) * ldv + h * HDP, 1.f);
    }
#pragma unroll
    for (int i = 0; i < HDP; ++i) {
        float f = 0.f;
        if (n < N && i < nb) {
#pragma unroll
            for (int d = 0; d < HDP; ++d)
                if (d < hd) f = fmaf(kr[d], srf[i * hd + d], f);
            f = fminf(fmaxf(f * feat_scale, 0.f) + eps, limit);
        }
        skf[threadIdx.x][i] = f;
        svv[threadIdx.x][i] = (n < N && i < hd) ? vr[i] : 0.f;
    }
    __syncthreads();
    float* dst = partial + (((size_t)b * gridDim.y + h) * nchunks + chunk) * (size_t)(HDP * (HDP + 1));
    for (int e = threadIdx.x; e < HDP * (HDP + 1); e += blockDim.x) {
        const int i = e / (HDP + 1), d = e - i * (HDP + 1);
        float a = 0.f;
        if (d < HDP) {
            for (int t = 0; t < LA_CHUNK; ++t) a = fmaf(skf[t][i], svv[t][d], a);
        } else {
            for (int t = 0; t < LA_CHUNK; ++t) a += skf[t][i];
        }
        dst[e] = a;
    }
}

template <int HDP>
__global__ void __launch_bounds__(LA_CHUNK) linattn_apply_kernel(const __half* __restrict__ q, int ldq, const float* __restrict__ rf,
                                                                 int nb, int hd, int N, int nchunks, float feat_scale, float eps,
                                                                 float limit, const float* __restrict__ partial,
                                                                 __half* __restrict__ out, int ldo) {
    __shared__ float skv[HDP * (HDP + 1)];
    __shared__ float srf[HDP * HDP];
    const int h = blockIdx.y, b = blockIdx.z;
    const float* src = partial + (((size_t)b * gridDim.y + h) * nchunks) * (size_t)(HDP * (HDP + 1));
    for (int e = threadIdx.x; e < HDP * (HDP + 1); e += blockDim.x) {
        float a = 0.f;
        for (int c = 0; c < nchunks; ++c) a += src[(size_t)c * (HDP * (HDP + 1)) + e];
        skv[e] = a;
    }
    for (int i = threadIdx.x; i < nb * hd; i += blockDim.x) srf[i] = rf[i];
    __syncthreads();
    const int n = blockIdx.x * LA_CHUNK + threadIdx.x;
    if (n >= N) return;
    float qr[HDP];
    load_row<HDP>(qr, q + ((size_t)b * N + n) * ldq + h * HDP, 1.f);
    float num[HDP];
#pragma unroll
    for (int d = 0; d < HDP; ++d) num[d] = 0.f;
    float den = 0.f;
#pragma unroll
    for (int i = 0; i < HDP; ++i) {
        if (i < nb) {
            float f = 0.f;
#pragma unroll
            for (int d = 0; d < HDP; ++d)
                if (d < hd) f = fmaf(qr[d], srf[i * hd + d], f);
            f = fminf(fmaxf(f * feat_scale, 0.f) + eps, limit);
#pragma unroll
            for (int d = 0; d < HDP; ++d) num[d] = fmaf(f, skv[i * (HDP + 1) + d], num[d]);
            den = fmaf(f, skv[i * (HDP + 1) + HDP], den);
        }
    }
    den = fmaxf(den, eps);
    __half* o = out + ((size_t)b * N + n) * ldo + h * HDP;
#pragma unroll
    for (int c = 0; c < HDP; c += 8) {
        Half8 hv;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a0 = fminf(fmaxf(num[c + 2 * j], -limit), limit) / den;
            const float a1 = fminf(fmaxf(num[c + 2 * j + 1], -limit), limit) / den;
            hv.v[j] = __floats2half2_rn(c + 2 * j < hd ? a0 : 0.f, c + 2 * j + 1 < hd ? a1 : 0.f);
        }
        *reinterpret_cast<Half8*>(o + c) = hv;
    }
}

// adaptive average pooling (F.adaptive_avg_pool2d bin edges: floor(i*H/h) .. ceil((i+1)*H/h))
__global__ void __launch_bounds__(256) adaptive_avgpool_kernel(const __half* __restrict__ x, int ldx, int B, int H, int W, int C,
                                                               int h, int w, __half* __restrict__ out, int ldo) {
    const int cv = C >> 3;
    const long long total = (long long)B * h * w * cv;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % cv) << 3;
    const int ox = (int)((i / cv) % w), oy = (int)((i / ((long long)cv * w)) % h);
    const long long b = i / ((long long)cv * w * h);
    const int y0 = (oy * H) / h, y1 = ((oy + 1) * H + h - 1) / h;
    const int x0 = (ox * W) / w, x1 = ((ox + 1) * W + w - 1) / w;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int yy = y0; yy < y1; ++yy)
        for (int xx = x0; xx < x1; ++xx) {
            const Half8 hv = *reinterpret_cast<const Half8*>(x + ((size_t)(b * H + yy) * W + xx) * ldx + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(hv.v[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
        }
    const float inv = 1.f / (float)((y1 - y0) * (x1 - x0));
    Half8 o;
#pragma unroll
    for (int j = 0; j < 4; ++j) o.v[j] = __floats2half2_rn(acc[2 * j] * inv, acc[2 * j + 1] * inv);
    *reinterpret_cast<Half8*>(out + ((size_t)(b * h + oy) * w + ox) * ldo + c) = o;
}

}  // namespace ym

using namespace ym;

#define LOG2E 1.4426950408889634f

static inline bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

extern "C" int ym_groupnorm_stats(const void* x, int x_f32, int ld, int B, int HW, int C, int G, float eps, const float* gamma,
                                  const float* beta, float* scale, float* shift, void* stream) {
    YM_CHECK_ARG(x && scale && shift, "ym_groupnorm_stats: null pointer");
    YM_CHECK_ARG(G >= 1 && C % G == 0 && C / G <= 256, "ym_groupnorm_stats: C=%d must be divisible by G=%d (<=256 channels per group)", C, G);
    if (B == 0) return YM_OK;
    if (x_f32) gn_stats_launch<float>((const float*)x, ld, B, HW, C, G, eps, gamma, beta, scale, shift, (cudaStream_t)stream);
    else gn_stats_launch<__half>((const __half*)x, ld, B, HW, C, G, eps, gamma, beta, scale, shift, (cudaStream_t)stream);
    YM_CHECK_LAUNCH("groupnorm_stats");
    return YM_OK;
}

extern "C" int ym_layernorm_nhwc(const void* x, int ldx, const float* gamma, const float* beta, float eps, void* out, int ldo,
                                 long long rows, int C, void* stream) {
    YM_CHECK_ARG(x && gamma && beta && out, "ym_layernorm_nhwc: null pointer");
    YM_CHECK_ARG(C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && al16(x) && al16(out), "ym_layernorm_nhwc: multiples of 8 / 16-byte alignment");
    if (rows == 0) return YM_OK;
    layernorm_kernel<<<(int)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const __half*)x, ldx, gamma, beta, eps, (__half*)out,
                                                                              ldo, rows, C);
    YM_CHECK_LAUNCH("layernorm");
    return YM_OK;
}

#define ATTN_ARGS_OK(name)                                                                                                   \
    YM_CHECK_ARG(q && k && v && out, name ": null pointer");                                                                 \
    YM_CHECK_ARG(hdp == 8 || hdp == 16 || hdp == 24 || hdp == 32, name ": (padded) head_dim must be 8/16/24/32 (got %d)", hdp); \
    YM_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && al16(q) && al16(k) && al16(v) && al16(out),   \
                 name ": pitches must be multiples of 8 and pointers 16-byte aligned")

extern "C" int ym_attn_small(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int batch, int heads, int hdp,
                             int Nq, int Nkv, float scale, void* out, int ldo, void* stream) {
    ATTN_ARGS_OK("ym_attn_small");
    YM_CHECK_ARG(Nq >= 1 && Nkv >= 1 && heads >= 1, "ym_attn_small: empty problem");
    if (batch == 0) return YM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const float sl = scale * LOG2E;
    // two query rows per thread where the register budget allows it (head_dim <= 16) and the problem is large enough to
    // still fill the GPU with 256-query CTAs
    const bool two = hdp <= 16 && (long long)((Nq + 255) / 256) * heads * batch >= 2 * 148;
    const int qpc = two ? 256 : 128;
    dim3 grid((Nq + qpc - 1) / qpc, heads, batch);
#define AS_LAUNCH(H, Q)                                                                                                      \
    attn_small_kernel<H, Q><<<grid, AT_THREADS, 0, st>>>((const __half*)q, ldq, (const __half*)k, ldk, (const __half*)v, ldv, Nq, \
                                                         Nkv, sl, (__half*)out, ldo)
    switch (hdp) {
        case 8: if (two) AS_LAUNCH(8, 2); else AS_LAUNCH(8, 1); break;
        case 16: if (two) AS_LAUNCH(16, 2); else AS_LAUNCH(16, 1); break;
        case 24: AS_LAUNCH(24, 1); break;
        default: AS_LAUNCH(32, 1); break;
    }
#undef AS_LAUNCH
    YM_CHECK_LAUNCH("attn_small");
    return YM_OK;
}

extern "C" int ym_attn_window(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int B, int H, int W, int heads,
                              int hdp, int win, int shift, const void* padq, const void* padk, const void* padv, float scale,
                              void* out, int ldo, void* stream) {
    ATTN_ARGS_OK("ym_attn_window");
    YM_CHECK_ARG(win >= 1 && win <= 8, "ym_attn_window: window size must be in 1..8 (got %d)", win);
    YM_CHECK_ARG(shift >= 0 && shift < win, "ym_attn_window: shift must be in [0, win)");
    if (B == 0) return YM_OK;
    const int nwy = (H + win - 1) / win, nwx = (W + win - 1) / win;
    dim3 grid(nwy * nwx, heads, B);
    cudaStream_t st = (cudaStream_t)stream;
    const float sl = scale * LOG2E;
#define AW_LAUNCH(HD)                                                                                                        \
    attn_window_kernel<HD><<<grid, 64, 0, st>>>((const __half*)q, ldq, (const __half*)k, ldk, (const __half*)v, ldv, H, W, win,  \
                                                shift, (const __half*)padq, (const __half*)padk, (const __half*)padv, sl,        \
                                                (__half*)out, ldo)
    switch (hdp) {
        case 8: AW_LAUNCH(8); break;
        case 16: AW_LAUNCH(16); break;
        case 24: AW_LAUNCH(24); break;
        default: AW_LAUNCH(32); break;
    }
#undef AW_LAUNCH
    YM_CHECK_LAUNCH("attn_window");
    return YM_OK;
}

extern "C" int ym_deform_sample(const float* oa, int ldoa, const void* v, int ldv, int B, int H, int W, int heads, int hd, int np,
                                int align_corners, void* out, int ldo, void* stream) {
    YM_CHECK_ARG(oa && v && out, "ym_deform_sample: null pointer");
    YM_CHECK_ARG(hd % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && al16(v) && al16(out), "ym_deform_sample: head_dim / pitches must be multiples of 8");
    YM_CHECK_ARG(np >= 1 && np <= 16 && ldoa >= heads * np * 3, "ym_deform_sample: bad n_points / offset pitch");
    if (B == 0) return YM_OK;
    const long long total = (long long)B * H * W * heads * (hd / 8);
    deform_sample_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(oa, ldoa, (const __half*)v, ldv, B, H, W, heads,
                                                                                       hd, np, align_corners, (__half*)out, ldo);
    YM_CHECK_LAUNCH("deform_sample");
    return YM_OK;
}

// scratch: fp32 [rows*HID] hidden + [2*B*HID] GroupNorm affine
extern "C" long long ym_token_router_scratch_floats(int B, int HW, int HID) { return (long long)B * HW * HID + 2LL * B * HID; }

extern "C" int ym_token_router(const void* x, int ldx, int B, int HW, int C, const float* w1, int HID, int G, const float* gn_w,
                               const float* gn_b, float gn_eps, const float* w2, const float* b2, int E, int topk,
                               const float* temp_dev, float temp, float* weights, int* idx, float* scratch, void* stream) {
    YM_CHECK_ARG(x && w1 && w2 && b2 && weights && scratch, "ym_token_router: null pointer");
    YM_CHECK_ARG(C % 8 == 0 && ldx % 8 == 0 && al16(x), "ym_token_router: C / pitch must be multiples of 8");
    YM_CHECK_ARG(HID >= 1 && HID <= RT_MAXH && E >= 1 && E <= 4 && topk >= 1 && topk <= E, "ym_token_router: HID<=32, E<=4, 1<=topk<=E");
    YM_CHECK_ARG(G >= 1 && HID % G == 0, "ym_token_router: GroupNorm groups must divide the hidden width");
    YM_CHECK_ARG((size_t)HID * C * 4 <= 48 * 1024, "ym_token_router: router weights exceed 48 KB of shared memory");
    if (B == 0) return YM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = (long long)B * HW;
    float* hidden = scratch;
    float* sc = scratch + rows * HID;
    float* sh = sc + (size_t)B * HID;
    token_router_hidden_kernel<<<(int)((rows + 127) / 128), 128, (size_t)HID * C * 4, st>>>((const __half*)x, ldx, w1, C, HID, hidden, rows);
    YM_CHECK_LAUNCH("token_router_hidden");
    gn_stats_launch<float>(hidden, HID, B, HW, HID, G, gn_eps, gn_w, gn_b, sc, sh, st);
    YM_CHECK_LAUNCH("token_router_gn");
    token_router_finish_kernel<<<(int)((rows + 127) / 128), 128, 0, st>>>(hidden, sc, sh, w2, b2, HID, E, topk, temp_dev, temp, HW,
                                                                           weights, idx, rows);
    YM_CHECK_LAUNCH("token_router_finish");
    return YM_OK;
}

extern "C" long long ym_linear_attn_scratch_floats(int batch, int heads, int hdp, int N) {
    return (long long)batch * heads * ((N + LA_CHUNK - 1) / LA_CHUNK) * hdp * (hdp + 1);
}

extern "C" int ym_linear_attn(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int batch, int heads, int hdp,
                              int hd, int nb, int N, const float* rf, float eps, float limit, float* scratch, void* out, int ldo,
                              void* stream) {
    ATTN_ARGS_OK("ym_linear_attn");
    YM_CHECK_ARG(rf && scratch, "ym_linear_attn: null pointer");
    YM_CHECK_ARG(hd >= 1 && hd <= hdp && nb >= 1 && nb <= hdp, "ym_linear_attn: hd / nb must be <= padded head_dim");
    if (batch == 0) return YM_OK;
    const int nchunks = (N + LA_CHUNK - 1) / LA_CHUNK;
    dim3 grid(nchunks, heads, batch);
    cudaStream_t st = (cudaStream_t)stream;
    const float fs = 1.f / sqrtf((float)nb);
#define LA_LAUNCH(H)                                                                                                            \
    do {                                                                                                                        \
        linattn_reduce_kernel<H><<<grid, LA_CHUNK, 0, st>>>((const __half*)k, ldk, (const __half*)v, ldv, rf, nb, hd, N, fs, eps, limit, \
                                                            scratch);                                                           \
        linattn_apply_kernel<H><<<grid, LA_CHUNK, 0, st>>>((const __half*)q, ldq, rf, nb, hd, N, nchunks, fs, eps, limit, scratch,      \
                                                           (__half*)out, ldo);                                                  \
    } while (0)
    switch (hdp) {
        case 8: LA_LAUNCH(8); break;
        case 16: LA_LAUNCH(16); break;
        case 24: LA_LAUNCH(24); break;
        default: LA_LAUNCH(32); break;
    }
#undef LA_LAUNCH
    YM_CHECK_LAUNCH("linear_attn");
    return YM_OK;
}

extern "C" int ym_adaptive_avgpool_nhwc(const void* x, int ldx, int B, int H, int W, int C, int h, int w, void* out, int ldo,
                                        void* stream) {
    YM_CHECK_ARG(x && out, "ym_adaptive_avgpool_nhwc: null pointer");
    YM_CHECK_ARG(C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && al16(x) && al16(out), "ym_adaptive_avgpool_nhwc: multiples of 8");
    YM_CHECK_ARG(h >= 1 && w >= 1 && h <= H && w <= W, "ym_adaptive_avgpool_nhwc: bad output size");
    if (B == 0) return YM_OK;
    const long long total = (long long)B * h * w * (C / 8);
    adaptive_avgpool_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __half*)x, ldx, B, H, W, C, h, w,
                                                                                          (__half*)out, ldo);
    YM_CHECK_LAUNCH("adaptive_avgpool");
    return YM_OK;
}
