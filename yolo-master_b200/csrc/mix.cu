// Element-wise glue of the mixture blocks (HBM-bound, 16-byte vectors, NHWC fp16 with row pitches).
//   ym_ew_nhwc: layer-scale residuals, token-weighted expert accumulation, GLU gate, GELU, per-(image,channel) affine,
//               sigmoid / multiplicative gates of the gated MoE family.
#include "ym_common.cuh"

namespace ym {

enum EwOp {
    EW_SCALE_RES = 0,   // out = a + chan[c] * b                       (A2C2f gamma block.py:1879; ls1/ls2, ls_attn/ls_ffn)
    EW_TOKEN_ACC = 1,   // out = (a ? a : 0) + tok[row*ldt + toff] * b (MoT blend mot/block.py:347-364, MoA mix moa/block.py:232-244)
    EW_GLU = 2,         // out = sigmoid(a) * b                        (mot/experts.py:168)
    EW_GELU = 3,        // out = gelu(a) (exact, erf)                  (nn.GELU in mot/experts.py:225,365)
    EW_AFFINE = 4,      // out = tok * [silu](a * sc[img,c] + sh[img,c]) + b   (GroupNorm apply; img = row / rows_per_img)
    EW_LERP = 5,        // out = t*a + (1-t)*b, t = p0[0]              (moa/heads.py:371-375)
    EW_SIGMOID = 6,     // out = sigmoid(a)                            (gate heads of moe/gated.py:1165-1172,1206-1208)
    EW_MUL_GATE = 7,    // out = a * (1 + t*b), t = p0[0]              (VisualDetailGate gated.py:1176-1178, t = tanh(detail_scale))
    EW_MUL = 8,         // out = a * b                                 (context * gate, gated.py:1219-1221)
};

__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); }

template <int OP>
__global__ void __launch_bounds__(256) ew_kernel(const __half* __restrict__ a, int lda, const __half* __restrict__ b, int ldb,
                                                 const float* __restrict__ p0, const float* __restrict__ p1,
                                                 const float* __restrict__ tok, int ldt, int toff,
                                                 int rows_per_img, int act, __half* __restrict__ out, int ldo, long long rows,
                                                 int C) {
    const int cv = C >> 3;
    const long long total = rows * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / cv;
        const int c = (int)(i - row * cv) << 3;
        float va[8], vb[8];
        if (a) {
            const Half8 h = *reinterpret_cast<const Half8*>(a + row * lda + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h.v[j]); va[2 * j] = f.x; va[2 * j + 1] = f.y; }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) va[j] = 0.f;
        }
        if (b) {
            const Half8 h = *reinterpret_cast<const Half8*>(b + row * ldb + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h.v[j]); vb[2 * j] = f.x; vb[2 * j + 1] = f.y; }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) vb[j] = 0.f;
        }
        float r[8];
        if (OP == EW_SCALE_RES) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = fmaf(p0[c + j], vb[j], va[j]);
        } else if (OP == EW_TOKEN_ACC) {
            const float w = tok[row * ldt + toff];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = fmaf(w, vb[j], va[j]);
        } else if (OP == EW_GLU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = vb[j] / (1.f + __expf(-va[j]));
        } else if (OP == EW_GELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = gelu_f(va[j]);
        } else if (OP == EW_LERP) {
            const float t = p0[0];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = t * va[j] + (1.f - t) * vb[j];
        } else if (OP == EW_SIGMOID) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = 1.f / (1.f + __expf(-va[j]));
        } else if (OP == EW_MUL_GATE) {
            const float t = p0[0];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = va[j] * fmaf(t, vb[j], 1.f);
        } else if (OP == EW_MUL) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = va[j] * vb[j];
        } else {
            const long long img = row / rows_per_img;
            const float w = tok ? tok[row * ldt + toff] : 1.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float v = fmaf(va[j], p0[img * C + c + j], p1[img * C + c + j]);
                if (act) v = silu_f(v);
                r[j] = fmaf(w, v, vb[j]);
            }
        }
        Half8 o;
#pragma unroll
        for (int j = 0; j < 4; ++j) o.v[j] = __floats2half2_rn(r[2 * j], r[2 * j + 1]);
        *reinterpret_cast<Half8*>(out + row * ldo + c) = o;
    }
}


// Routed, dilated depthwise 3x3 (DiversifiedExpertGroup.dw_layers, moe/gated.py:2267-2280, v0_14 zoo): image b runs the taps AND the
// dilation of expert e = route[b] (dilation 1 + e / 2 in the reference, passed as a table), zero padding = dilation.  One thread per
// pixel and 8 channels, 16-byte loads of the activation and of the tap row, fp32 accumulation in tap order, fp16 output.
__global__ void __launch_bounds__(256) dw3_routed_kernel(const __half* __restrict__ x, int ldx, const __half* __restrict__ w,
                                                         const int* __restrict__ route, int route_stride,
                                                         const int* __restrict__ dil, int E, int B, int H, int W, int C,
                                                         __half* __restrict__ out, int ldo) {
    const int cv = C >> 3;
    const long long total = (long long)B * H * W * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pix = i / cv;
        const int c = (int)(i - pix * cv) << 3;
        const int b = (int)(pix / ((long long)H * W));
        const int rem = (int)(pix - (long long)b * H * W);
        const int y = rem / W, xx = rem - y * W;
        int e = route[(long long)b * route_stride];
        e = e < 0 ? 0 : (e >= E ? E - 1 : e);
        const int d = dil[e];
        const __half* wt = w + (long long)e * 9 * C + c;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int sy = y + (ky - 1) * d;
            if (sy < 0 || sy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int sx = xx + (kx - 1) * d;
                if (sx < 0 || sx >= W) continue;
                const Half8 hv = *reinterpret_cast<const Half8*>(x + (((long long)b * H + sy) * W + sx) * ldx + c);
                const Half8 hw = *reinterpret_cast<const Half8*>(wt + (ky * 3 + kx) * C);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 fv = __half22float2(hv.v[j]), fw = __half22float2(hw.v[j]);
                    acc[2 * j] = fmaf(fv.x, fw.x, acc[2 * j]);
                    acc[2 * j + 1] = fmaf(fv.y, fw.y, acc[2 * j + 1]);
                }
            }
        }
        Half8 o;
#pragma unroll
        for (int j = 0; j < 4; ++j) o.v[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
        *reinterpret_cast<Half8*>(out + pix * ldo + c) = o;
    }
}

// Per-image normalisation (scale = rstd, shift = -mean * rstd from ym_groupnorm_stats with unit gamma) -> the routed expert's affine:
// scale' = scale * gamma[e], shift' = shift * gamma[e] + beta[e], e = route[b]  (the per-expert GroupNorm of dw_layers[e][1]).
__global__ void __launch_bounds__(256) route_affine_kernel(float* __restrict__ scale, float* __restrict__ shift,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const int* __restrict__ route, int route_stride, int E, int B, int C,
                                                           const float* __restrict__ route_w) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * C) return;
    const int b = (int)(i / C), c = (int)(i - (long long)b * C);
    int e = route[(long long)b * route_stride];
    e = e < 0 ? 0 : (e >= E ? E - 1 : e);
    const float g = gamma[(long long)e * C + c];
    const float rw = route_w != nullptr ? route_w[b] : 1.f;      // routing weight of image b folded into the affine
    scale[i] = scale[i] * g * rw;
    shift[i] = fmaf(shift[i], g, beta[(long long)e * C + c]) * rw;
}

}  // namespace ym

using namespace ym;

// a, b, out: fp16 [rows][ld*] (either of a/b may be null = zeros where the op allows); p0/p1: fp32 parameters of the op
// (chan[C] | scale,shift[imgs*C] | t); tok: fp32 per-token weights [rows*ldt].  C % 8 == 0, pitches % 8 == 0.
extern "C" int ym_ew_nhwc(int op, const void* a, int lda, const void* b, int ldb, const float* p0, const float* p1,
                          const float* tok, int ldt, int toff, int rows_per_img, int act, void* out, int ldo, long long rows, int C,
                          void* stream) {
    YM_CHECK_ARG(out, "ym_ew_nhwc: null output");
    YM_CHECK_ARG(C % 8 == 0 && ldo % 8 == 0 && (!a || lda % 8 == 0) && (!b || ldb % 8 == 0), "ym_ew_nhwc: multiples of 8");
    YM_CHECK_ARG(op >= 0 && op <= 8, "ym_ew_nhwc: unknown op %d", op);
    YM_CHECK_ARG(op != EW_SCALE_RES || (p0 && b), "ym_ew_nhwc: op 0 needs chan and b");
    YM_CHECK_ARG(op != EW_TOKEN_ACC || (tok && b), "ym_ew_nhwc: op 1 needs tok and b");
    YM_CHECK_ARG(op != EW_LERP || (p0 && a && b), "ym_ew_nhwc: op 5 needs t, a and b");
    YM_CHECK_ARG((op != EW_GLU) || (a && b), "ym_ew_nhwc: GLU needs a and b");
    YM_CHECK_ARG((op != EW_GELU && op != EW_SIGMOID) || a, "ym_ew_nhwc: GELU / sigmoid need a");
    YM_CHECK_ARG(op != EW_MUL_GATE || (p0 && a && b), "ym_ew_nhwc: op 7 needs t, a and b");
    YM_CHECK_ARG(op != EW_MUL || (a && b), "ym_ew_nhwc: op 8 needs a and b");
    YM_CHECK_ARG((op != EW_AFFINE) || (a && p0 && p1 && rows_per_img > 0), "ym_ew_nhwc: affine needs a, scale, shift, rows_per_img");
    if (rows == 0) return YM_OK;
    const long long total = rows * (C / 8);
    long long nb = (total + 255) / 256;
    if (nb > 148LL * 16) nb = 148LL * 16;
    cudaStream_t st = (cudaStream_t)stream;
#define EW_LAUNCH(OP)                                                                                                       \
    do {                                                                                                                    \
        auto kfn = ew_kernel<OP>;                                                                                           \
        YM_LAUNCH(kfn, (int)nb, 256, 0, st, (const __half*)a, lda, (const __half*)b, ldb, p0, p1, tok, ldt, toff, rows_per_img, \
                  act, (__half*)out, ldo, rows, C);                                                                         \
    } while (0)
    switch (op) {
        case EW_SCALE_RES: EW_LAUNCH(EW_SCALE_RES); break;
        case EW_TOKEN_ACC: EW_LAUNCH(EW_TOKEN_ACC); break;
        case EW_GLU: EW_LAUNCH(EW_GLU); break;
        case EW_GELU: EW_LAUNCH(EW_GELU); break;
        case EW_LERP: EW_LAUNCH(EW_LERP); break;
        case EW_SIGMOID: EW_LAUNCH(EW_SIGMOID); break;
        case EW_MUL_GATE: EW_LAUNCH(EW_MUL_GATE); break;
        case EW_MUL: EW_LAUNCH(EW_MUL); break;
        default: EW_LAUNCH(EW_AFFINE); break;
    }
#undef EW_LAUNCH
    YM_CHECK_LAUNCH("ew_nhwc");
    return YM_OK;
}

// DiversifiedExpertGroup.dw_layers[e][0] (moe/gated.py:2272-2277) for e = route[b * route_stride]: x fp16 [B][H][W][ldx] ->
// out fp16 [B][H][W][ldo], w fp16 [E][9][C] tap-major, dil int32 [E] (device).  C, ldx, ldo multiples of 8.
extern "C" int ym_dwconv3_routed_nhwc(const void* x, int ldx, const void* w, const int* route, int route_stride, const int* dil, int E,
                                      int B, int H, int W, int C, void* out, int ldo, void* stream) {
    YM_CHECK_ARG(x && w && route && dil && out, "ym_dwconv3_routed_nhwc: null pointer");
    YM_CHECK_ARG(C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && ldx >= C && ldo >= C, "ym_dwconv3_routed_nhwc: multiples of 8");
    YM_CHECK_ARG(E >= 1 && B >= 0 && H > 0 && W > 0 && route_stride >= 1, "ym_dwconv3_routed_nhwc: sizes");
    if (B == 0) return YM_OK;
    const long long total = (long long)B * H * W * (C / 8);
    long long nb = (total + 255) / 256;
    if (nb > 148LL * 16) nb = 148LL * 16;
    YM_LAUNCH(dw3_routed_kernel, (int)nb, 256, 0, (cudaStream_t)stream, (const __half*)x, ldx, (const __half*)w, route, route_stride, dil,
              E, B, H, W, C, (__half*)out, ldo);
    YM_CHECK_LAUNCH("dwconv3_routed");
    return YM_OK;
}

// In place on scale / shift fp32 [B][C]: the routed expert's GroupNorm affine (gamma, beta fp32 [E][C]).
extern "C" int ym_route_affine(float* scale, float* shift, const float* gamma, const float* beta, const int* route, int route_stride,
                               int E, int B, int C, const float* route_w, void* stream) {
    YM_CHECK_ARG(scale && shift && gamma && beta && route, "ym_route_affine: null pointer");
    YM_CHECK_ARG(E >= 1 && B >= 0 && C > 0 && route_stride >= 1, "ym_route_affine: sizes");
    if (B == 0) return YM_OK;
    const long long total = (long long)B * C;
    YM_LAUNCH(route_affine_kernel, (int)((total + 255) / 256), 256, 0, (cudaStream_t)stream, scale, shift, gamma, beta, route,
              route_stride, E, B, C, route_w);
    YM_CHECK_LAUNCH("route_affine");
    return YM_OK;
}
