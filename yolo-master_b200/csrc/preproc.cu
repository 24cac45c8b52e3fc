// Predictor pre- and post-processing around the forward pass.  HBM-bound byte / index work:
//   ym_letterbox_u8   LetterBox (cv2.resize INTER_LINEAR fixed point + constant border) + BGR->RGB + HWC->CHW (+ /255, fp16)
//   ym_scale_boxes    ops.scale_boxes + clip_boxes on the (B, K, 6) / ragged NMS result rows
#include "ym_common.cuh"

#include "preproc_core.cuh"

namespace ym {

constexpr int LB_TX = 64, LB_TY = 4, LB_PX = 4;   // 256 threads, each 4 consecutive output pixels of one row

template <typename T>
__device__ __forceinline__ T lb_cvt(int v);
template <>
__device__ __forceinline__ uint8_t lb_cvt<uint8_t>(int v) { return (uint8_t)v; }
template <>
__device__ __forceinline__ float lb_cvt<float>(int v) { return __fdiv_rn((float)v, 255.f); }   // im.float() / 255 predictor.py:173-175
template <>
__device__ __forceinline__ __half lb_cvt<__half>(int v) { return __float2half_rn(__fdiv_rn((float)v, 255.f)); }

template <typename T>
struct alignas(sizeof(T) * 4) Vec4 {
    T v[4];
};

// out: CHW ? [B][3][H][W] : [B][H][W][3].  grid = (ceil(W / 256), ceil(H / 4), B).
template <typename T, bool CHW>
__global__ void __launch_bounds__(LB_TX* LB_TY) letterbox_kernel(const uint8_t* __restrict__ src, long long src_stride,
                                                                 const LbTap* __restrict__ xt, const LbTap* __restrict__ yt,
                                                                 T* __restrict__ out, const LbGeom g) {
    const int dy = blockIdx.y * LB_TY + threadIdx.y;
    const int dx0 = (blockIdx.x * LB_TX + threadIdx.x) * LB_PX;
    if (dy >= g.H || dx0 >= g.W) return;
    const uint8_t* s = src + (long long)blockIdx.z * src_stride;
    T* o = out + (long long)blockIdx.z * 3 * g.H * g.W;
    int v[LB_PX][3];
#pragma unroll
    for (int j = 0; j < LB_PX; ++j) {
        if (dx0 + j < g.W) lb_output_pixel(s, g, xt, yt, dx0 + j, dy, v[j]);
        else v[j][0] = v[j][1] = v[j][2] = 0;
    }
    if (CHW) {
        const bool vec = (g.W % LB_PX) == 0;   // rows and planes then start on a 4-element boundary
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            T* row = o + ((long long)c * g.H + dy) * g.W + dx0;
            if (vec) {
                Vec4<T> pk;
#pragma unroll
                for (int j = 0; j < LB_PX; ++j) pk.v[j] = lb_cvt<T>(v[j][c]);
                *reinterpret_cast<Vec4<T>*>(row) = pk;
            } else {
#pragma unroll
                for (int j = 0; j < LB_PX; ++j)
                    if (dx0 + j < g.W) row[j] = lb_cvt<T>(v[j][c]);
            }
        }
    } else {
        T* row = o + ((long long)dy * g.W + dx0) * 3;
#pragma unroll
        for (int j = 0; j < LB_PX; ++j)
            if (dx0 + j < g.W) {
#pragma unroll
                for (int c = 0; c < 3; ++c) row[3 * j + c] = lb_cvt<T>(v[j][c]);
            }
    }
}

constexpr int SB_MAX_IMG = 128;
struct ScaleBoxParams {
    float p[SB_MAX_IMG][5];   // gain, pad_x, pad_y, w0, h0 of each image
};

struct DivRn {
    __device__ __forceinline__ float operator()(float a, float b) const { return __fdiv_rn(a, b); }
};

__global__ void __launch_bounds__(256) scale_boxes_kernel(float* __restrict__ boxes, int ld, long long n, int rows_per_img,
                                                          const int* __restrict__ row_img, int padding, int xywh,
                                                          const __grid_constant__ ScaleBoxParams sp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int img = row_img ? row_img[i] : (int)(i / rows_per_img);
    scale_box(boxes + i * ld, sp.p[img], padding, xywh, DivRn());
}

struct ScaleCoordParams {
    float p[5];
};
__global__ void __launch_bounds__(256) scale_coords_kernel(float* __restrict__ coords, int ld, long long n, int padding, int normalize,
                                                           const ScaleCoordParams sp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    scale_coord(coords + i * ld, sp.p, padding, normalize, DivRn());
}

constexpr int KPT_MAX_LEVELS = 8;
struct KptLevels {
    const float* kpt[KPT_MAX_LEVELS];   // fp32 [B][h][w][nk]
    int h[KPT_MAX_LEVELS], w[KPT_MAX_LEVELS], a0[KPT_MAX_LEVELS + 1];   // a0 = first anchor of the level
    float stride[KPT_MAX_LEVELS];
    int nl;
};
struct SigmoidRn {
    __device__ __forceinline__ float operator()(float v) const { return __fdiv_rn(1.f, 1.f + expf(-v)); }
};

// y fp32 [B][nk][A]; one thread per element, anchor fastest (coalesced stores)
__global__ void __launch_bounds__(256) kpts_decode_kernel(const KptLevels lv, int B, int nk, int ndim, int A, float* __restrict__ y) {
    const long long total = (long long)B * nk * A;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(i % A), k = (int)((i / A) % nk), b = (int)(i / ((long long)A * nk));
        int l = 0;
        while (l + 1 < lv.nl && a >= lv.a0[l + 1]) ++l;
        const int p = a - lv.a0[l], gy = p / lv.w[l], gx = p - gy * lv.w[l];
        const float raw = lv.kpt[l][((long long)b * lv.h[l] * lv.w[l] + p) * nk + k];
        y[i] = kpt_decode_value(raw, k % ndim, ndim, gx, gy, lv.stride[l], SigmoidRn());
    }
}

struct SinCosDev {
    __device__ __forceinline__ void operator()(float a, float* s, float* c) const { sincosf(a, s, c); }
};

// yin fp32 [B][4+nc][A] (xywh dense decode) -> yout fp32 [B][4+nc+1][A]: rotated centre, w, h, class scores, angle
__global__ void __launch_bounds__(256) obb_finish_kernel(const KptLevels lv, int B, int nc, int A, const float* __restrict__ yin,
                                                         float* __restrict__ yout) {
    const int rows = 4 + nc;
    const long long total = (long long)B * A;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(i % A), b = (int)(i / A);
        int l = 0;
        while (l + 1 < lv.nl && a >= lv.a0[l + 1]) ++l;
        const int p = a - lv.a0[l], gy = p / lv.w[l], gx = p - gy * lv.w[l];
        const float* src = yin + (long long)b * rows * A + a;
        float* dst = yout + (long long)b * (rows + 1) * A + a;
        float ox, oy, ang;
        obb_rotate(src[0], src[A], lv.kpt[l][(long long)b * lv.h[l] * lv.w[l] + p], gx, gy, lv.stride[l], &ox, &oy, &ang, SigmoidRn(), SinCosDev());
        dst[0] = ox;
        dst[A] = oy;
        for (int r = 2; r < rows; ++r) dst[(long long)r * A] = src[(long long)r * A];
        dst[(long long)rows * A] = ang;
    }
}

}  // namespace ym

using namespace ym;

extern "C" int ym_obb_finish(int nl, const void* const* angle, const int* hs, const int* ws, const float* strides, int B, int nc,
                             const float* yin, float* yout, void* stream) {
    YM_CHECK_ARG(angle && hs && ws && strides && yin && yout, "ym_obb_finish: null pointer");
    YM_CHECK_ARG(nl >= 1 && nl <= KPT_MAX_LEVELS && nc >= 1, "ym_obb_finish: 1..%d levels", KPT_MAX_LEVELS);
    if (B <= 0) return YM_OK;
    KptLevels lv;
    lv.nl = nl;
    int A = 0;
    for (int l = 0; l < nl; ++l) {
        YM_CHECK_ARG(angle[l] && hs[l] > 0 && ws[l] > 0, "ym_obb_finish: bad level %d", l);
        lv.kpt[l] = (const float*)angle[l]; lv.h[l] = hs[l]; lv.w[l] = ws[l]; lv.stride[l] = strides[l]; lv.a0[l] = A;
        A += hs[l] * ws[l];
    }
    lv.a0[nl] = A;
    long long nb = ((long long)B * A + 255) / 256;
    if (nb > 148LL * 16) nb = 148LL * 16;
    YM_LAUNCH(obb_finish_kernel, (int)nb, 256, 0, (cudaStream_t)stream, lv, B, nc, A, yin, yout);
    YM_CHECK_LAUNCH("obb_finish");
    return YM_OK;
}

extern "C" int ym_kpts_decode(int nl, const void* const* kpt, const int* hs, const int* ws, const float* strides, int B, int nk,
                              int ndim, float* y, void* stream) {
    YM_CHECK_ARG(kpt && hs && ws && strides && y, "ym_kpts_decode: null pointer");
    YM_CHECK_ARG(nl >= 1 && nl <= KPT_MAX_LEVELS && nk >= 1 && ndim >= 1 && ndim <= 3 && nk % ndim == 0,
                 "ym_kpts_decode: 1..%d levels, ndim 1 (copy), 2 or 3 dividing nk", KPT_MAX_LEVELS);
    if (B <= 0) return YM_OK;
    KptLevels lv;
    lv.nl = nl;
    int A = 0;
    for (int l = 0; l < nl; ++l) {
        YM_CHECK_ARG(kpt[l] && hs[l] > 0 && ws[l] > 0, "ym_kpts_decode: bad level %d", l);
        lv.kpt[l] = (const float*)kpt[l]; lv.h[l] = hs[l]; lv.w[l] = ws[l]; lv.stride[l] = strides[l]; lv.a0[l] = A;
        A += hs[l] * ws[l];
    }
    lv.a0[nl] = A;
    const long long total = (long long)B * nk * A;
    long long nb = (total + 255) / 256;
    if (nb > 148LL * 16) nb = 148LL * 16;
    YM_LAUNCH(kpts_decode_kernel, (int)nb, 256, 0, (cudaStream_t)stream, lv, B, nk, ndim, A, y);
    YM_CHECK_LAUNCH("kpts_decode");
    return YM_OK;
}

extern "C" int ym_letterbox_u8(const void* src, long long src_stride, int B, int sh, int sw, int src_pitch, const void* xtab,
                               const void* ytab, int area2x, int nw, int nh, int top, int left, int pad_value, int swap_rb,
                               void* out, int out_dtype, int chw, int H, int W, void* stream) {
    YM_CHECK_ARG(src && out, "ym_letterbox_u8: null pointer");
    YM_CHECK_ARG(B >= 0 && B <= 65535, "ym_letterbox_u8: 0 <= B <= 65535 (got %d)", B);
    YM_CHECK_ARG(sh > 0 && sw > 0 && sh <= 65535 && sw <= 65535 && src_pitch >= 3 * sw, "ym_letterbox_u8: bad source %dx%d pitch %d",
                 sh, sw, src_pitch);
    YM_CHECK_ARG(nw > 0 && nh > 0 && top >= 0 && left >= 0 && top + nh <= H && left + nw <= W,
                 "ym_letterbox_u8: resized %dx%d at (%d,%d) does not fit %dx%d", nh, nw, top, left, H, W);
    YM_CHECK_ARG(area2x ? (sw == 2 * nw && sh == 2 * nh) : (xtab && ytab), "ym_letterbox_u8: tables missing / not an exact 2x downscale");
    YM_CHECK_ARG(out_dtype >= 0 && out_dtype <= 2, "ym_letterbox_u8: out_dtype 0 uint8, 1 fp16, 2 fp32");
    YM_CHECK_ARG(pad_value >= 0 && pad_value <= 255, "ym_letterbox_u8: pad value");
    if (B == 0) return YM_OK;
    LbGeom g;
    g.sh = sh; g.sw = sw; g.src_pitch = src_pitch; g.nw = nw; g.nh = nh; g.top = top; g.left = left; g.H = H; g.W = W;
    g.pad = pad_value; g.swap_rb = swap_rb ? 1 : 0; g.area2x = area2x ? 1 : 0;
    dim3 grid((W + LB_TX * LB_PX - 1) / (LB_TX * LB_PX), (H + LB_TY - 1) / LB_TY, B), block(LB_TX, LB_TY);
    YM_CHECK_ARG(grid.y <= 65535, "ym_letterbox_u8: output too tall");
    cudaStream_t st = (cudaStream_t)stream;
    const uint8_t* s = (const uint8_t*)src;
    const LbTap *xt = (const LbTap*)xtab, *yt = (const LbTap*)ytab;
#define LB_LAUNCH(T, CHW)                                                      \
    do {                                                                       \
        auto kfn = letterbox_kernel<T, CHW>;                                   \
        YM_LAUNCH(kfn, grid, block, 0, st, s, src_stride, xt, yt, (T*)out, g); \
    } while (0)
    if (chw) {
        if (out_dtype == 0) LB_LAUNCH(uint8_t, true);
        else if (out_dtype == 1) LB_LAUNCH(__half, true);
        else LB_LAUNCH(float, true);
    } else {
        if (out_dtype == 0) LB_LAUNCH(uint8_t, false);
        else if (out_dtype == 1) LB_LAUNCH(__half, false);
        else LB_LAUNCH(float, false);
    }
#undef LB_LAUNCH
    YM_CHECK_LAUNCH("letterbox_u8");
    return YM_OK;
}

extern "C" int ym_scale_boxes(float* boxes, int ld, long long n, int rows_per_img, const int* row_img, int n_img,
                              const float* params_host, int padding, int xywh, void* stream) {
    YM_CHECK_ARG(n == 0 || boxes, "ym_scale_boxes: null boxes");
    YM_CHECK_ARG(ld >= 4, "ym_scale_boxes: row pitch >= 4");
    YM_CHECK_ARG(n_img >= 1 && n_img <= SB_MAX_IMG && params_host, "ym_scale_boxes: 1..%d images per call", SB_MAX_IMG);
    YM_CHECK_ARG(row_img || (rows_per_img > 0 && n <= (long long)rows_per_img * n_img), "ym_scale_boxes: rows_per_img * n_img < n");
    if (n == 0) return YM_OK;
    ScaleBoxParams sp;
    memset(&sp, 0, sizeof(sp));
    memcpy(sp.p, params_host, sizeof(float) * 5 * n_img);
    YM_LAUNCH(scale_boxes_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, boxes, ld, n, rows_per_img, row_img, padding,
              xywh, sp);
    YM_CHECK_LAUNCH("scale_boxes");
    return YM_OK;
}

// ops.scale_coords (utils/ops.py:596-631) in place for the points of ONE image: coords fp32 rows of pitch ld >= 2 with (x, y) in the
// first two columns (keypoints (n, nk, 2 | 3): ld = 2 | 3, rows = n * nk); params_host = (gain, pad_x, pad_y, w0, h0).
extern "C" int ym_scale_coords(float* coords, int ld, long long n, const float* params_host, int padding, int normalize, void* stream) {
    YM_CHECK_ARG(n == 0 || coords, "ym_scale_coords: null coords");
    YM_CHECK_ARG(ld >= 2 && n >= 0 && params_host, "ym_scale_coords: row pitch >= 2");
    if (n == 0) return YM_OK;
    ScaleCoordParams sp;
    memcpy(sp.p, params_host, sizeof(sp.p));
    YM_LAUNCH(scale_coords_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, coords, ld, n, padding, normalize, sp);
    YM_CHECK_LAUNCH("scale_coords");
    return YM_OK;
}
