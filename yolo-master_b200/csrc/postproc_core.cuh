// Per-element bodies of the task post-processing that follows the Segment and OBB heads:
//   * mask assembly   - ops.process_mask (ultralytics/utils/ops.py:500-528): coefficients x prototypes, bilinear upsampling to the
//                       network input size (F.interpolate, align_corners=False), crop to the box (crop_mask :477-497), > 0 -> uint8;
//   * rotated NMS     - the `rotated` branch of non_max_suppression (utils/nms.py:148-152): TorchNMS.fast_nms (:193-242) with
//                       batch_probiou (utils/metrics.py:293-326, covariance :224-242) - candidate j survives iff NO higher-scoring
//                       candidate i < j (suppressed or not) has ProbIoU >= iou_thres.
// Everything here is __host__ __device__ and free of CUDA built-ins, so that the same code is compared with the oracle in the
// GPU-less build container (tests/native/cuda_host_emu.h builds the whole translation unit with g++).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#ifndef YM_HD
#define YM_HD __host__ __device__ __forceinline__
#endif
#else
#ifndef YM_HD
#define YM_HD inline
#endif
#endif

namespace ym {
namespace pp {

// ---------------------------------------------------------------------------------------------------------------- masks
// One axis of ATen's upsample_bilinear2d with align_corners=False and no scale_factor (the `size=` call of ops.py:520):
// scale = in / out (fp32), src = scale * (dst + 0.5) - 0.5 clamped at 0, i0 = min(floor(src), in - 1), i1 = i0 + (i0 < in - 1),
// lambda = clamp(src - i0, 0, 1).
struct Lerp {
    int i0, i1;
    float w0, w1;
};
YM_HD Lerp lerp_axis(int dst, int in_size, float scale) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    int i0 = (int)floorf(src);
    if (i0 > in_size - 1) i0 = in_size - 1;
    float l = src - (float)i0;
    l = l < 0.f ? 0.f : (l > 1.f ? 1.f : l);
    Lerp r;
    r.i0 = i0;
    r.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    r.w0 = 1.f - l;
    r.w1 = l;
    return r;
}

// crop_mask (ops.py:494-495): a pixel survives iff x1 <= col < x2 and y1 <= row < y2 (fp32 comparisons; NaN boxes keep nothing).
YM_HD bool in_box(float col, float row, float x1, float y1, float x2, float y2) {
    return col >= x1 && col < x2 && row >= y1 && row < y2;
}

// One output pixel of process_mask.  lg: the detection's fp32 logit plane (mh x mw) = coefficients @ prototypes.
// upsample != 0: pixel (x, y) of the (oh x ow) network-input frame, box in that frame.
// upsample == 0: pixel of the prototype plane, box already multiplied by (mw / w, mh / h) (ops.py:523-526).
YM_HD unsigned char mask_pixel(const float* lg, int mh, int mw, int x, int y, int upsample, float sx, float sy, float x1, float y1,
                               float x2, float y2) {
    if (!in_box((float)x, (float)y, x1, y1, x2, y2)) return 0;
    float v;
    if (upsample) {
        const Lerp lx = lerp_axis(x, mw, sx), ly = lerp_axis(y, mh, sy);
        const float* r0 = lg + (long long)ly.i0 * mw;
        const float* r1 = lg + (long long)ly.i1 * mw;
        v = ly.w0 * (lx.w0 * r0[lx.i0] + lx.w1 * r0[lx.i1]) + ly.w1 * (lx.w0 * r1[lx.i0] + lx.w1 * r1[lx.i1]);
    } else {
        v = lg[(long long)y * mw + x];
    }
    return v > 0.f ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------- rotated NMS
struct RBox {        // one sorted candidate: class-offset centre and its Gaussian (metrics.py:224-242), 32 bytes
    float x, y;      // cx + cls * max_wh, cy + cls * max_wh  (nms.py:148,151)
    float a, b, c;   // covariance [[a, c], [c, b]]
    float det;       // clamp(a * b - c^2, 0)
    int anchor;      // source column of the prediction
    int pad;
};

template <class SinCos>
YM_HD RBox make_rbox(float cx, float cy, float w, float h, float ang, float off, int anchor, SinCos sincos_fn) {
    const float ga = w * w / 12.f, gb = h * h / 12.f;
    float s, c;
    sincos_fn(ang, &s, &c);
    const float c2 = c * c, s2 = s * s;
    RBox r;
    r.x = cx + off;
    r.y = cy + off;
    r.a = ga * c2 + gb * s2;
    r.b = ga * s2 + gb * c2;
    r.c = (ga - gb) * c * s;
    const float d = r.a * r.b - r.c * r.c;
    r.det = d > 0.f ? d : 0.f;
    r.anchor = anchor;
    r.pad = 0;
    return r;
}

// batch_probiou(obb1 = p, obb2 = q) for one pair (metrics.py:307-326), eps = 1e-7.
template <class Log, class Exp>
YM_HD float probiou(const RBox& p, const RBox& q, Log log_fn, Exp exp_fn) {
    const float eps = 1e-7f;
    const float A = p.a + q.a, B = p.b + q.b, C = p.c + q.c;
    const float dx = p.x - q.x, dy = p.y - q.y;
    const float core = A * B - C * C;
    const float den = core + eps;
    const float t1 = ((A * dy * dy + B * dx * dx) / den) * 0.25f;
    const float t2 = ((C * (-dx) * dy) / den) * 0.5f;
    const float t3 = log_fn(core / (4.f * sqrtf(p.det * q.det) + eps) + eps) * 0.5f;
    float bd = t1 + t2 + t3;
    bd = bd < eps ? eps : (bd > 100.f ? 100.f : bd);     // clamp(eps, 100): NaN stays NaN and never reaches iou_thres
    const float hd = sqrtf(1.f - exp_fn(-bd) + eps);
    return 1.f - hd;
}

struct SinCosF {
    YM_HD void operator()(float a, float* s, float* c) const {
        *s = sinf(a);
        *c = cosf(a);
    }
};
struct LogF {
    YM_HD float operator()(float v) const { return logf(v); }
};
struct ExpF {
    YM_HD float operator()(float v) const { return expf(v); }
};

YM_HD uint32_t f2key(float f) {
    union { float f; uint32_t u; } v;
    v.f = f;
    return (v.u & 0x80000000u) ? ~v.u : (v.u | 0x80000000u);
}

}  // namespace pp
}  // namespace ym
