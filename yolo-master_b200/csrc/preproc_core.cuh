// Per-pixel / per-box arithmetic of the predictor's pre- and post-processing, shared by the kernels in preproc.cu and by the
// host-compiled check in tests/native/preproc_host.cpp (g++ includes this header with the qualifiers defined away, so the very
// same integer code is compared with the oracle in the GPU-less build container).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define YM_HD __host__ __device__ __forceinline__
#else
#define YM_HD inline
#endif

namespace ym {

// Geometry of one letterboxed frame (LetterBox.get_params data/augment.py:1742-1786).
struct LbGeom {
    int sh, sw, src_pitch;   // source frame, bytes per source row (3 interleaved uint8 channels)
    int nw, nh;              // resized ("new_unpad") width / height
    int top, left;           // border in front of the resized image
    int H, W;                // output frame
    int pad;                 // border value (114)
    int swap_rb;             // 1: output channel c reads source channel 2-c (BGR -> RGB, engine/predictor.py:169)
    int area2x;              // 1: exact 2x downscale in both axes -> cv2's INTER_AREA fast path
};

// One axis table entry: idx = i0 | i1 << 16 (source indices, already clipped), wgt = a0 | a1 << 16 (11-bit weights, a0+a1=2048).
struct LbTap {
    uint32_t idx, wgt;
};

// cv2.resize(INTER_LINEAR) on 8-bit data, one output pixel of the resized image (x < nw, y < nh), three channels.
//   horizontal:  S = p[x0]*a0 + p[x1]*a1                                   (HResizeLinear, int32)
//   vertical:    v = (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2    (VResizeLinear<uchar,int,short>)
YM_HD void lb_resized_pixel(const uint8_t* src, const LbGeom& g, const LbTap* xt, const LbTap* yt, int x, int y, int (&v)[3]) {
    if (g.area2x) {
        const uint8_t* r0 = src + (long long)(2 * y) * g.src_pitch + 6 * x;
        const uint8_t* r1 = r0 + g.src_pitch;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = (r0[c] + r0[3 + c] + r1[c] + r1[3 + c] + 2) >> 2;
        return;
    }
    const LbTap tx = xt[x], ty = yt[y];
    const int x0 = (int)(tx.idx & 0xffffu) * 3, x1 = (int)(tx.idx >> 16) * 3;
    const int a0 = (int)(tx.wgt & 0xffffu), a1 = (int)(tx.wgt >> 16);
    const int b0 = (int)(ty.wgt & 0xffffu), b1 = (int)(ty.wgt >> 16);
    const uint8_t* r0 = src + (long long)(ty.idx & 0xffffu) * g.src_pitch;
    const uint8_t* r1 = src + (long long)(ty.idx >> 16) * g.src_pitch;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int s0 = r0[x0 + c] * a0 + r0[x1 + c] * a1;
        const int s1 = r1[x0 + c] * a0 + r1[x1 + c] * a1;
        int o = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
        v[c] = o < 0 ? 0 : (o > 255 ? 255 : o);
    }
}

// One pixel of the letterboxed OUTPUT frame (copyMakeBorder BORDER_CONSTANT augment.py:1807-1810), in output channel order.
YM_HD void lb_output_pixel(const uint8_t* src, const LbGeom& g, const LbTap* xt, const LbTap* yt, int dx, int dy, int (&v)[3]) {
    const int x = dx - g.left, y = dy - g.top;
    if (x < 0 || x >= g.nw || y < 0 || y >= g.nh) {
        v[0] = v[1] = v[2] = g.pad;
        return;
    }
    lb_resized_pixel(src, g, xt, yt, x, y, v);
    if (g.swap_rb) {
        const int t = v[0];
        v[0] = v[2];
        v[2] = t;
    }
}

// ops.scale_boxes + clip_boxes (utils/ops.py:119-158,174-201) for one box: p = (gain, pad_x, pad_y, w0, h0).
// `div` is an IEEE round-to-nearest division (the translation units are built with --use_fast_math, so it is passed in).
template <typename Div>
YM_HD void scale_box(float* b, const float* p, int padding, int xywh, Div div) {
    float x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
    if (padding) {
        x1 -= p[1];
        y1 -= p[2];
        if (!xywh) {
            x2 -= p[1];
            y2 -= p[2];
        }
    }
    x1 = div(x1, p[0]);
    y1 = div(y1, p[0]);
    x2 = div(x2, p[0]);
    y2 = div(y2, p[0]);
    if (!xywh) {   // clamp_(0, w) keeps NaN
        x1 = x1 < 0.f ? 0.f : (x1 > p[3] ? p[3] : x1);
        y1 = y1 < 0.f ? 0.f : (y1 > p[4] ? p[4] : y1);
        x2 = x2 < 0.f ? 0.f : (x2 > p[3] ? p[3] : x2);
        y2 = y2 < 0.f ? 0.f : (y2 > p[4] ? p[4] : y2);
    }
    b[0] = x1;
    b[1] = y1;
    b[2] = x2;
    b[3] = y2;
}

// ops.scale_coords + clip_coords (utils/ops.py:596-631,204-225) for one point (x, y[, ...]): p = (gain, pad_x, pad_y, w0, h0).
template <typename Div>
YM_HD void scale_coord(float* c, const float* p, int padding, int normalize, Div div) {
    float x = c[0], y = c[1];
    if (padding) {
        x -= p[1];
        y -= p[2];
    }
    x = div(x, p[0]);
    y = div(y, p[0]);
    x = x < 0.f ? 0.f : (x > p[3] ? p[3] : x);      // clamp_(0, w) keeps NaN
    y = y < 0.f ? 0.f : (y > p[4] ? p[4] : y);
    if (normalize) {
        x = div(x, p[3]);
        y = div(y, p[4]);
    }
    c[0] = x;
    c[1] = y;
}

// Pose.kpts_decode head.py:644-664 for one output element y[b][k][a] (k = keypoint*ndim + d): x / y are (v*2 + grid coordinate) *
// stride (the reference's anchor - 0.5 IS the integer grid coordinate), the optional visibility channel is a sigmoid.
// v points at the level's fp32 NHWC tower output [B][h][w][nk]; (gx, gy) is the anchor's cell, `sig` an IEEE-accurate sigmoid.
template <typename Sig>
YM_HD float kpt_decode_value(float raw, int d, int ndim, int gx, int gy, float stride, Sig sig) {
    if (ndim == 1) return raw;                       // plain NHWC levels -> (B, n, A) gather: Segment's mask coefficients
    if (ndim == 3 && d == 2) return sig(raw);
    return (raw * 2.0f + (float)(d == 0 ? gx : gy)) * stride;
}

// OBB head (head.py:477-500, dist2rbox utils/tal.py:447-453) for one anchor, applied to the axis-aligned dense decode: the centre offset
// (xf, yf) = (rb - lt) / 2 is recovered from the decoded centre, rotated by the predicted angle and re-anchored; w, h stay.
//   angle = (sigmoid(raw) - 0.25) * pi;  x' = (xf cos - yf sin + ax) * s;  y' = (xf sin + yf cos + ay) * s,  anchor (ax, ay) = cell + 0.5.
template <typename Sig, typename SinCos>
YM_HD void obb_rotate(float cx, float cy, float raw, int gx, int gy, float stride, float* ox, float* oy, float* oang, Sig sig, SinCos sc) {
    const float ang = (sig(raw) - 0.25f) * 3.14159265358979323846f;
    const float ax = (float)gx + 0.5f, ay = (float)gy + 0.5f;
    const float xf = cx / stride - ax, yf = cy / stride - ay;
    float s, c;
    sc(ang, &s, &c);
    *ox = (xf * c - yf * s + ax) * stride;
    *oy = (xf * s + yf * c + ay) * stride;
    *oang = ang;
}

}  // namespace ym
