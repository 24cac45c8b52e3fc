// TEST INFRASTRUCTURE (g++ only, no CUDA): runs the per-pixel / per-box code of yolo-master_b200/csrc/preproc_core.cuh on the
// host, over the same index space the kernels cover, so tests/test_preproc_host.py can compare the integer arithmetic of
// ym_letterbox_u8 / ym_scale_boxes with the oracle in the GPU-less build container.  Built by the test into a temp dir.
#include <math.h>

#include "preproc_core.cuh"

using namespace ym;

extern "C" void host_letterbox_u8(const uint8_t* src, int sh, int sw, int src_pitch, const void* xtab, const void* ytab,
                                  int area2x, int nw, int nh, int top, int left, int pad_value, int swap_rb, uint8_t* out, int chw,
                                  int H, int W) {
    LbGeom g;
    g.sh = sh; g.sw = sw; g.src_pitch = src_pitch; g.nw = nw; g.nh = nh; g.top = top; g.left = left; g.H = H; g.W = W;
    g.pad = pad_value; g.swap_rb = swap_rb; g.area2x = area2x;
    for (int dy = 0; dy < H; ++dy)
        for (int dx = 0; dx < W; ++dx) {
            int v[3];
            lb_output_pixel(src, g, (const LbTap*)xtab, (const LbTap*)ytab, dx, dy, v);
            for (int c = 0; c < 3; ++c) {
                if (chw) out[((long long)c * H + dy) * W + dx] = (uint8_t)v[c];
                else out[((long long)dy * W + dx) * 3 + c] = (uint8_t)v[c];
            }
        }
}

struct SigmoidHost {
    float operator()(float v) const { return 1.f / (1.f + expf(-v)); }
};

// one level at a time: kpt fp32 [B][h][w][nk] -> y[b][k][a0 + p] of the [B][nk][A] output
extern "C" void host_kpts_decode_level(const float* kpt, int h, int w, float stride, int a0, int B, int nk, int ndim, int A, float* y) {
    for (int b = 0; b < B; ++b)
        for (int k = 0; k < nk; ++k)
            for (int p = 0; p < h * w; ++p)
                y[((long long)b * nk + k) * A + a0 + p] =
                    kpt_decode_value(kpt[((long long)b * h * w + p) * nk + k], k % ndim, ndim, p % w, p / w, stride, SigmoidHost());
}

struct SinCosHost {
    void operator()(float a, float* s, float* c) const { *s = sinf(a); *c = cosf(a); }
};

// one level: angle fp32 [B][h][w][1]; yin [B][rows][A] -> yout [B][rows+1][A] for the anchors a0 .. a0 + h*w
extern "C" void host_obb_finish_level(const float* angle, int h, int w, float stride, int a0, int B, int nc, int A, const float* yin,
                                      float* yout) {
    const int rows = 4 + nc;
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < h * w; ++p) {
            const float* src = yin + (long long)b * rows * A + a0 + p;
            float* dst = yout + (long long)b * (rows + 1) * A + a0 + p;
            float ox, oy, ang;
            obb_rotate(src[0], src[A], angle[(long long)b * h * w + p], p % w, p / w, stride, &ox, &oy, &ang, SigmoidHost(), SinCosHost());
            dst[0] = ox; dst[A] = oy;
            for (int r = 2; r < rows; ++r) dst[(long long)r * A] = src[(long long)r * A];
            dst[(long long)rows * A] = ang;
        }
}

struct DivHost {
    float operator()(float a, float b) const { return a / b; }
};

extern "C" void host_scale_boxes(float* boxes, int ld, long long n, int rows_per_img, const int* row_img, const float* params,
                                 int padding, int xywh) {
    for (long long i = 0; i < n; ++i) {
        const int img = row_img ? row_img[i] : (int)(i / rows_per_img);
        scale_box(boxes + i * ld, params + 5 * img, padding, xywh, DivHost());
    }
}
