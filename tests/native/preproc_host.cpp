// TEST INFRASTRUCTURE (g++ only, no CUDA): runs the per-pixel / per-box code of yolo-master_b200/csrc/preproc_core.cuh on the
// host, over the same index space the kernels cover, so tests/test_preproc_host.py can compare the integer arithmetic of
// ym_letterbox_u8 / ym_scale_boxes with the oracle in the GPU-less build container.  Built by the test into a temp dir.
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
