// TEST INFRASTRUCTURE (g++ only, no CUDA): runs the per-image algorithm of yolo-master_b200/csrc/nms_large_core.cuh with a host
// executor (ex.all(f) = f(tid) for every tid, in order - the barrier is implicit), so tests/test_nms_large_host.py can compare the
// large-candidate NMS path with the NMS oracle in the GPU-less build container.  Same argument list as ym_nms_batched_large
// minus scratch / stream.
#include <vector>

#include "nms_large_core.cuh"

using namespace ym::nmsl;

struct HostExec {
    int nthr;
    template <class F>
    void all(F f) {
        for (int t = 0; t < nthr; ++t) f(t);
    }
};

extern "C" int host_nms_batched_large(const float* pred, int B, int nc, int A, float conf_thres, float iou_thres, int max_det,
                                      int max_nms, float max_wh, float* out, int* out_count, int* out_idx) {
    int NP = 1;
    while (NP < A) NP <<= 1;
    std::vector<float> conf((size_t)B * A);
    std::vector<int> cls((size_t)B * A);
    for (int b = 0; b < B; ++b)
        for (int a = 0; a < A; ++a) {                          // best class per anchor, first maximum wins
            const float* p = pred + ((long long)b * (4 + nc) + 4) * A + a;
            float best = p[0];
            int bi = 0;
            for (int c = 1; c < nc; ++c)
                if (p[(long long)c * A] > best) { best = p[(long long)c * A]; bi = c; }
            conf[(size_t)b * A + a] = best;
            cls[(size_t)b * A + a] = bi;
        }
    std::vector<unsigned long long> keys((size_t)B * NP);
    std::vector<Box4> sbox((size_t)B * A);
    std::vector<unsigned char> sup((size_t)B * A);
    for (long long i = 0; i < (long long)B * max_det * 6; ++i) out[i] = 0.f;
    for (long long i = 0; i < (long long)B * max_det; ++i) out_idx[i] = -1;
    Args a;
    a.pred = pred; a.conf = conf.data(); a.cls = cls.data(); a.nc = nc; a.A = A; a.NP = NP; a.conf_thres = conf_thres;
    a.iou_thres = iou_thres; a.max_wh = max_wh; a.max_det = max_det; a.max_nms = max_nms; a.keys = keys.data(); a.sbox = sbox.data();
    a.sup = sup.data(); a.out = out; a.out_count = out_count; a.out_idx = out_idx;
    HostExec ex{1024};
    Shared sh;
    for (int b = 0; b < B; ++b) image(a, b, ex, sh);
    return 0;
}
