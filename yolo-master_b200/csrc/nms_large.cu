// Batched NMS, large-candidate path (see nms_large_core.cuh): taken by utils/nms.py only for batches in which some image exceeds
// the 16384-candidate capacity of the shared-memory kernel in nms.cu (validation at conf 0.001, utils/nms.py:142-146).
#include "ym_common.cuh"

#include "nms_large_core.cuh"

namespace ym {

struct CtaExec {
    int nthr;
    template <class F>
    __host__ __device__ __forceinline__ void all(F f) {
#if defined(__CUDA_ARCH__) || defined(YM_HOST_EMU)
        f((int)threadIdx.x);
        __syncthreads();
#endif
    }
};

// best class per anchor (first maximum wins, like torch.max).  pred: [B][4+nc][A]
__global__ void __launch_bounds__(256) nmsl_best_class_kernel(const float* __restrict__ pred, int B, int nc, int A,
                                                              float* __restrict__ conf, int* __restrict__ cls) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * A) return;
    const int b = (int)(i / A), a = (int)(i % A);
    const float* p = pred + ((long long)b * (4 + nc) + 4) * A + a;
    float best = p[0];
    int bi = 0;
    for (int c = 1; c < nc; ++c) {
        const float v = p[(long long)c * A];
        if (v > best) { best = v; bi = c; }
    }
    conf[i] = best;
    cls[i] = bi;
}

__global__ void __launch_bounds__(1024) nms_large_kernel(const nmsl::Args a) {
    __shared__ nmsl::Shared sh;
    CtaExec ex{(int)blockDim.x};
    nmsl::image(a, (int)blockIdx.x, ex, sh);
}

static int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

}  // namespace ym

using namespace ym;

extern "C" long long ym_nms_large_scratch_bytes(int B, int A) {
    const long long NP = next_pow2(A);
    return (long long)B * A * (4 + 4 + 16 + 1) + (long long)B * NP * 8 + 256;
}

// Same contract as ym_nms_batched mode 0 (out fp32 [B][max_det][6] xyxy/conf/cls, out_count int32 [B], out_idx int32 [B][max_det]),
// any number of candidates per image; the first max_nms by score enter the suppression (utils/nms.py:142-146).
extern "C" int ym_nms_batched_large(const float* pred, int B, int nc, int A, float conf_thres, float iou_thres, int max_det,
                                    int max_nms, float max_wh, float* out, int* out_count, int* out_idx, void* scratch, void* stream) {
    YM_CHECK_ARG(pred && out && out_count && out_idx && scratch, "ym_nms_batched_large: null pointer");
    YM_CHECK_ARG(max_det >= 1 && max_det <= nmsl::MAX_KEEP, "ym_nms_batched_large: max_det must be in 1..512");
    YM_CHECK_ARG(B >= 0 && nc >= 1 && A >= 1 && A <= (1 << 24) && max_nms >= 1, "ym_nms_batched_large: sizes");
    if (B == 0) return YM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int NP = next_pow2(A);
    unsigned char* p = (unsigned char*)(((uintptr_t)scratch + 15) & ~(uintptr_t)15);
    nmsl::Args a;
    a.keys = (unsigned long long*)p;            p += (size_t)B * NP * 8;
    a.sbox = (nmsl::Box4*)p;                    p += (size_t)B * A * 16;
    float* conf = (float*)p;                    p += (size_t)B * A * 4;
    int* cls = (int*)p;                         p += (size_t)B * A * 4;
    a.sup = p;
    a.pred = pred; a.conf = conf; a.cls = cls; a.nc = nc; a.A = A; a.NP = NP; a.conf_thres = conf_thres; a.iou_thres = iou_thres;
    a.max_wh = max_wh; a.max_det = max_det; a.max_nms = max_nms; a.out = out; a.out_count = out_count; a.out_idx = out_idx;
    cudaMemsetAsync(out, 0, (size_t)B * max_det * 6 * sizeof(float), st);
    cudaMemsetAsync(out_idx, 0xff, (size_t)B * max_det * sizeof(int), st);
    const long long total = (long long)B * A;
    YM_LAUNCH(nmsl_best_class_kernel, (int)((total + 255) / 256), 256, 0, st, pred, B, nc, A, conf, cls);
    YM_CHECK_LAUNCH("nmsl_best_class");
    YM_LAUNCH(nms_large_kernel, B, 1024, 0, st, a);
    YM_CHECK_LAUNCH("nms_large");
    return YM_OK;
}
