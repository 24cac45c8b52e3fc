// Task post-processing after the Segment and OBB heads (see postproc_core.cuh): mask assembly and rotated NMS.
//
// Mask assembly (ops.process_mask).  Two launches on the caller's stream:
//   mask_logits_kernel  logits[n][p] = sum_c coef[n][c] * proto[c][p]  (fp32 accumulate, c ascending): one thread per prototype
//                       pixel and 8 detections, so each prototype value is read once per 8 detections (the 32 x mh x mw plane is
//                       L2-resident: 3.3 MB at 160 x 160); coefficients sit in shared memory.
//   mask_finish_kernel  one thread per 4 consecutive output pixels: crop test first (pixels outside the box cost no loads), bilinear
//                       tap of the logit plane, > 0, one 32-bit store.  HBM-bound on the uint8 output: n * H * W bytes.
// Rotated NMS (nms.py `rotated` branch = fast_nms + ProbIoU).  Four launches:
//   rnms_best_class_kernel, rnms_sort_kernel (one CTA per image: keys, bitonic sort, Gaussian of each sorted candidate),
//   rnms_suppress_kernel (grid = candidate tiles x images: candidate j against every i < j, tiles of i staged in shared memory -
//   fast-NMS has no sequential dependence, so the n^2 / 2 pair tests spread over the whole GPU), rnms_emit_kernel (ordered
//   compaction of the first max_det survivors).
#include "ym_common.cuh"

#include "postproc_core.cuh"

namespace ym {

constexpr int ML_DETS = 8;        // detections per mask_logits thread
constexpr int ML_MAX_NM = 64;     // prototype channels held in shared memory
constexpr int RN_TILE = 256;

template <typename T>
__device__ __forceinline__ float proto_at(const T* p, long long i);
template <>
__device__ __forceinline__ float proto_at<float>(const float* p, long long i) { return p[i]; }
template <>
__device__ __forceinline__ float proto_at<__half>(const __half* p, long long i) { return __half2float(p[i]); }

template <typename T>
__global__ void __launch_bounds__(256) mask_logits_kernel(const T* __restrict__ protos, int nm, int P, const float* __restrict__ dets,
                                                          int ld, int n, int coef_col, float* __restrict__ logits) {
    __shared__ float coef[ML_DETS][ML_MAX_NM];
    const int n0 = (int)blockIdx.y * ML_DETS;
    for (int i = threadIdx.x; i < ML_DETS * nm; i += blockDim.x) {
        const int k = i / nm, c = i % nm;
        coef[k][c] = (n0 + k < n) ? dets[(long long)(n0 + k) * ld + coef_col + c] : 0.f;
    }
    __syncthreads();
    const int p = (int)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float acc[ML_DETS];
#pragma unroll
    for (int k = 0; k < ML_DETS; ++k) acc[k] = 0.f;
    for (int c = 0; c < nm; ++c) {
        const float v = proto_at<T>(protos, (long long)c * P + p);
#pragma unroll
        for (int k = 0; k < ML_DETS; ++k) acc[k] += coef[k][c] * v;
    }
#pragma unroll
    for (int k = 0; k < ML_DETS; ++k)
        if (n0 + k < n) logits[(long long)(n0 + k) * P + p] = acc[k];
}

struct MaskArgs {
    const float* logits;   // [n][mh * mw]
    const float* dets;     // [n][ld], box at columns 0..3
    int ld, n, mh, mw, oh, ow, upsample;
    float sx, sy;          // mw / ow, mh / oh as fp32 (interpolation scale with upsample, box ratio without)
    unsigned char* out;    // [n][oh][ow]
};

__global__ void __launch_bounds__(256) mask_finish_kernel(const MaskArgs a) {
    const int qw = (a.ow + 3) >> 2;                                   // 4-pixel groups per row
    const long long total = (long long)a.n * a.oh * qw;
    const bool vec = (a.ow & 3) == 0 && (((uintptr_t)a.out) & 3) == 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(t % qw);
        const long long ry = t / qw;
        const int y = (int)(ry % a.oh), d = (int)(ry / a.oh);
        const float* bx = a.dets + (long long)d * a.ld;
        float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
        if (!a.upsample) {                                            // bboxes * ratios (ops.py:523-526): one fp32 rounding each
            x1 = __fmul_rn(x1, a.sx); x2 = __fmul_rn(x2, a.sx);
            y1 = __fmul_rn(y1, a.sy); y2 = __fmul_rn(y2, a.sy);
        }
        const float* lg = a.logits + (long long)d * a.mh * a.mw;
        unsigned char v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = q * 4 + k;
            v[k] = x < a.ow ? pp::mask_pixel(lg, a.mh, a.mw, x, y, a.upsample, a.sx, a.sy, x1, y1, x2, y2) : 0;
        }
        unsigned char* o = a.out + ((long long)d * a.oh + y) * a.ow + q * 4;
        if (vec) {
            *(uint32_t*)o = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (q * 4 + k < a.ow) o[k] = v[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------ rotated NMS
struct RnmsArgs {
    const float* pred;     // [B][4 + nc + 1][A]: xywh, class scores, angle
    float* conf;           // [B][A]
    int* cls;              // [B][A]
    unsigned long long* keys;   // [B][NP]
    pp::RBox* rbox;        // [B][A] sorted candidates
    unsigned char* sup;    // [B][A]
    int* ncand;            // [B]
    int nc, A, NP, max_det, max_nms;
    float conf_thres, iou_thres, max_wh;
    float* out;            // [B][max_det][7]  x, y, w, h, conf, cls, angle
    int* out_count;        // [B]
    int* out_idx;          // [B][max_det]
};

// best class per anchor (first maximum wins, like torch.max - nms.py:133)
__global__ void __launch_bounds__(256) rnms_best_class_kernel(const float* __restrict__ pred, int B, int nc, int A,
                                                              float* __restrict__ conf, int* __restrict__ cls) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * A) return;
    const int b = (int)(i / A), an = (int)(i % A);
    const float* p = pred + ((long long)b * (5 + nc) + 4) * A + an;
    float best = p[0];
    int bi = 0;
    for (int c = 1; c < nc; ++c) {
        const float v = p[(long long)c * A];
        if (v > best) { best = v; bi = c; }
    }
    conf[i] = best;
    cls[i] = bi;
}

__global__ void __launch_bounds__(1024) rnms_sort_kernel(const RnmsArgs a) {
    __shared__ int cnt[1024];
    __shared__ int n_sh;
    const int b = (int)blockIdx.x, tid = (int)threadIdx.x, nthr = (int)blockDim.x;
    const int A = a.A, NP = a.NP;
    const float* pb = a.pred + (long long)b * (5 + a.nc) * A;
    const float* cb = a.conf + (long long)b * A;
    const int* kb = a.cls + (long long)b * A;
    unsigned long long* keys = a.keys + (long long)b * NP;
    {
        int c = 0;
        for (int i = tid; i < NP; i += nthr) {
            const bool ok = i < A && cb[i] > a.conf_thres;           // nms.py:76,134
            keys[i] = ok ? (((unsigned long long)pp::f2key(cb[i]) << 32) | (unsigned)(0x7fffffff - i)) : 0ull;
            c += ok ? 1 : 0;
        }
        cnt[tid] = c;
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int t = 0; t < nthr; ++t) n += cnt[t];
        n_sh = n > a.max_nms ? a.max_nms : n;                         // nms.py:142-146
        a.ncand[b] = n_sh;
    }
    __syncthreads();
    // descending by (score, -anchor): argsort(descending=True) with ties towards the lower anchor
    for (int size = 2; size <= NP; size <<= 1)
        for (int strd = size >> 1; strd > 0; strd >>= 1) {
            for (int i = tid; i < (NP >> 1); i += nthr) {
                const int lo = 2 * i - (i & (strd - 1)), hi = lo + strd;
                const bool desc = (lo & size) == 0;
                const unsigned long long x0 = keys[lo], x1 = keys[hi];
                if ((x0 < x1) == desc) { keys[lo] = x1; keys[hi] = x0; }
            }
            __syncthreads();
        }
    const int n = n_sh;
    pp::RBox* rb = a.rbox + (long long)b * A;
    unsigned char* sup = a.sup + (long long)b * A;
    for (int i = tid; i < n; i += nthr) {
        const int an = 0x7fffffff - (int)(keys[i] & 0xffffffffull);
        const float off = __fmul_rn((float)kb[an], a.max_wh);         // x[:, 5:6] * max_wh (nms.py:148)
        rb[i] = pp::make_rbox(pb[an], pb[(long long)A + an], pb[2ll * A + an], pb[3ll * A + an], pb[(long long)(4 + a.nc) * A + an],
                              off, an, pp::SinCosF());
        sup[i] = 0;
    }
}

__global__ void __launch_bounds__(RN_TILE) rnms_suppress_kernel(const RnmsArgs a) {
    __shared__ pp::RBox tile[RN_TILE];
    const int b = (int)blockIdx.y, tid = (int)threadIdx.x;
    const int n = a.ncand[b];
    const int j0 = (int)blockIdx.x * RN_TILE;
    if (j0 >= n) return;                                              // uniform per CTA
    const pp::RBox* rb = a.rbox + (long long)b * a.A;
    const int j = j0 + tid;
    const bool live = j < n;
    pp::RBox me;
    if (live) me = rb[j];
    bool dead = false;
    const int jmax = (j0 + RN_TILE < n ? j0 + RN_TILE : n) - 1;       // highest j of this CTA: needs i in [0, jmax)
    for (int i0 = 0; i0 < jmax; i0 += RN_TILE) {
        if (i0 + tid < n) tile[tid] = rb[i0 + tid];
        __syncthreads();
        if (live && !dead) {
            const int iend = (j - i0) < RN_TILE ? (j - i0) : RN_TILE; // i < j only (ious.triu_(diagonal=1), nms.py:228)
            for (int t = 0; t < iend; ++t)
                if (pp::probiou(tile[t], me, pp::LogF(), pp::ExpF()) >= a.iou_thres) { dead = true; break; }
        }
        __syncthreads();
    }
    if (live && dead) a.sup[(long long)b * a.A + j] = 1;
}

__global__ void __launch_bounds__(1024) rnms_emit_kernel(const RnmsArgs a) {
    __shared__ int cnt[1024];
    __shared__ int total;
    const int b = (int)blockIdx.x, tid = (int)threadIdx.x, nthr = (int)blockDim.x;
    const int n = a.ncand[b], A = a.A;
    const unsigned char* sup = a.sup + (long long)b * A;
    const pp::RBox* rb = a.rbox + (long long)b * A;
    const int per = (n + nthr - 1) / nthr;
    const int s0 = tid * per, s1 = (s0 + per < n) ? s0 + per : n;    // contiguous chunk: the survivors keep the score order
    int c = 0;
    for (int i = s0; i < s1; ++i) c += sup[i] ? 0 : 1;
    cnt[tid] = c;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int t = 0; t < nthr; ++t) {
            const int v = cnt[t];
            cnt[t] = run;
            run += v;
        }
        total = run;
    }
    __syncthreads();
    const float* pb = a.pred + (long long)b * (5 + a.nc) * A;
    float* ob = a.out + (long long)b * a.max_det * 7;
    int* ib = a.out_idx + (long long)b * a.max_det;
    int s = cnt[tid];
    for (int i = s0; i < s1 && s < a.max_det; ++i) {                  // i = i[:max_det] (nms.py:160)
        if (sup[i]) continue;
        const int an = rb[i].anchor;
        ob[s * 7 + 0] = pb[an];
        ob[s * 7 + 1] = pb[(long long)A + an];
        ob[s * 7 + 2] = pb[2ll * A + an];
        ob[s * 7 + 3] = pb[3ll * A + an];
        ob[s * 7 + 4] = a.conf[(long long)b * A + an];
        ob[s * 7 + 5] = (float)a.cls[(long long)b * A + an];
        ob[s * 7 + 6] = pb[(long long)(4 + a.nc) * A + an];
        ib[s] = an;
        ++s;
    }
    if (tid == 0) a.out_count[b] = total < a.max_det ? total : a.max_det;
}

static int pp_next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

}  // namespace ym

using namespace ym;

// fp32 logit planes of ym_process_mask: n * mh * mw floats (+ alignment slack)
extern "C" long long ym_process_mask_scratch_bytes(int n, int mh, int mw) {
    return (long long)(n > 0 ? n : 0) * mh * mw * 4 + 256;
}

// ops.process_mask(protos, masks_in, bboxes, shape, upsample) (ultralytics/utils/ops.py:500-528) for the detections of ONE image.
//   protos      [nm][mh][mw] fp16 (proto_dtype 1) or fp32 (2), CHW like the reference's `protos`
//   dets        fp32 rows of pitch ld: xyxy box in `shape` = (in_h, in_w) coordinates at columns 0..3, the nm mask coefficients at
//               columns coef_col .. coef_col + nm (the NMS output rows: coef_col = 6)
//   out         uint8 [n][in_h][in_w] when upsample, else [n][mh][mw]
extern "C" int ym_process_mask(const void* protos, int proto_dtype, int nm, int mh, int mw, const float* dets, int ld, int n,
                               int coef_col, int in_h, int in_w, int upsample, unsigned char* out, void* scratch, void* stream) {
    YM_CHECK_ARG(n >= 0 && n <= 65535 * ML_DETS, "ym_process_mask: 0 <= n <= %d (got %d)", 65535 * ML_DETS, n);
    if (n == 0) return YM_OK;                                         // ops.py:515-516: an empty (0, h, w) stack
    YM_CHECK_ARG(protos && dets && out && scratch, "ym_process_mask: null pointer");
    YM_CHECK_ARG(proto_dtype == 1 || proto_dtype == 2, "ym_process_mask: proto_dtype 1 (fp16) or 2 (fp32)");
    YM_CHECK_ARG(nm >= 1 && nm <= ML_MAX_NM, "ym_process_mask: 1 <= nm <= %d (got %d)", ML_MAX_NM, nm);
    YM_CHECK_ARG(mh >= 1 && mw >= 1 && in_h >= 1 && in_w >= 1 && (long long)mh * mw <= (1ll << 30), "ym_process_mask: sizes");
    YM_CHECK_ARG(coef_col >= 4 && ld >= coef_col + nm, "ym_process_mask: row pitch %d < coef_col %d + nm %d", ld, coef_col, nm);
    cudaStream_t st = (cudaStream_t)stream;
    float* logits = (float*)(((uintptr_t)scratch + 15) & ~(uintptr_t)15);
    const int P = mh * mw;
    dim3 g1((unsigned)((P + 255) / 256), (unsigned)((n + ML_DETS - 1) / ML_DETS));
    if (proto_dtype == 1) {
        YM_LAUNCH(mask_logits_kernel<__half>, g1, 256, 0, st, (const __half*)protos, nm, P, dets, ld, n, coef_col, logits);
    } else {
        YM_LAUNCH(mask_logits_kernel<float>, g1, 256, 0, st, (const float*)protos, nm, P, dets, ld, n, coef_col, logits);
    }
    YM_CHECK_LAUNCH("mask_logits");
    MaskArgs a;
    a.logits = logits; a.dets = dets; a.ld = ld; a.n = n; a.mh = mh; a.mw = mw; a.upsample = upsample ? 1 : 0;
    a.oh = upsample ? in_h : mh;
    a.ow = upsample ? in_w : mw;
    a.sx = (float)mw / (float)in_w;                                   // area_pixel_compute_scale<float> / width_ratio (ops.py:523)
    a.sy = (float)mh / (float)in_h;
    a.out = out;
    const long long total = (long long)n * a.oh * ((a.ow + 3) >> 2);
    long long blocks = (total + 255) / 256;
    if (blocks > 148ll * 64) blocks = 148ll * 64;                     // grid-stride beyond 64 CTAs per SM
    YM_LAUNCH(mask_finish_kernel, (unsigned)blocks, 256, 0, st, a);
    YM_CHECK_LAUNCH("mask_finish");
    return YM_OK;
}

extern "C" long long ym_nms_rotated_scratch_bytes(int B, int A) {
    const long long NP = pp_next_pow2(A);
    return (long long)B * A * (4 + 4 + (long long)sizeof(pp::RBox) + 1) + (long long)B * NP * 8 + (long long)B * 4 + 256;
}

// non_max_suppression(..., rotated=True) (ultralytics/utils/nms.py:13-171) for a batch: pred fp32 [B][4 + nc + 1][A] = xywh, class
// scores, angle (the OBB head's eval output).  out fp32 [B][max_det][7] = x, y, w, h, conf, cls, angle in score order,
// out_count int32 [B], out_idx int32 [B][max_det] (anchor of each kept row; -1 past the count).
extern "C" int ym_nms_rotated(const float* pred, int B, int nc, int A, float conf_thres, float iou_thres, int max_det, int max_nms,
                              float max_wh, float* out, int* out_count, int* out_idx, void* scratch, void* stream) {
    YM_CHECK_ARG(pred && out && out_count && out_idx && scratch, "ym_nms_rotated: null pointer");
    YM_CHECK_ARG(max_det >= 1 && max_nms >= 1, "ym_nms_rotated: max_det, max_nms >= 1");
    YM_CHECK_ARG(B >= 0 && B <= 65535 && nc >= 1 && A >= 1 && A <= (1 << 24), "ym_nms_rotated: sizes");
    if (B == 0) return YM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int NP = pp_next_pow2(A);
    unsigned char* p = (unsigned char*)(((uintptr_t)scratch + 15) & ~(uintptr_t)15);
    RnmsArgs a;
    a.keys = (unsigned long long*)p;   p += (size_t)B * NP * 8;
    a.rbox = (pp::RBox*)p;             p += (size_t)B * A * sizeof(pp::RBox);
    a.conf = (float*)p;                p += (size_t)B * A * 4;
    a.cls = (int*)p;                   p += (size_t)B * A * 4;
    a.ncand = (int*)p;                 p += (size_t)B * 4;
    a.sup = p;
    a.pred = pred; a.nc = nc; a.A = A; a.NP = NP; a.max_det = max_det; a.max_nms = max_nms; a.conf_thres = conf_thres;
    a.iou_thres = iou_thres; a.max_wh = max_wh; a.out = out; a.out_count = out_count; a.out_idx = out_idx;
    cudaMemsetAsync(out, 0, (size_t)B * max_det * 7 * sizeof(float), st);
    cudaMemsetAsync(out_idx, 0xff, (size_t)B * max_det * sizeof(int), st);
    const long long total = (long long)B * A;
    YM_LAUNCH(rnms_best_class_kernel, (unsigned)((total + 255) / 256), 256, 0, st, pred, B, nc, A, a.conf, a.cls);
    YM_CHECK_LAUNCH("rnms_best_class");
    YM_LAUNCH(rnms_sort_kernel, (unsigned)B, 1024, 0, st, a);
    YM_CHECK_LAUNCH("rnms_sort");
    const int cap = A < max_nms ? A : max_nms;
    YM_LAUNCH(rnms_suppress_kernel, dim3((unsigned)((cap + RN_TILE - 1) / RN_TILE), (unsigned)B), RN_TILE, 0, st, a);
    YM_CHECK_LAUNCH("rnms_suppress");
    YM_LAUNCH(rnms_emit_kernel, (unsigned)B, 1024, 0, st, a);
    YM_CHECK_LAUNCH("rnms_emit");
    return YM_OK;
}
