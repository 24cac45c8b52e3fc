// Batched NMS and Cluster-Weighted NMS on device (sm_100a), one CTA per image, no host round trips.
//
// mode 0  ultralytics `non_max_suppression` (utils/nms.py:13-171, single-label class-aware branch) with the greedy kernel of
//         `TorchNMS.nms` (:245-302): candidates with best-class conf > conf_thres, class offset cls*max_wh ADDED IN FP32
//         (so the reference's own coordinate quantisation at large offsets is reproduced and keep decisions are bit-exact),
//         score-descending greedy suppression of IoU > iou_thres, first max_det survivors, boxes returned as xyxy.
// mode 1  CW-NMS of the C++ deployment code (examples/.../cpp/src/common.cpp:56-198): conf >= conf_thres, offset
//         2*max(w,h)+8192 and IoU in float64, greedy survivors' boxes replaced by the score-and-proximity weighted mean of
//         their cluster, w = s*exp(-(1-IoU)^2/sigma) over the top-3000 pool, guard sum_w > 1e-6, clip to frame, drop empty.
//
// Replaces a Python per-image loop around a third-party NMS kernel (torchvision) / a `while` loop with one sync per kept
// box.  Pipeline per image: (1) wide kernel: best class + confidence per anchor (coalesced over anchors);
// (2) CTA: compaction -> bitonic sort of (score, index) keys in shared memory -> greedy sweep with the suppression flags in
// shared memory (<= max_det iterations, each a parallel IoU pass) -> optional cluster refinement (warp per survivor).
#include <type_traits>

#include "ym_common.cuh"

namespace ym {

constexpr int NMS_CAP = 16384;   // candidates per image held in shared memory (keys 8 B + flags 1 B)

__device__ __forceinline__ uint32_t nms_f2key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// (1) best class per anchor.  pred: [B][4+nc][A]
__global__ void __launch_bounds__(256) nms_best_class_kernel(const float* __restrict__ pred, int B, int nc, int A,
                                                             float* __restrict__ conf, int* __restrict__ cls) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * A) return;
    const int b = (int)(i / A), a = (int)(i % A);
    const float* p = pred + ((long long)b * (4 + nc) + 4) * A + a;
    float best = p[0];
    int bi = 0;
    for (int c = 1; c < nc; ++c) {
        const float v = p[(long long)c * A];
        if (v > best) { best = v; bi = c; }      // first maximum wins, like torch.max
    }
    conf[i] = best;
    cls[i] = bi;
}

template <typename T>
struct BoxT { T x1, y1, x2, y2; };

template <int MODE>
__global__ void __launch_bounds__(1024) nms_image_kernel(const float* __restrict__ pred, int nc, int A,
                                                         const float* __restrict__ conf, const int* __restrict__ cls,
                                                         float conf_thres, float iou_thres, int max_det, int max_nms, float max_wh,
                                                         float sigma, float frame_w, float frame_h, float* __restrict__ out,
                                                         int* __restrict__ out_count, int* __restrict__ out_idx,
                                                         float4* __restrict__ sbox_g /* [B][NMS_CAP] sorted boxes */,
                                                         int* __restrict__ err) {
    using T = typename std::conditional<MODE == 1, double, float>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);   // [NMS_CAP]
    unsigned char* sup = reinterpret_cast<unsigned char*>(keys + NMS_CAP);         // [NMS_CAP]
    __shared__ int s_n, s_cur, s_nk;
    __shared__ int s_keep[512];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* pb = pred + (long long)b * (4 + nc) * A;
    const float* cb = conf + (long long)b * A;
    const int* kb = cls + (long long)b * A;
    float4* sbox = sbox_g + (long long)b * NMS_CAP;

    // ---- compaction (anchor order is irrelevant after the sort; the key carries the anchor index for tie-breaking)
    if (tid == 0) { s_n = 0; s_nk = 0; s_cur = 0; }
    __syncthreads();
    for (int a = tid; a < A; a += blockDim.x) {
        const float c = cb[a];
        const bool ok = (MODE == 1) ? !(c < conf_thres) : (c > conf_thres);
        if (ok) {
            const int slot = atomicAdd(&s_n, 1);
            if (slot < NMS_CAP) keys[slot] = ((unsigned long long)nms_f2key(c) << 32) | (unsigned)(0x7fffffff - a);
        }
    }
    __syncthreads();
    int n = s_n;
    if (n > NMS_CAP) {   // more candidates than the shared-memory sorter holds: report instead of silently truncating
        if (tid == 0) { atomicExch(err, 1); out_count[b] = 0; }
        return;
    }
    int npow = 1;
    while (npow < n) npow <<= 1;
    for (int i = n + tid; i < npow; i += blockDim.x) keys[i] = 0ull;
    __syncthreads();
    // ---- bitonic sort, descending by (score, -anchor): ties resolve towards the lower anchor index
    for (int size = 2; size <= npow; size <<= 1) {
        for (int strd = size >> 1; strd > 0; strd >>= 1) {
            for (int i = tid; i < (npow >> 1); i += blockDim.x) {
                const int lo = 2 * i - (i & (strd - 1));
                const int hi = lo + strd;
                const bool desc = ((lo & size) == 0);
                const unsigned long long x0 = keys[lo], x1 = keys[hi];
                if ((x0 < x1) == desc) { keys[lo] = x1; keys[hi] = x0; }
            }
            __syncthreads();
        }
    }
    if (MODE == 0 && n > max_nms) n = max_nms;       // nms.py:142-146
    // ---- sorted boxes -> global scratch (xyxy for mode 0, x,y,w,h for mode 1), flags cleared
    for (int i = tid; i < n; i += blockDim.x) {
        const int a = 0x7fffffff - (int)(keys[i] & 0xffffffffull);
        const float cx = pb[a], cy = pb[(long long)A + a], w = pb[2ll * A + a], h = pb[3ll * A + a];
        float4 bx;
        if (MODE == 0) {   // xywh2xyxy: xy -+ wh/2
            const float hw = w / 2, hh = h / 2;
            bx = make_float4(cx - hw, cy - hh, cx + hw, cy + hh);
        } else {           // decode_candidates common.cpp:106-109 (pad 0, scale 1)
            bx = make_float4(cx - 0.5f * w, cy - 0.5f * h, w, h);
        }
        sbox[i] = bx;
        sup[i] = 0;
    }
    __syncthreads();

    auto offbox = [&](int i) -> BoxT<T> {
        const float4 bx = sbox[i];
        const int a = 0x7fffffff - (int)(keys[i] & 0xffffffffull);
        const int c = kb[a];
        BoxT<T> r;
        if constexpr (MODE == 0) {
            const float off = __fmul_rn((float)c, max_wh);            // x[:, 5:6] * max_wh, then boxes + c (two roundings)
            r.x1 = __fadd_rn(bx.x, off); r.y1 = __fadd_rn(bx.y, off); r.x2 = __fadd_rn(bx.z, off); r.y2 = __fadd_rn(bx.w, off);
        } else {
            const double OFF = 2.0 * fmax((double)frame_w, (double)frame_h) + 8192.0;
            r.x1 = (double)bx.x + c * OFF; r.y1 = (double)bx.y + c * OFF; r.x2 = r.x1 + (double)bx.z; r.y2 = r.y1 + (double)bx.w;
        }
        return r;
    };
    auto iou = [&](const BoxT<T>& p, const BoxT<T>& q) -> T {
        if constexpr (MODE == 0) {   // TorchNMS.nms :279-291
            const float w = fmaxf(__fsub_rn(fminf(p.x2, q.x2), fmaxf(p.x1, q.x1)), 0.f);
            const float h = fmaxf(__fsub_rn(fminf(p.y2, q.y2), fmaxf(p.y1, q.y1)), 0.f);
            const float inter = __fmul_rn(w, h);
            const float ap = __fmul_rn(__fsub_rn(p.x2, p.x1), __fsub_rn(p.y2, p.y1));
            const float aq = __fmul_rn(__fsub_rn(q.x2, q.x1), __fsub_rn(q.y2, q.y1));
            return __fdiv_rn(inter, __fsub_rn(__fadd_rn(ap, aq), inter));
        } else {           // box_iou common.cpp:56-67
            const double w = fmax(0.0, fmin((double)p.x2, (double)q.x2) - fmax((double)p.x1, (double)q.x1));
            const double h = fmax(0.0, fmin((double)p.y2, (double)q.y2) - fmax((double)p.y1, (double)q.y1));
            const double inter = w * h;
            const double uni = (p.x2 - p.x1) * (p.y2 - p.y1) + (q.x2 - q.x1) * (q.y2 - q.y1) - inter;
            return uni > 0 ? inter / uni : 0.0;
        }
    };

    // ---- greedy sweep: at most max_det survivors are ever needed (output is keep[:max_det], nms.py:160)
    int search_from = 0;
    while (true) {
        if (tid < 32) {   // next unsuppressed candidate at or after search_from (= previous survivor + 1)
            int cur = search_from, found = -1;
            while (cur < n) {
                const int j = cur + tid;
                const unsigned m = __ballot_sync(0xffffffffu, j < n && !sup[j]);
                if (m) { found = cur + __ffs(m) - 1; break; }
                cur += 32;
            }
            if (tid == 0) {
                s_cur = found < 0 ? n : found;
                if (found >= 0 && s_nk < 512) s_keep[s_nk++] = found;
            }
        }
        __syncthreads();
        const int i = s_cur, nk = s_nk;
        // mode 1 (CW-NMS): survivors that clip to nothing are dropped AFTER the sweep and do not count towards max_det
        // (common.cpp:180-197), so the sweep collects up to the 512-entry survivor table instead of stopping at max_det
        if (i >= n || nk >= (MODE == 1 ? 512 : max_det)) break;
        const BoxT<T> bi = offbox(i);
        for (int j = i + 1 + tid; j < n; j += blockDim.x) {
            if (!sup[j] && iou(bi, offbox(j)) > (T)iou_thres) sup[j] = 1;
        }
        search_from = i + 1;
        __syncthreads();
    }
    __syncthreads();
    const int nk_cap = MODE == 1 ? 512 : max_det;
    const int nk = s_nk < nk_cap ? s_nk : nk_cap;

    // ---- emit (mode 1: cluster-weighted refinement, one warp per survivor, then clip / drop empty in survivor order)
    float* ob = out + (long long)b * max_det * 6;
    int* ib = out_idx + (long long)b * max_det;
    if (MODE == 0) {
        for (int s = tid; s < nk; s += blockDim.x) {
            const int i = s_keep[s];
            const int a = 0x7fffffff - (int)(keys[i] & 0xffffffffull);
            const float4 bx = sbox[i];
            ob[s * 6 + 0] = bx.x; ob[s * 6 + 1] = bx.y; ob[s * 6 + 2] = bx.z; ob[s * 6 + 3] = bx.w;
            ob[s * 6 + 4] = cb[a]; ob[s * 6 + 5] = (float)kb[a];
            ib[s] = a;
        }
        if (tid == 0) out_count[b] = nk;
    } else {
        double* refined = reinterpret_cast<double*>(keys + NMS_CAP) + (NMS_CAP / 8 + 8);   // after `sup`, 8-byte aligned: [512][4]
        const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
        const int pool = n < 3000 ? n : 3000;                                              // common.cpp:153-158
        for (int s = warp; s < nk; s += nwarp) {
            const int k = s_keep[s];
            const BoxT<T> bk = offbox(k);
            double sw = 0, ax = 0, ay = 0, ax2 = 0, ay2 = 0;
            for (int m = lane; m < pool; m += 32) {
                const double ov = iou(bk, offbox(m));
                if (ov <= (double)iou_thres) continue;
                const int am = 0x7fffffff - (int)(keys[m] & 0xffffffffull);
                const double w = (double)cb[am] * exp(-((1.0 - ov) * (1.0 - ov)) / (double)sigma);
                const float4 bm = sbox[m];
                sw += w; ax += w * (double)bm.x; ay += w * (double)bm.y;
                ax2 += w * ((double)bm.x + (double)bm.z); ay2 += w * ((double)bm.y + (double)bm.w);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sw += __shfl_xor_sync(0xffffffffu, sw, o); ax += __shfl_xor_sync(0xffffffffu, ax, o);
                ay += __shfl_xor_sync(0xffffffffu, ay, o); ax2 += __shfl_xor_sync(0xffffffffu, ax2, o);
                ay2 += __shfl_xor_sync(0xffffffffu, ay2, o);
            }
            if (lane == 0) {
                const float4 bx = sbox[k];
                double x0 = bx.x, y0 = bx.y, w = bx.z, h = bx.w;
                if (sigma > 0.f && sw > 1e-6) { x0 = ax / sw; y0 = ay / sw; w = fmax(0.0, ax2 / sw - x0); h = fmax(0.0, ay2 / sw - y0); }
                refined[s * 4 + 0] = x0; refined[s * 4 + 1] = y0; refined[s * 4 + 2] = w; refined[s * 4 + 3] = h;
            }
        }
        __syncthreads();
        if (tid == 0) {   // clip to frame, drop empty boxes, keep survivor order (common.cpp:180-197)
            int m = 0;
            for (int s = 0; s < nk && m < max_det; ++s) {
                const double x0 = fmax(refined[s * 4], 0.0), y0 = fmax(refined[s * 4 + 1], 0.0);
                const double x1 = fmin(refined[s * 4] + refined[s * 4 + 2], (double)frame_w);
                const double y1 = fmin(refined[s * 4 + 1] + refined[s * 4 + 3], (double)frame_h);
                if (x1 - x0 > 0 && y1 - y0 > 0) {
                    const int i = s_keep[s];
                    const int a = 0x7fffffff - (int)(keys[i] & 0xffffffffull);
                    ob[m * 6 + 0] = (float)x0; ob[m * 6 + 1] = (float)y0; ob[m * 6 + 2] = (float)(x1 - x0); ob[m * 6 + 3] = (float)(y1 - y0);
                    ob[m * 6 + 4] = cb[a]; ob[m * 6 + 5] = (float)kb[a];
                    ib[m] = a;
                    ++m;
                }
            }
            out_count[b] = m;
        }
    }
}

}  // namespace ym

using namespace ym;

extern "C" long long ym_nms_scratch_bytes(int B, int A) {
    return (long long)B * A * 8 + (long long)B * NMS_CAP * 16 + 64;
}

// pred fp32 [B][4+nc][A] (xywh centre boxes + class scores).  out fp32 [B][max_det][6], out_count int32 [B],
// out_idx int32 [B][max_det] (anchor index of every output row).  mode 0: out rows = (x1,y1,x2,y2,conf,cls);
// mode 1 (CW-NMS): out rows = (x,y,w,h,conf,cls) clipped to the frame.  Returns non-zero (and sets the error) if an image
// has more than 16384 candidates above the confidence threshold.
extern "C" int ym_nms_batched(const float* pred, int B, int nc, int A, float conf_thres, float iou_thres, int max_det, int max_nms,
                              float max_wh, int mode, float sigma, float frame_w, float frame_h, float* out, int* out_count,
                              int* out_idx, void* scratch, void* stream) {
    YM_CHECK_ARG(pred && out && out_count && out_idx && scratch, "ym_nms_batched: null pointer");
    YM_CHECK_ARG(max_det >= 1 && max_det <= 512, "ym_nms_batched: max_det must be in 1..512");
    YM_CHECK_ARG(mode == 0 || mode == 1, "ym_nms_batched: mode");
    YM_CHECK_ARG(B >= 0 && nc >= 1 && A >= 1, "ym_nms_batched: sizes");
    if (B == 0) return YM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    float* conf = (float*)scratch;
    int* cls = (int*)(conf + (long long)B * A);
    float4* sbox = (float4*)((((uintptr_t)(cls + (long long)B * A)) + 15) & ~(uintptr_t)15);
    int* err = (int*)(sbox + (long long)B * NMS_CAP);   // global overflow flag kept in the scratch tail
    cudaMemsetAsync(err, 0, sizeof(int), st);
    cudaMemsetAsync(out, 0, (size_t)B * max_det * 6 * sizeof(float), st);
    cudaMemsetAsync(out_idx, 0xff, (size_t)B * max_det * sizeof(int), st);
    const long long total = (long long)B * A;
    nms_best_class_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(pred, B, nc, A, conf, cls);
    YM_CHECK_LAUNCH("nms_best_class");
    const size_t smem = (size_t)NMS_CAP * 8 + NMS_CAP + 64 + 512 * 4 * 8;
    if (mode == 0) {
        cudaFuncSetAttribute(nms_image_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        nms_image_kernel<0><<<B, 1024, smem, st>>>(pred, nc, A, conf, cls, conf_thres, iou_thres, max_det, max_nms, max_wh, sigma,
                                                  frame_w, frame_h, out, out_count, out_idx, sbox, err);
    } else {
        cudaFuncSetAttribute(nms_image_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        nms_image_kernel<1><<<B, 1024, smem, st>>>(pred, nc, A, conf, cls, conf_thres, iou_thres, max_det, max_nms, max_wh, sigma,
                                                  frame_w, frame_h, out, out_count, out_idx, sbox, err);
    }
    YM_CHECK_LAUNCH("nms_image");
    return YM_OK;
}

// 1 if the last ym_nms_batched on this scratch buffer overflowed the 16384-candidate sorter (read after a stream sync).
extern "C" int ym_nms_overflowed(const void* scratch, int B, int A, void* stream) {
    const float* conf = (const float*)scratch;
    const int* cls = (const int*)(conf + (long long)B * A);
    const float4* sbox = (const float4*)((((uintptr_t)(cls + (long long)B * A)) + 15) & ~(uintptr_t)15);
    const int* err = (const int*)(sbox + (long long)B * NMS_CAP);
    int h = 0;
    cudaMemcpyAsync(&h, err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    cudaStreamSynchronize((cudaStream_t)stream);
    return h;
}
