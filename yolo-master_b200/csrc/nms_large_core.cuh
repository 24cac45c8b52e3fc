// Large-candidate path of the batched NMS (mode 0): images with more candidates above the confidence threshold than the
// shared-memory sorter of nms.cu holds (16384) - validation at conf 0.001 produces up to max_nms = 30000 (utils/nms.py:142-146).
// Same decisions as the shared-memory path (and as the oracle): best-class confidence > conf_thres, score-descending order with
// ties towards the lower anchor, the first max_nms candidates, class offset added in fp32, greedy suppression of IoU > iou_thres,
// first max_det survivors.  Keys, boxes and flags live in global scratch; one CTA per image.
//
// The per-image algorithm is ONE function, written against an executor: `ex.all(f)` runs f(tid) for every thread of the CTA and
// then synchronises.  On the device that is "call f(threadIdx.x); __syncthreads()", on the host (tests/native/nms_large_host.cpp,
// g++) a loop over tid - so the code below is compared with the NMS oracle in the GPU-less build container.  Scalar control flow
// between the `ex.all` calls only reads what an earlier step left in the shared block, so it is uniform across the CTA.
#pragma once
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

// Single IEEE roundings, never contracted into FMAs (the translation unit is built with --use_fast_math); plain operators on the
// host, where the baseline x86-64 target has no FMA to contract into.
#ifdef __CUDA_ARCH__
#define YM_RN_ADD(a, b) __fadd_rn(a, b)
#define YM_RN_SUB(a, b) __fsub_rn(a, b)
#define YM_RN_MUL(a, b) __fmul_rn(a, b)
#define YM_RN_DIV(a, b) __fdiv_rn(a, b)
#else
#define YM_RN_ADD(a, b) ((a) + (b))
#define YM_RN_SUB(a, b) ((a) - (b))
#define YM_RN_MUL(a, b) ((a) * (b))
#define YM_RN_DIV(a, b) ((a) / (b))
#endif

namespace ym {
namespace nmsl {

constexpr int MAX_KEEP = 512;

struct Box4 {
    float x1, y1, x2, y2;
};

struct Shared {            // one per CTA
    int n, cur, nk, from;
    int keep[MAX_KEEP];
    int cnt[1024];         // per-thread candidate counts
};

struct Args {
    const float* pred;     // [B][4+nc][A]  xywh + class scores
    const float* conf;     // [B][A] best-class confidence
    const int* cls;        // [B][A] best class
    int nc, A, NP;         // NP = power of two >= A
    float conf_thres, iou_thres, max_wh;
    int max_det, max_nms;
    unsigned long long* keys;   // [B][NP]
    Box4* sbox;                 // [B][A]
    unsigned char* sup;         // [B][A]
    float* out;                 // [B][max_det][6]
    int* out_count;             // [B]
    int* out_idx;               // [B][max_det]
};

YM_HD uint32_t f2key(float f) {
    union { float f; uint32_t u; } v;
    v.f = f;
    return (v.u & 0x80000000u) ? ~v.u : (v.u | 0x80000000u);
}
YM_HD float fmaxf_(float a, float b) { return a > b ? a : b; }
YM_HD float fminf_(float a, float b) { return a < b ? a : b; }

YM_HD Box4 offset_box(const Box4& b, int c, float max_wh) {
    const float off = YM_RN_MUL((float)c, max_wh);                    // x[:, 5:6] * max_wh, then boxes + c (two roundings)
    Box4 r;
    r.x1 = YM_RN_ADD(b.x1, off); r.y1 = YM_RN_ADD(b.y1, off); r.x2 = YM_RN_ADD(b.x2, off); r.y2 = YM_RN_ADD(b.y2, off);
    return r;
}
YM_HD float iou(const Box4& p, const Box4& q) {                       // TorchNMS.nms utils/nms.py:279-291
    const float w = fmaxf_(YM_RN_SUB(fminf_(p.x2, q.x2), fmaxf_(p.x1, q.x1)), 0.f);
    const float h = fmaxf_(YM_RN_SUB(fminf_(p.y2, q.y2), fmaxf_(p.y1, q.y1)), 0.f);
    const float inter = YM_RN_MUL(w, h);
    const float ap = YM_RN_MUL(YM_RN_SUB(p.x2, p.x1), YM_RN_SUB(p.y2, p.y1));
    const float aq = YM_RN_MUL(YM_RN_SUB(q.x2, q.x1), YM_RN_SUB(q.y2, q.y1));
    return YM_RN_DIV(inter, YM_RN_SUB(YM_RN_ADD(ap, aq), inter));
}

template <class Exec>
YM_HD void image(const Args& a, int b, Exec& ex, Shared& sh) {
    const int A = a.A, NP = a.NP, nthr = ex.nthr;
    const float* pb = a.pred + (long long)b * (4 + a.nc) * A;
    const float* cb = a.conf + (long long)b * A;
    const int* kb = a.cls + (long long)b * A;
    unsigned long long* keys = a.keys + (long long)b * NP;
    Box4* sbox = a.sbox + (long long)b * A;
    unsigned char* sup = a.sup + (long long)b * A;

    // keys for every anchor (0 = not a candidate): no compaction needed, the sort moves the candidates to the front
    ex.all([&](int tid) {
        int c = 0;
        for (int i = tid; i < NP; i += nthr) {
            const bool ok = i < A && cb[i] > a.conf_thres;
            keys[i] = ok ? (((unsigned long long)f2key(cb[i]) << 32) | (unsigned)(0x7fffffff - i)) : 0ull;
            c += ok ? 1 : 0;
        }
        sh.cnt[tid] = c;
    });
    ex.all([&](int tid) {
        if (tid != 0) return;
        int n = 0;
        for (int t = 0; t < nthr; ++t) n += sh.cnt[t];
        sh.n = n > a.max_nms ? a.max_nms : n;                         // nms.py:142-146
        sh.nk = 0;
        sh.from = 0;
        sh.cur = 0;
    });
    // bitonic sort, descending by (score, -anchor): ties resolve towards the lower anchor index
    for (int size = 2; size <= NP; size <<= 1)
        for (int strd = size >> 1; strd > 0; strd >>= 1)
            ex.all([&](int tid) {
                for (int i = tid; i < (NP >> 1); i += nthr) {
                    const int lo = 2 * i - (i & (strd - 1)), hi = lo + strd;
                    const bool desc = (lo & size) == 0;
                    const unsigned long long x0 = keys[lo], x1 = keys[hi];
                    if ((x0 < x1) == desc) { keys[lo] = x1; keys[hi] = x0; }
                }
            });
    const int n = sh.n;
    ex.all([&](int tid) {                                             // sorted xyxy boxes (xywh2xyxy: xy -+ wh/2), flags cleared
        for (int i = tid; i < n; i += nthr) {
            const int an = 0x7fffffff - (int)(keys[i] & 0xffffffffull);
            const float cx = pb[an], cy = pb[(long long)A + an], w = pb[2ll * A + an], h = pb[3ll * A + an];
            const float hw = YM_RN_DIV(w, 2.f), hh = YM_RN_DIV(h, 2.f);
            Box4 bx;
            bx.x1 = YM_RN_SUB(cx, hw); bx.y1 = YM_RN_SUB(cy, hh); bx.x2 = YM_RN_ADD(cx, hw); bx.y2 = YM_RN_ADD(cy, hh);
            sbox[i] = bx;
            sup[i] = 0;
        }
    });
    // greedy sweep: at most max_det survivors are ever needed (output is keep[:max_det], nms.py:160)
    while (true) {
        ex.all([&](int tid) {
            if (tid != 0) return;
            int cur = sh.from;
            while (cur < n && sup[cur]) ++cur;
            sh.cur = cur;
            if (cur < n && sh.nk < MAX_KEEP) sh.keep[sh.nk++] = cur;
        });
        const int i = sh.cur;
        if (i >= n || sh.nk >= a.max_det) break;
        ex.all([&](int tid) {
            const int ai = 0x7fffffff - (int)(keys[i] & 0xffffffffull);
            const Box4 bi = offset_box(sbox[i], kb[ai], a.max_wh);
            for (int j = i + 1 + tid; j < n; j += nthr) {
                if (sup[j]) continue;
                const int aj = 0x7fffffff - (int)(keys[j] & 0xffffffffull);
                if (iou(bi, offset_box(sbox[j], kb[aj], a.max_wh)) > a.iou_thres) sup[j] = 1;
            }
            if (tid == 0) sh.from = i + 1;
        });
    }
    const int nk = sh.nk < a.max_det ? sh.nk : a.max_det;
    ex.all([&](int tid) {
        float* ob = a.out + (long long)b * a.max_det * 6;
        int* ib = a.out_idx + (long long)b * a.max_det;
        for (int s = tid; s < nk; s += nthr) {
            const int i = sh.keep[s];
            const int an = 0x7fffffff - (int)(keys[i] & 0xffffffffull);
            const Box4 bx = sbox[i];
            ob[s * 6 + 0] = bx.x1; ob[s * 6 + 1] = bx.y1; ob[s * 6 + 2] = bx.x2; ob[s * 6 + 3] = bx.y2;
            ob[s * 6 + 4] = cb[an]; ob[s * 6 + 5] = (float)kb[an];
            ib[s] = an;
        }
        if (tid == 0) a.out_count[b] = nk;
    });
}

}  // namespace nmsl
}  // namespace ym
