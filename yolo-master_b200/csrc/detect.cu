// Detect head post-processing on device (head.py:173-258):
//   anchors/strides (tal.py:398-411) + dist2bbox (tal.py:414-423) + sigmoid + end2end two-stage top-k
//   (max over classes -> top-k anchors -> flatten k x nc -> top-k) in ONE kernel, one CTA per image.
// Also the dense decode (B, 4+nc, A) used by the NMS path (head.py:173-184).
//
// Selection runs on the raw fp32 logits (sigmoid is monotonic), with radix-select thresholds and a final
// 512-wide bitonic sort; scores are emitted as sigmoid(logit).  Ties are broken towards the lower index.
#include "ym_common.cuh"

namespace ym {

struct DetectLevels {
    const float* box[4];  // [B, hw_l, 4*reg_max] fp32 (reg_max == 1)
    const float* cls[4];  // [B, hw_l, nc] fp32 logits
    int h[4], w[4];
    float stride[4];
    int off[5];  // anchor offsets, off[nl] = A
    int nl;
};

__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// Block-wide: k-th largest key among keys[0..n).  Returns T; *take_eq = how many keys == T belong to the top-k.
__device__ uint32_t radix_select_kth(const uint32_t* keys, int n, int k, uint32_t* hist, int* sh_misc, int* take_eq) {
    uint32_t prefix = 0, mask = 0;
    int remaining = k;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        // (a warp-aggregated variant - __match_any_sync + one atomic per distinct digit - was measured: 147 -> 163 us, reverted)
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t u = keys[i];
            if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // descending scan of the 256 bins by one warp: lane l owns bins 255-8l .. 248-8l (a single thread walking the bins
            // was ~30 us of dependent shared-memory reads per image over the 8 digit passes of the two selections)
            const int lane = threadIdx.x;
            int local[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { local[j] = (int)hist[255 - (lane * 8 + j)]; sum += local[j]; }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int excl = incl - sum;
            if (excl < remaining && remaining <= incl) {      // exactly one lane: its bins contain the remaining-th largest key
                int cum = excl, digit = 255 - lane * 8 - 7, left = remaining - excl;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (cum + local[j] >= remaining) { digit = 255 - (lane * 8 + j); left = remaining - cum; break; }
                    cum += local[j];
                }
                sh_misc[0] = digit;
                sh_misc[1] = left;
            }
        }
        __syncthreads();
        prefix |= ((uint32_t)sh_misc[0]) << shift;
        mask |= 255u << shift;
        remaining = sh_misc[1];
        __syncthreads();
    }
    *take_eq = remaining;
    return prefix;
}

// Collect the indices of the top-k keys into list[0..k): all keys > T (any order) then the first take_eq keys == T.
__device__ void collect_topk(const uint32_t* keys, int n, uint32_t T, int take_eq, int* list, int* sh_cnt) {
    if (threadIdx.x == 0) *sh_cnt = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (keys[i] > T) list[atomicAdd(sh_cnt, 1)] = i;
    __syncthreads();
    if (threadIdx.x < 32) {  // ordered scan of the equal keys by one warp
        int base = *sh_cnt, taken = 0;
        for (int i0 = 0; i0 < n && taken < take_eq; i0 += 32) {
            const int i = i0 + threadIdx.x;
            const bool eq = i < n && keys[i] == T;
            const unsigned m = __ballot_sync(0xffffffffu, eq);
            const int rank = __popc(m & ((1u << threadIdx.x) - 1u));
            if (eq && taken + rank < take_eq) list[base + taken + rank] = i;
            taken += __popc(m);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void anchor_of(const DetectLevels& L, int a, int& lvl, int& r) {
    lvl = 0;
    while (lvl + 1 < L.nl && a >= L.off[lvl + 1]) ++lvl;
    r = a - L.off[lvl];
}

// Stage 1, full-GPU: per-anchor max logit over classes -> sortable keys [B][A].
// General shape: one warp per anchor row.
__global__ void __launch_bounds__(256) detect_rowmax_kernel(const DetectLevels L, int nc, int B, uint32_t* __restrict__ keys) {
    pdl_prologue();
    const int A = L.off[L.nl];
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= (long long)B * A) return;
    const int b = (int)(row / A), a = (int)(row % A);
    int lvl, r;
    anchor_of(L, a, lvl, r);
    const float* src = L.cls[lvl] + ((long long)b * L.h[lvl] * L.w[lvl] + r) * nc;
    float m = -INFINITY;
    for (int c = lane; c < nc; c += 32) m = fmaxf(m, src[c]);
    m = warp_max(m);
    if (lane == 0) keys[row] = f2key(m);
}

// nc % 4 == 0 (16-byte rows): a warp takes EIGHT consecutive anchor rows of one (level, image) - 8 * nc floats of contiguous memory - as
// float4 loads dealt lane by lane (nc = 80: five independent, fully coalesced 512-byte requests per warp instead of three 4-byte loads
// per lane behind one another: the row-per-warp kernel moved 1.4 TB/s, profiles/r02_launch_roofline.txt), reduces each float4, and
// eight lanes finish the rows from a warp-private shared-memory strip.
constexpr int RM_ROWS = 8, RM_MAXQ = 64;          // nc <= 256
struct RowmaxGroups { int first[5]; };            // first[l] = index of level l's first 8-row group; first[nl] = total groups
__global__ void __launch_bounds__(256) detect_rowmax8_kernel(const DetectLevels L, const RowmaxGroups G, int nc, int B, uint32_t* __restrict__ keys) {
    pdl_prologue();
    __shared__ float part[8][RM_ROWS * RM_MAXQ];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = blockIdx.x * 8 + warp;
    if (grp >= G.first[L.nl]) return;
    int lvl = 0;
    while (lvl + 1 < L.nl && grp >= G.first[lvl + 1]) ++lvl;
    const int hw = L.h[lvl] * L.w[lvl], gpi = (hw + RM_ROWS - 1) / RM_ROWS;      // groups per image
    const int gl = grp - G.first[lvl];
    const int b = gl / gpi, r0 = (gl - b * gpi) * RM_ROWS;
    const int rows = min(RM_ROWS, hw - r0), q = nc >> 2, total = rows * q;
    const float4* src = reinterpret_cast<const float4*>(L.cls[lvl] + ((long long)b * hw + r0) * nc);
    float* mine = part[warp];
#pragma unroll 4
    for (int i = lane; i < total; i += 32) {
        const float4 v = src[i];
        mine[i] = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    }
    __syncwarp();
    if (lane < rows) {
        float m = -INFINITY;
        for (int c = 0; c < q; ++c) m = fmaxf(m, mine[lane * q + ((c + lane) % q)]);      // rotated start: the 8 lanes hit different banks
        keys[(long long)b * L.off[L.nl] + L.off[lvl] + r0 + lane] = f2key(m);
    }
}

__global__ void __launch_bounds__(1024) detect_topk_kernel(const DetectLevels L, int nc, int kdet, float* __restrict__ out,
                                                           int* __restrict__ out_anchor, const uint32_t* __restrict__ keys_g) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int A = L.off[L.nl];
    const int nkeys_max = max(A, kdet * nc);
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem_raw);              // [max(A, kdet*nc)]
    int* cand = reinterpret_cast<int*>(keys + nkeys_max);               // [kdet] anchor ids
    int* sel = cand + kdet;                                             // [kdet] flat ids
    unsigned long long* sortbuf = reinterpret_cast<unsigned long long*>(sel + kdet + ((nkeys_max + 2 * kdet) & 1));  // [512], 8-byte aligned
    __shared__ uint32_t hist[256];
    __shared__ int misc[4];
    const int b = blockIdx.x, tid = threadIdx.x;

    // ---- stage 1 result (detect_rowmax_kernel): per-anchor max-logit keys
    for (int a = tid; a < A; a += blockDim.x) keys[a] = keys_g[(long long)b * A + a];
    __syncthreads();
    int take_eq;
    uint32_t T = radix_select_kth(keys, A, kdet, hist, misc, &take_eq);
    collect_topk(keys, A, T, take_eq, cand, &misc[2]);
    // order candidates by (score desc, anchor asc) so that the flat index matches topk's sorted output
    for (int i = tid; i < 512; i += blockDim.x)
        sortbuf[i] = i < kdet ? (((unsigned long long)keys[cand[i]] << 32) | (uint32_t)(0x7fffffff - cand[i])) : 0ull;
    __syncthreads();
    for (int size = 2; size <= 512; size <<= 1) {
        for (int strd = size >> 1; strd > 0; strd >>= 1) {
            for (int i = tid; i < 256; i += blockDim.x) {
                const int lo = 2 * i - (i & (strd - 1));
                const int hi = lo + strd;
                const bool desc = ((lo & size) == 0);
                const unsigned long long x0 = sortbuf[lo], x1 = sortbuf[hi];
                if ((x0 < x1) == desc) { sortbuf[lo] = x1; sortbuf[hi] = x0; }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < kdet; i += blockDim.x) cand[i] = 0x7fffffff - (int)(sortbuf[i] & 0xffffffffull);
    __syncthreads();

    // ---- stage 2: gather kdet x nc logits, select top-kdet
    const int n2 = kdet * nc;
    for (int i = tid; i < n2; i += blockDim.x) {
        const int ci = i / nc, c = i - ci * nc;
        int lvl, r;
        anchor_of(L, cand[ci], lvl, r);
        keys[i] = f2key(L.cls[lvl][((long long)b * L.h[lvl] * L.w[lvl] + r) * nc + c]);
    }
    __syncthreads();
    T = radix_select_kth(keys, n2, kdet, hist, misc, &take_eq);
    collect_topk(keys, n2, T, take_eq, sel, &misc[2]);
    for (int i = tid; i < 512; i += blockDim.x)
        sortbuf[i] = i < kdet ? (((unsigned long long)keys[sel[i]] << 32) | (uint32_t)(0x7fffffff - sel[i])) : 0ull;
    __syncthreads();
    for (int size = 2; size <= 512; size <<= 1) {
        for (int strd = size >> 1; strd > 0; strd >>= 1) {
            for (int i = tid; i < 256; i += blockDim.x) {
                const int lo = 2 * i - (i & (strd - 1));
                const int hi = lo + strd;
                const bool desc = ((lo & size) == 0);
                const unsigned long long x0 = sortbuf[lo], x1 = sortbuf[hi];
                if ((x0 < x1) == desc) { sortbuf[lo] = x1; sortbuf[hi] = x0; }
            }
            __syncthreads();
        }
    }
    // ---- emit
    for (int i = tid; i < kdet; i += blockDim.x) {
        const unsigned long long e = sortbuf[i];
        const int flat = 0x7fffffff - (int)(e & 0xffffffffull);
        const float logit = key2f((uint32_t)(e >> 32));
        const int ci = flat / nc, c = flat - ci * nc;
        const int a = cand[ci];
        int lvl, r;
        anchor_of(L, a, lvl, r);
        const int ay = r / L.w[lvl], ax = r - ay * L.w[lvl];
        const float4 d = *reinterpret_cast<const float4*>(L.box[lvl] + ((long long)b * L.h[lvl] * L.w[lvl] + r) * 4);
        const float s = L.stride[lvl];
        const float px = (float)ax + 0.5f, py = (float)ay + 0.5f;
        float* o = out + ((long long)b * kdet + i) * 6;
        o[0] = (px - d.x) * s;
        o[1] = (py - d.y) * s;
        o[2] = (px + d.z) * s;
        o[3] = (py + d.w) * s;
        o[4] = 1.f / (1.f + expf(-logit));
        o[5] = (float)c;
        if (out_anchor) out_anchor[(long long)b * kdet + i] = a;
    }
}

// Dense decode: y[b, 0:4, a] = box (xyxy if xyxy else xywh) * stride, y[b, 4+c, a] = sigmoid(logit)   (head.py:173-194)
// DFL (block.py:63-85): distance of one side = sum_i i * softmax_i(bins), bins = reg_max logits of that side.
__device__ __forceinline__ float dfl_side(const float* __restrict__ p, int reg_max) {
    if (reg_max == 1) return p[0];
    float m = p[0];
    for (int i = 1; i < reg_max; ++i) m = fmaxf(m, p[i]);
    float s = 0.f, e = 0.f;
    for (int i = 0; i < reg_max; ++i) {
        const float w = expf(p[i] - m);
        s += w;
        e += w * (float)i;
    }
    return e / s;
}

__global__ void __launch_bounds__(256) detect_dense_kernel(const DetectLevels L, int nc, int reg_max, int xyxy, int B,
                                                           float* __restrict__ y) {
    pdl_prologue();
    const int A = L.off[L.nl];
    const int no = 4 + nc;
    const long long total = (long long)B * no * A;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int a = (int)(idx % A);
    const int ch = (int)((idx / A) % no);
    const int b = (int)(idx / ((long long)A * no));
    int lvl, r;
    anchor_of(L, a, lvl, r);
    const long long row = (long long)b * L.h[lvl] * L.w[lvl] + r;
    float v;
    if (ch >= 4) {
        v = 1.f / (1.f + expf(-L.cls[lvl][row * nc + (ch - 4)]));
    } else {
        // channel layout of the box tower: side * reg_max + bin, sides = (l, t, r, b)   head.py:186-194
        const float* d = L.box[lvl] + row * 4 * reg_max;
        const int axis = ch & 1;                       // 0: x (sides l,r), 1: y (sides t,b)
        const float lo = dfl_side(d + axis * reg_max, reg_max), hi = dfl_side(d + (axis + 2) * reg_max, reg_max);
        const int ay = r / L.w[lvl], ax = r - ay * L.w[lvl];
        const float pc = (float)(axis ? ay : ax) + 0.5f, s = L.stride[lvl];
        const float c1 = pc - lo, c2 = pc + hi;
        if (xyxy) v = (ch < 2 ? c1 : c2) * s;
        else v = (ch < 2 ? (c1 + c2) * 0.5f : (c2 - c1)) * s;
    }
    y[idx] = v;
}

}  // namespace ym

using namespace ym;

static int fill_levels(DetectLevels& L, int nl, const void* const* box, const void* const* cls, const int* hs, const int* ws,
                       const float* strides) {
    if (nl < 1 || nl > 4) { ym_set_error("detect: 1..4 levels supported (got %d)", nl); return YM_ERR_ARG; }
    memset(&L, 0, sizeof(L));
    L.nl = nl;
    int off = 0;
    for (int i = 0; i < nl; ++i) {
        if (!box[i] || !cls[i]) { ym_set_error("detect: null level pointer"); return YM_ERR_ARG; }
        L.box[i] = (const float*)box[i]; L.cls[i] = (const float*)cls[i];
        L.h[i] = hs[i]; L.w[i] = ws[i]; L.stride[i] = strides[i];
        L.off[i] = off; off += hs[i] * ws[i];
    }
    L.off[nl] = off;
    return YM_OK;
}

// out: [B, k, 6] fp32 (x1,y1,x2,y2,score,cls) with k = min(max_det, A); out_anchor (nullable): [B, k] int32
extern "C" int ym_detect_topk(int nl, const void* const* box, const void* const* cls, const int* hs, const int* ws,
                              const float* strides, int B, int nc, int max_det, float* out, int* out_anchor, void* scratch,
                              void* stream) {
    YM_CHECK_ARG(out && scratch, "ym_detect_topk: null output / scratch (B*A uint32)");
    DetectLevels L;
    int rc = fill_levels(L, nl, box, cls, hs, ws, strides);
    if (rc) return rc;
    const int A = L.off[nl];
    const int k = max_det < A ? max_det : A;
    YM_CHECK_ARG(k >= 1 && k <= 512, "ym_detect_topk: max_det must be in 1..512 (got %d)", max_det);
    YM_CHECK_ARG(nc >= 1, "ym_detect_topk: nc");
    if (B == 0) return YM_OK;
    const size_t nkeys = (size_t)(A > k * nc ? A : k * nc);
    const size_t smem = nkeys * 4 + (size_t)k * 8 + 8 + 512 * 8;
    YM_CHECK_ARG(smem <= 227 * 1024, "ym_detect_topk: %zu bytes of shared memory needed (A=%d) exceeds 227 KB", smem, A);
    cudaError_t e = cudaFuncSetAttribute(detect_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ym_set_error("ym_detect_topk: smem attr: %s", cudaGetErrorString(e)); return YM_ERR_CUDA; }
    const long long rows = (long long)B * A;
    bool aligned = nc % 4 == 0 && nc / 4 <= RM_MAXQ;
    for (int i = 0; i < nl; ++i) aligned = aligned && (((uintptr_t)cls[i]) & 15) == 0;
    if (aligned) {
        RowmaxGroups G;
        int acc = 0;
        for (int i = 0; i < nl; ++i) { G.first[i] = acc; acc += B * ((hs[i] * ws[i] + RM_ROWS - 1) / RM_ROWS); }
        for (int i = nl; i < 5; ++i) G.first[i] = acc;
        launch_pdl(detect_rowmax8_kernel, (acc + 7) / 8, 256, 0, (cudaStream_t)stream, L, G, nc, B, (uint32_t*)scratch);
    } else {
        launch_pdl(detect_rowmax_kernel, (int)((rows * 32 + 255) / 256), 256, 0, (cudaStream_t)stream, L, nc, B, (uint32_t*)scratch);
    }
    YM_CHECK_LAUNCH("detect_rowmax");
    launch_pdl(detect_topk_kernel, B, 1024, smem, (cudaStream_t)stream, L, nc, k, out, out_anchor, (const uint32_t*)scratch);
    YM_CHECK_LAUNCH("detect_topk");
    return YM_OK;
}

extern "C" int ym_detect_dense(int nl, const void* const* box, const void* const* cls, const int* hs, const int* ws,
                               const float* strides, int B, int nc, int reg_max, int xyxy, float* y, void* stream) {
    YM_CHECK_ARG(y, "ym_detect_dense: null output");
    YM_CHECK_ARG(reg_max >= 1 && reg_max <= 64, "ym_detect_dense: reg_max must be in 1..64 (got %d)", reg_max);
    DetectLevels L;
    int rc = fill_levels(L, nl, box, cls, hs, ws, strides);
    if (rc) return rc;
    if (B == 0) return YM_OK;
    const long long total = (long long)B * (4 + nc) * L.off[nl];
    launch_pdl(detect_dense_kernel, (int)((total + 255) / 256), 256, 0, (cudaStream_t)stream, L, nc, reg_max, xyxy, B, y);
    YM_CHECK_LAUNCH("detect_dense");
    return YM_OK;
}
