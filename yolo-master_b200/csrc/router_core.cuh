// Finish of EfficientSpatialRouter for ONE image by ONE warp (routers.py:283-304, :185-265): spatial mean of the hidden map (from the
// per-tile partial sums of router_fused_kernel) -> logits -> softmax -> top-k (lowest index wins ties) -> renormalised weights.
// Shared by router_finish_kernel (router.cu) and the routed statistics pass of the expert FFN (tc_moe.cu), which runs it in its
// prologue instead of a separate launch: the same instruction sequence, the same bits.
#pragma once
#include "ym_common.cuh"

namespace ym {

struct RouterFin {
    const float* partial;      // [B][nblk][Cr]
    int nblk, Cr, npix;
    const float* w2;           // [E][Cr]
    const float* scale2;       // [E]
    const float* shift2;       // [E]
    int E, topk;
    int* idx_out;              // [B][topk]
    float* w_out;              // [B][topk]
    float* probs_out;          // [B][E] or null
};

// hm, pr: 64 floats of shared memory each.  ids / vals are valid in lane 0 only.  `write`: store idx / w / probs of image b.
__device__ __forceinline__ void router_finish_warp(const RouterFin& r, int b, int lane, float* hm, float* pr, int (&ids)[8], float (&vals)[8],
                                                   bool write) {
    for (int c = lane; c < r.Cr; c += 32) {
        float s = 0.f;
        for (int i = 0; i < r.nblk; ++i) s += r.partial[((long long)b * r.nblk + i) * r.Cr + c];
        hm[c] = s / (float)r.npix;
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int e = lane; e < r.E; e += 32) {
        float acc = 0.f;
        for (int c = 0; c < r.Cr; ++c) acc = fmaf(r.w2[e * r.Cr + c], hm[c], acc);
        const float lg = fmaf(acc, r.scale2[e], r.shift2[e]);
        pr[e] = lg;
        mx = fmaxf(mx, lg);
    }
    mx = warp_max(mx);
    __syncwarp();
    float sum = 0.f;
    for (int e = lane; e < r.E; e += 32) {
        const float v = expf(pr[e] - mx);
        pr[e] = v;
        sum += v;
    }
    sum = warp_sum(sum);
    __syncwarp();
    for (int e = lane; e < r.E; e += 32) {
        pr[e] = pr[e] / sum;
        if (write && r.probs_out) r.probs_out[(long long)b * r.E + e] = pr[e];
    }
    __syncwarp();
    if (lane == 0) {
        unsigned long long taken = 0ull;
        float tot = 0.f;
        for (int j = 0; j < r.topk; ++j) {
            int best = -1;
            float bv = -INFINITY;
            for (int e = 0; e < r.E; ++e) {
                if ((taken >> e) & 1ull) continue;
                if (pr[e] > bv) { bv = pr[e]; best = e; }
            }
            taken |= 1ull << best;
            vals[j] = bv;
            ids[j] = best;
            tot += bv;
        }
        tot = fmaxf(tot, 1e-6f);
        for (int j = 0; j < r.topk; ++j) {
            vals[j] = vals[j] / tot;
            if (write) {
                r.idx_out[b * r.topk + j] = ids[j];
                r.w_out[b * r.topk + j] = vals[j];
            }
        }
    }
}

}  // namespace ym
