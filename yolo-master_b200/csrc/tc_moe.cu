// Routed expert FFN of OptimizedMOEImproved on tcgen05 with the hidden activation kept on chip (sm_100a).
//
// Reference: SimpleExpert = 1x1 conv -> GroupNorm -> SiLU -> 1x1 conv -> GroupNorm (moe/experts.py:73-88), applied to image b by its
// top-k experts and combined (moe/modules.py:1128-1157).  Routing is per IMAGE, so a routed problem p = (image b, rank j) is a plain
// GEMM chain over that image's HW token rows with expert e = route_idx[p]'s weights - no gather, no permutation.
//
// What made the mma.sync version (gemm_conv.cu) slow is traffic, not math: GEMM1 wrote the hidden h [P][HW][HID] (105 MB per P3 block at
// bs32) and GEMM2 read it back, because GroupNorm-1 needs the statistics of the WHOLE image before any element can be normalised.
// Here GEMM1 is cheap enough (K = C = 64 / 128) to run twice:
//   pass 1  ym_moe_ffn_stats  : h = x W1[e]^T per 128-row tile (tcgen05.mma into tensor memory), epilogue = GroupNorm-1 partial sums
//                               of the fp16-rounded h; nothing but 2 * HID/8 floats per (problem, strip) is written
//   (ym_gn_finalize_tiles     : partial sums -> per-(problem, channel) scale / shift, unchanged arithmetic)
//   pass 2  ym_moe_ffn_fused  : h again (bit-identical: same MMA, same operands), epilogue-1 = round to fp16, GroupNorm-1 affine, SiLU,
//                               pack to fp16 and write it back into TENSOR MEMORY as the A operand of a TS-mode tcgen05.mma
//                               o = a W2[e]^T; epilogue-2 = o -> fp16 -> global + GroupNorm-2 partial sums.
// h never exists in global memory: per P3 block the chain moves x (read by both passes, L2-resident between the k ranks) and o.
//
// One CTA = one strip of consecutive 128-row tiles of ONE problem (weights fetched once per CTA by TMA), 288 threads:
//   warp 0 lane 0 : every TMA (x tiles double-buffered, W1 / W2 once) and every tcgen05.mma, static order
//                   GEMM1(0); per tile i: [P2: wait a(i) -> GEMM2(i)]; GEMM1(i+1); refill x(i+2)
//   warps 1-8     : epilogue, thread = token row (tensor-memory lane) x one HALF of the columns: warps w and w+4 share a lane quarter and
//                   split the columns (four epilogue warps per CTA left two warps per SM sub-partition: 14 % warp occupancy, latency-bound
//                   at 72 / 86 us per P3 launch in the second capture); hand-offs through mbarriers only
// Partial statistics are accumulated per thread over the strip and reduced once, in a fixed order (bit-reproducible).
#include <cuda.h>

#include "tc_common.cuh"
#include "router_core.cuh"

namespace ym {

constexpr int MF_BM = 128, MF_THREADS = 288;   // warp 0: TMA + MMA issuer; warps 1-8: epilogue (two per tensor-memory lane quarter)

__device__ __forceinline__ void mf_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mf_tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void mf_tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// GroupNorm partial sums that the CONSUMER kernel finalises itself (one launch less per GroupNorm): [problems][tiles][Cn / 8][2] as
// written by the statistics epilogues, with the routed expert's gamma / beta tables [E][Cn].  stats == nullptr: the caller passes
// finished scale / shift arrays instead (ym_gn_finalize_tiles).
struct GnRaw {
    const float* stats; int tiles, groups; float count, eps;
    const float* gamma; const float* beta;
    const float* route_w;      // optional per-problem factor folded into the affine (the routing weight before the combine)
};

// One warp: scale / shift of channel group `grp` of problem `pr` - the arithmetic of gn_finalize_kernel (gemm_conv.cu) operation for
// operation, so both ways of finalising give the same bits.
__device__ __forceinline__ void gn_group_affine(const GnRaw& g, int Cn, int pr, int e, int grp, int lane, float* scale_out, float* shift_out) {
    const int cpg = Cn / g.groups, tpg = cpg / 8, nt8 = Cn / 8;
    float s = 0.f, q = 0.f;
    for (int i = lane; i < g.tiles * tpg; i += 32) {
        const int mt = i / tpg, t = grp * tpg + i % tpg;
        const float* src = g.stats + (((long long)pr * g.tiles + mt) * nt8 + t) * 2;
        s += src[0];
        q += src[1];
    }
    s = ym::warp_sum(s);
    q = ym::warp_sum(q);
    const float mean = s / g.count;
    const float var = fmaxf(q / g.count - mean * mean, 0.f);
    const float rstd = rsqrtf(var + g.eps);
    const float rw = g.route_w ? g.route_w[pr] : 1.f;
    for (int j = lane; j < cpg; j += 32) {
        const int c = grp * cpg + j;
        const float gm = g.gamma[e * Cn + c], bt = g.beta[e * Cn + c];
        scale_out[c] = rw * rstd * gm;
        shift_out[c] = rw * (bt - mean * rstd * gm);
    }
}

struct MoeFfnParams {
    const int* route_idx;      // [P] expert of problem p (negative = dropped route: the CTA exits)
    int a_div;                 // image of problem p = p / a_div (top_k)
    int HW, mtiles, tiles_per_strip;
    const float* a_scale;      // pass 2: GroupNorm-1 affine per (problem, hidden channel) [P][HID]
    const float* a_shift;
    __half* out;               // pass 2: o [P][HW][C]
    float* stats;              // partial sums [P][strips][NS/2][2], NS/2 = (pass 1 ? HID : C) / 8 eight-channel slices
    GnRaw gn1;                 // pass 2: GroupNorm-1 finalised here when gn1.stats != nullptr (a_scale / a_shift unused)
    RouterFin rf;              // pass 1: the router's finish runs in the prologue when rf.partial != nullptr (route_idx is then an OUTPUT)
};

struct MfBars {
    uint64_t w_full, x_full[2], d1_full, a_full, d2_full;
    uint32_t tmem_slot;
};

// STAGE 1: statistics of h only.  STAGE 2: the fused chain.
template <int C, int HID, int STAGE>
// 9 warps are allocated as 10 (pairs): two CTAs per SM leave 65536 / (20 * 32) = 102 -> 96 registers per thread.  (__maxnreg__(112) fitted the
// code without spills but silently halved the occupancy: 14 % warps active in profiles/r02_moe_ncu.txt, second capture.)
__global__ void __launch_bounds__(MF_THREADS, (HID <= 128) ? 2 : 1)
moe_ffn_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
               const __grid_constant__ CUtensorMap map_w2, const MoeFfnParams p) {
    constexpr int KC1 = C / 64;                        // 64-wide k chunks of GEMM1 (K = C)
    constexpr int KC2 = HID / 64;                      // 64-wide k chunks of GEMM2 (K = HID)
    constexpr int X_BYTES = MF_BM * C * 2, W1_BYTES = HID * C * 2, W2_BYTES = C * HID * 2;
    constexpr uint32_t D1_COLS = HID, A_COLS = HID / 2, D2_COLS = C;
    constexpr uint32_t TMEM_COLS = (STAGE == 1) ? (HID <= 128 ? 128 : 256) : (HID <= 128 ? 256 : 512);
    constexpr int NS = (STAGE == 1 ? HID : C) / 8 * 2;   // floats of partial statistics per thread / per strip
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
    unsigned char* sX = smem;                          // [2][KC1][128 rows x 128 B]
    unsigned char* sW1 = sX + 2 * X_BYTES;             // [KC1][HID rows x 128 B]
    unsigned char* sW2 = sW1 + W1_BYTES;               // [KC2][C rows x 128 B]           (pass 2)
    float* sAff = reinterpret_cast<float*>(sW2 + (STAGE == 2 ? W2_BYTES : 0));   // [2][HID] GroupNorm-1 scale | shift (pass 2)
    unsigned char* sStage = reinterpret_cast<unsigned char*>(sAff + 2 * HID);     // [8 warps][32 rows][64 B] output staging (pass 2)
    __shared__ MfBars bars;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int prob = blockIdx.y, strip = blockIdx.x;
    const int t0 = strip * p.tiles_per_strip;
    const int nt = min(p.tiles_per_strip, p.mtiles - t0);

    if (tid == 0) {
        tc::mbar_init(&bars.w_full, 1);
        tc::mbar_init(&bars.x_full[0], 1);
        tc::mbar_init(&bars.x_full[1], 1);
        tc::mbar_init(&bars.d1_full, 1);
        tc::mbar_init(&bars.a_full, 8);                // one arrive per epilogue warp
        tc::mbar_init(&bars.d2_full, 1);
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&bars.tmem_slot, TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = bars.tmem_slot;
    pdl_prologue();                                    // set-up done: stage the next kernel, then wait for the producers of x / routes / affine
    int e;
    if (STAGE == 1 && p.rf.partial != nullptr) {
        // the router's finish (spatial mean -> logits -> softmax -> top-k) for this CTA's image, by warp 0 of EVERY CTA of the image: a few
        // hundred flops instead of a launch; strip 0 of route 0 publishes idx / w / probs for the kernels that follow
        __shared__ float rf_hm[64], rf_pr[64];
        __shared__ int rf_e;
        if (warp == 0) {
            int ids[8];
            float vals[8];
            const int img_r = prob / p.a_div, jr = prob - img_r * p.a_div;
            router_finish_warp(p.rf, img_r, lane, rf_hm, rf_pr, ids, vals, strip == 0 && jr == 0);
            if (lane == 0) rf_e = ids[jr];
        }
        __syncthreads();
        e = rf_e;
    } else {
        e = p.route_idx[prob];
    }
    const bool active = e >= 0 && nt > 0;              // CTA-uniform
    if (active) {
        const int img = prob / p.a_div;
        if (STAGE == 2) {
            if (p.gn1.stats != nullptr) {
                for (int grp = warp; grp < p.gn1.groups; grp += MF_THREADS / 32) gn_group_affine(p.gn1, HID, prob, e, grp, lane, sAff, sAff + HID);
            } else {
                for (int i = tid; i < HID; i += MF_THREADS) {
                    sAff[i] = p.a_scale[(long long)prob * HID + i];
                    sAff[HID + i] = p.a_shift[(long long)prob * HID + i];
                }
            }
        }
        __syncthreads();
        const uint32_t t_d1 = tmem_base, t_a = tmem_base + D1_COLS, t_d2 = tmem_base + D1_COLS + A_COLS;

        if (warp == 0) {
            if (lane == 0) {
                // ============================================ TMA + MMA issuer ============================================
                auto load_x = [&](int i) {
                    unsigned char* dst = sX + (i & 1) * X_BYTES;
                    mf_expect_tx(&bars.x_full[i & 1], (uint32_t)X_BYTES);
#pragma unroll
                    for (int kc = 0; kc < KC1; ++kc)
                        mf_tma_load_3d(dst + kc * (MF_BM * 128), &map_x, kc * 64, (t0 + i) * MF_BM, img, &bars.x_full[i & 1]);
                };
                mf_expect_tx(&bars.w_full, (uint32_t)(W1_BYTES + (STAGE == 2 ? W2_BYTES : 0)));
#pragma unroll
                for (int kc = 0; kc < KC1; ++kc) mf_tma_load_2d(sW1 + kc * (HID * 128), &map_w1, kc * 64, e * HID, &bars.w_full);
                if (STAGE == 2) {
#pragma unroll
                    for (int kc = 0; kc < KC2; ++kc) mf_tma_load_2d(sW2 + kc * (C * 128), &map_w2, kc * 64, e * C, &bars.w_full);
                }
                load_x(0);
                if (nt > 1) load_x(1);
                const uint32_t idesc1 = tc::make_idesc_f16(MF_BM, HID, 0);
                const uint32_t idesc2 = tc::make_idesc_f16(MF_BM, C, 0);
                auto gemm1 = [&](int i) {
                    const uint32_t xa = smem_u32(sX + (i & 1) * X_BYTES), wa = smem_u32(sW1);
#pragma unroll
                    for (int kc = 0; kc < KC1; ++kc) {
                        const uint64_t ad = tc::make_desc_sw128(xa + kc * (MF_BM * 128)), bd = tc::make_desc_sw128(wa + kc * (HID * 128));
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) tc::mma_f16_ss(t_d1, ad + 2 * ks, bd + 2 * ks, idesc1, (kc | ks) ? 1u : 0u);
                    }
                    tc::mma_commit(&bars.d1_full);
                };
                tc::mbar_wait(&bars.w_full, 0);
                tc::mbar_wait(&bars.x_full[0], 0);
                gemm1(0);
                for (int i = 0; i < nt; ++i) {
                    // the epilogue has read D1(i) (and, pass 2, written the A operand): D1 and x buffer i&1 are free
                    tc::mbar_wait(&bars.a_full, i & 1);
                    tc::fence_after_sync();
                    if (STAGE == 2) {
                        const uint32_t wa = smem_u32(sW2);
#pragma unroll
                        for (int kc = 0; kc < KC2; ++kc) {
                            const uint64_t bd = tc::make_desc_sw128(wa + kc * (C * 128));
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)      // 16 hidden channels = 8 tensor-memory columns of the packed A operand
                                tc::mma_f16_ts(t_d2, t_a + 8 * (kc * 4 + ks), bd + 2 * ks, idesc2, (kc | ks) ? 1u : 0u);
                        }
                        tc::mma_commit(&bars.d2_full);
                    }
                    if (i + 1 < nt) {
                        tc::mbar_wait(&bars.x_full[(i + 1) & 1], ((i + 1) >> 1) & 1);
                        gemm1(i + 1);
                        if (i + 2 < nt) load_x(i + 2);       // buffer i&1: GEMM1(i) retired before a_full(i)
                    }
                }
            }
        } else {
            // ====================================== epilogue warps: thread = token row x one half of the columns ======================
            const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
            const int half = (warp - 1) >> 2;                               // 0: lower half of the columns, 1: upper half
            const int row_in_tile = (warp & 3) * 32 + lane;
            constexpr int NSH = NS / 2;                                     // statistics slots of this thread (its half of the columns)
            constexpr int H1 = HID / 2, H2 = C / 2;                         // columns per half: GEMM1 / GEMM2 accumulator
            float part[NSH];
#pragma unroll
            for (int i = 0; i < NSH; ++i) part[i] = 0.f;
            for (int i = 0; i < nt; ++i) {
                const int row = (t0 + i) * MF_BM + row_in_tile;
                const bool live = row < p.HW;
                tc::mbar_wait(&bars.d1_full, i & 1);
                tc::fence_after_sync();
#pragma unroll
                for (int c0 = 0; c0 < H1; c0 += 32) {
                    const int cb = half * H1 + c0;                          // first hidden channel of this 32-column chunk
                    uint32_t v[32];
                    tc::tmem_ld32(t_d1 + lane_sel + cb, v);
                    tc::tmem_ld_wait();
                    if (STAGE == 1) {
                        if (live) {
#pragma unroll
                            for (int q = 0; q < 16; ++q) {
                                const float2 r = __half22float2(__floats2half2_rn(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])));
                                const int sl = (c0 + 2 * q) >> 3;
                                part[2 * sl] += r.x + r.y;
                                part[2 * sl + 1] += r.x * r.x + r.y * r.y;
                            }
                        }
                    } else {
                        uint32_t pk[16];
#pragma unroll
                        for (int q4 = 0; q4 < 8; ++q4) {          // four channels per step: one 16-byte broadcast read each of scale / shift
                            const float4 sc4 = *reinterpret_cast<const float4*>(sAff + cb + 4 * q4);
                            const float4 sh4 = *reinterpret_cast<const float4*>(sAff + HID + cb + 4 * q4);
                            const float2 r0 = __half22float2(__floats2half2_rn(__uint_as_float(v[4 * q4]), __uint_as_float(v[4 * q4 + 1])));
                            const float2 r1 = __half22float2(__floats2half2_rn(__uint_as_float(v[4 * q4 + 2]), __uint_as_float(v[4 * q4 + 3])));
                            pk[2 * q4] = pack_half2(silu_f(fmaf(r0.x, sc4.x, sh4.x)), silu_f(fmaf(r0.y, sc4.y, sh4.y)));
                            pk[2 * q4 + 1] = pack_half2(silu_f(fmaf(r1.x, sc4.z, sh4.z)), silu_f(fmaf(r1.y, sc4.w, sh4.w)));
                        }
                        tc::tmem_st16(t_a + lane_sel + cb / 2, pk);
                    }
                }
                if (STAGE == 2) tc::tmem_st_wait();
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.a_full);
                if (STAGE == 2) {
                    tc::mbar_wait(&bars.d2_full, i & 1);
                    tc::fence_after_sync();
                    // o rows leave through a warp-private staging tile (thread = token row of the accumulator, but row-strided 16-byte
                    // stores choke the LSU: lg_throttle 2.25 / issue in the first capture); staged, every store instruction writes full sectors
                    unsigned char* stg = sStage + (warp - 1) * 2048;        // [32 rows][64 B]: one 32-column chunk at a time
                    const int tile_row0 = (t0 + i) * MF_BM + (warp & 3) * 32;
#pragma unroll
                    for (int c0 = 0; c0 < H2; c0 += 32) {
                        uint32_t v[32];
                        tc::tmem_ld32(t_d2 + lane_sel + half * H2 + c0, v);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8) {
                            Half8 hv;
                            float sm = 0.f, q2 = 0.f;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                hv.v[q] = __floats2half2_rn(__uint_as_float(v[c8 * 8 + 2 * q]), __uint_as_float(v[c8 * 8 + 2 * q + 1]));
                                const float2 r = __half22float2(hv.v[q]);
                                sm += r.x + r.y;
                                q2 += r.x * r.x + r.y * r.y;
                            }
                            if (live) {
                                const int sl = (c0 >> 3) + c8;
                                part[2 * sl] += sm;
                                part[2 * sl + 1] += q2;
                            }
                            *reinterpret_cast<uint4*>(stg + lane * 64 + ((c8 ^ ((lane >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(&hv);
                        }
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int idx = k * 32 + lane, rr = idx >> 2, ch = idx & 3;
                            if (tile_row0 + rr < p.HW)
                                *reinterpret_cast<uint4*>(p.out + ((long long)prob * p.HW + tile_row0 + rr) * C + half * H2 + c0 + ch * 8) =
                                    *reinterpret_cast<const uint4*>(stg + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
                        }
                        __syncwarp();
                    }
                    tc::fence_before_sync();       // D2 / the A operand are read out before the next tile's MMAs may overwrite them
                }
            }
            // ---- strip statistics: per-thread partials -> shared memory -> fixed-order column sums (bit-reproducible)
            // every MMA that read the x buffers has retired (d1_full / d2_full of the last tile were waited on), so sX is free
            float* red = reinterpret_cast<float*>(sX);
            const int r = tid - 32;                                         // 0..255: [half][row]
#pragma unroll
            for (int i = 0; i < NSH; ++i) red[r * NSH + i] = part[i];
            asm volatile("bar.sync 1, 256;\n" ::: "memory");       // the eight epilogue warps only
            if (r < NS) {
                const int h2 = r / NSH, jj = r - h2 * NSH;                  // statistics slot r belongs to column half h2
                float a = 0.f;
                for (int k = 0; k < 128; ++k) a += red[(h2 * 128 + k) * NSH + jj];
                p.stats[((long long)prob * gridDim.x + strip) * NS + r] = a;
            }
        }
    } else if (warp >= 1) {
        // dropped route / empty trailing strip: the statistics slots must still be defined (the finalize kernel sums every strip)
        const int r = tid - 32;
        if (r < NS) p.stats[((long long)prob * gridDim.x + strip) * NS + r] = 0.f;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---- combine: y = SiLU(x Ws^T + bs) + sum_j (o[b,j] * scale2[b,j] + shift2[b,j]) + x   (moe/modules.py:1144-1157) ------------------
// shared expert = 1x1 conv + folded BN + SiLU on the tensor core (tcgen05, accumulator double-buffered in tensor memory so the MMA of
// tile i+1 runs under the epilogue of tile i); the epilogue thread (= token row) adds the routed experts' GroupNorm-2-normalised outputs
// (routing weight folded into scale2 / shift2 by ym_gn_finalize_tiles) and the residual, one rounding to fp16, one 128 / 256-byte row store.
struct MoeCombineParams {
    const __half* __restrict__ x;   int ldx;        // residual read (the GEMM operand comes through TMA)
    const float* __restrict__ bias;                 // [C] folded BN bias of the shared expert
    const __half* __restrict__ o;                   // [B*topk][HW][C]
    const float* __restrict__ o_scale; const float* __restrict__ o_shift;   // [B*topk][C]
    __half* __restrict__ out;       int ldo;
    GnRaw gn2;                                      // GroupNorm-2 finalised here when gn2.stats != nullptr (o_scale / o_shift unused)
    const int* __restrict__ route_idx;              // with gn2: expert of problem img * topk + j
    int HW, mtiles, tiles_per_strip, topk, add_residual;
};
struct McBars {
    uint64_t w_full, x_full[2], d_full[2], d_free[2];
    uint32_t tmem_slot;
};

template <int C>
__global__ void __launch_bounds__(MF_THREADS, (C <= 64) ? 2 : 1)   // C = 128 holds 131 KB of shared memory: one CTA per SM either way
moe_combine_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const MoeCombineParams p) {
    constexpr int KC = C / 64;
    constexpr int X_BYTES = MF_BM * C * 2, W_BYTES = C * C * 2;
    constexpr uint32_t TMEM_COLS = 2 * C;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
    unsigned char* sX = smem;                          // [2][KC][128 rows x 128 B]
    unsigned char* sW = sX + 2 * X_BYTES;              // [KC][C rows x 128 B]
    float* sAff = reinterpret_cast<float*>(sW + W_BYTES);   // bias [C] | per rank j: scale [C], shift [C]   (topk <= 2)
    float* sStage = sAff + 5 * C;                           // [8 warps][32 rows][32 fp32] staging of SiLU(shared expert)
    __shared__ McBars bars;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int img = blockIdx.y, strip = blockIdx.x;
    const int t0 = strip * p.tiles_per_strip;
    const int nt = min(p.tiles_per_strip, p.mtiles - t0);
    if (tid == 0) {
        tc::mbar_init(&bars.w_full, 1);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bars.x_full[i], 1);
            tc::mbar_init(&bars.d_full[i], 1);
            tc::mbar_init(&bars.d_free[i], 8);
        }
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&bars.tmem_slot, TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = bars.tmem_slot;
    pdl_prologue();
    if (nt > 0) {
        for (int i = tid; i < C; i += MF_THREADS) sAff[i] = p.bias ? p.bias[i] : 0.f;
        if (p.gn2.stats != nullptr) {
            for (int pg = warp; pg < p.topk * p.gn2.groups; pg += MF_THREADS / 32) {
                const int j = pg / p.gn2.groups, grp = pg - j * p.gn2.groups, pr = img * p.topk + j;
                gn_group_affine(p.gn2, C, pr, p.route_idx[pr], grp, lane, sAff + C + (2 * j) * C, sAff + C + (2 * j + 1) * C);
            }
        } else {
            for (int i = tid; i < p.topk * C; i += MF_THREADS) {
                const int j = i / C, c = i - j * C;
                sAff[C + (2 * j) * C + c] = p.o_scale[((long long)img * p.topk + j) * C + c];
                sAff[C + (2 * j + 1) * C + c] = p.o_shift[((long long)img * p.topk + j) * C + c];
            }
        }
        __syncthreads();
        if (warp == 0) {
            if (lane == 0) {
                auto load_x = [&](int i) {
                    unsigned char* dst = sX + (i & 1) * X_BYTES;
                    mf_expect_tx(&bars.x_full[i & 1], (uint32_t)X_BYTES);
#pragma unroll
                    for (int kc = 0; kc < KC; ++kc)
                        mf_tma_load_3d(dst + kc * (MF_BM * 128), &map_x, kc * 64, (t0 + i) * MF_BM, img, &bars.x_full[i & 1]);
                };
                mf_expect_tx(&bars.w_full, (uint32_t)W_BYTES);
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) mf_tma_load_2d(sW + kc * (C * 128), &map_w, kc * 64, 0, &bars.w_full);
                load_x(0);
                if (nt > 1) load_x(1);
                const uint32_t idesc = tc::make_idesc_f16(MF_BM, C, 0);
                tc::mbar_wait(&bars.w_full, 0);
                for (int i = 0; i < nt; ++i) {
                    const int bf = i & 1;
                    tc::mbar_wait(&bars.x_full[bf], (i >> 1) & 1);
                    if (i >= 2) {                            // the epilogue has drained accumulator bf (tile i-2)
                        tc::mbar_wait(&bars.d_free[bf], ((i >> 1) - 1) & 1);
                        tc::fence_after_sync();
                    }
                    const uint32_t xa = smem_u32(sX + bf * X_BYTES), wa = smem_u32(sW);
#pragma unroll
                    for (int kc = 0; kc < KC; ++kc) {
                        const uint64_t ad = tc::make_desc_sw128(xa + kc * (MF_BM * 128)), bd = tc::make_desc_sw128(wa + kc * (C * 128));
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) tc::mma_f16_ss(tmem_base + bf * C, ad + 2 * ks, bd + 2 * ks, idesc, (kc | ks) ? 1u : 0u);
                    }
                    tc::mma_commit(&bars.d_full[bf]);
                    // x buffer bf is re-filled for tile i+2 once this tile's MMA has retired: the epilogue's d_free(i) arrive implies it
                    if (i + 2 < nt) {
                        tc::mbar_wait(&bars.d_full[bf], (i >> 1) & 1);
                        load_x(i + 2);
                    }
                }
            }
        } else {
            const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
            const int half = (warp - 1) >> 2;                               // this warp's half of the channels
            constexpr int H = C / 2;
            for (int i = 0; i < nt; ++i) {
                const int bf = i & 1;
                tc::mbar_wait(&bars.d_full[bf], (i >> 1) & 1);
                tc::fence_after_sync();
                // The accumulator arrives one token ROW per thread; the routed outputs, the residual and y live in global memory as rows of
                // 128 / 256 bytes.  Row-strided 16-byte accesses (32 sectors per instruction) throttled the LSU in the first version
                // (130 us per P3 launch), so SiLU(shared) is staged in a warp-private fp32 tile, 32 channels at a time, and the sum is formed
                // in (row, 8-channel group) ownership: every global load / store instruction then covers whole 32-byte sectors.
                float* stg = sStage + (warp - 1) * (32 * 32);               // [32 rows][32 fp32], 16-byte chunks XOR-swizzled by row
                const int tile_row0 = (t0 + i) * MF_BM + (warp & 3) * 32;
                const int cg = lane & 3;                                     // this lane's 8-channel group in the coalesced phase
#pragma unroll
                for (int c0 = 0; c0 < H; c0 += 32) {
                    const int cb = half * H + c0;                            // first channel of this chunk
                    uint32_t v[32];
                    tc::tmem_ld32(tmem_base + bf * C + lane_sel + cb, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(sAff + cb + c4 * 4);
                        float4 f;
                        f.x = silu_f(__uint_as_float(v[c4 * 4 + 0]) + b4.x);
                        f.y = silu_f(__uint_as_float(v[c4 * 4 + 1]) + b4.y);
                        f.z = silu_f(__uint_as_float(v[c4 * 4 + 2]) + b4.z);
                        f.w = silu_f(__uint_as_float(v[c4 * 4 + 3]) + b4.w);
                        *reinterpret_cast<float4*>(stg + lane * 32 + ((c4 ^ (lane & 7)) << 2)) = f;
                    }
                    __syncwarp();
                    float sc[2][8], sh[2][8];                                // this lane's channels: GroupNorm-2 affine of each routed rank
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (j < p.topk) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                sc[j][q] = sAff[C + (2 * j) * C + cb + cg * 8 + q];
                                sh[j][q] = sAff[C + (2 * j + 1) * C + cb + cg * 8 + q];
                            }
                        }
                    }
                    // all global reads of the four row groups first (12 independent 16-byte loads in flight per lane), then the arithmetic:
                    // in program order every load sat behind the previous group's store and paid its full latency (60 % of the samples)
                    #pragma unroll
                    for (int kb = 0; kb < 4; kb += 2) {          // two row groups at a time: 6 loads in flight, 24 registers
                    uint4 oq[2][2], xq[2];
#pragma unroll
                    for (int k2 = 0; k2 < 2; ++k2) {
                        const int k = kb + k2;
                        const int row = tile_row0 + ((k * 32 + lane) >> 2);
                        const bool ok = row < p.HW;
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            oq[k2][j] = (ok && j < p.topk) ? *reinterpret_cast<const uint4*>(p.o + (((long long)img * p.topk + j) * p.HW + row) * C + cb + cg * 8)
                                                          : make_uint4(0u, 0u, 0u, 0u);
                        xq[k2] = (ok && p.add_residual) ? *reinterpret_cast<const uint4*>(p.x + ((long long)img * p.HW + row) * p.ldx + cb + cg * 8)
                                                       : make_uint4(0u, 0u, 0u, 0u);
                    }
#pragma unroll
                    for (int k2 = 0; k2 < 2; ++k2) {
                        const int k = kb + k2;
                        const int rr = (k * 32 + lane) >> 2;
                        const int row = tile_row0 + rr;
                        if (row < p.HW) {
                            const float4 a0 = *reinterpret_cast<const float4*>(stg + rr * 32 + (((2 * cg) ^ (rr & 7)) << 2));
                            const float4 a1 = *reinterpret_cast<const float4*>(stg + rr * 32 + (((2 * cg + 1) ^ (rr & 7)) << 2));
                            float acc[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                if (j < p.topk) {
                                    const Half8 ov = *reinterpret_cast<const Half8*>(&oq[k2][j]);
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        const float2 f = __half22float2(ov.v[q]);
                                        acc[2 * q] += fmaf(f.x, sc[j][2 * q], sh[j][2 * q]);
                                        acc[2 * q + 1] += fmaf(f.y, sc[j][2 * q + 1], sh[j][2 * q + 1]);
                                    }
                                }
                            }
                            Half8 hv;
                            if (p.add_residual) {
                                const Half8 rv = *reinterpret_cast<const Half8*>(&xq[k2]);
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float2 rf = __half22float2(rv.v[q]);
                                    hv.v[q] = __floats2half2_rn(acc[2 * q] + rf.x, acc[2 * q + 1] + rf.y);
                                }
                            } else {
#pragma unroll
                                for (int q = 0; q < 4; ++q) hv.v[q] = __floats2half2_rn(acc[2 * q], acc[2 * q + 1]);
                            }
                            *reinterpret_cast<uint4*>(p.out + ((long long)img * p.HW + row) * p.ldo + cb + cg * 8) = *reinterpret_cast<const uint4*>(&hv);
                        }
                    }
                    }
                    __syncwarp();
                }
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.d_free[bf]);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

typedef CUresult (*MfEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static MfEncodeFn mf_encode() {
    static MfEncodeFn fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<MfEncodeFn>(q);
    }
    return fn;
}
static bool mf_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    cuuint32_t est[3] = {1, 1, 1};
    return mf_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, est,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int C, int HID, int STAGE>
static int mf_launch(const CUtensorMap& mx, const CUtensorMap& mw1, const CUtensorMap& mw2, const MoeFfnParams& p, int strips, int P,
                     cudaStream_t st) {
    size_t smem = (size_t)2 * MF_BM * C * 2 + (size_t)HID * C * 2 + 1024;
    if (STAGE == 2) smem += (size_t)C * HID * 2 + 2 * HID * sizeof(float) + (size_t)8 * 2048;
    auto kern = moe_ffn_kernel<C, HID, STAGE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ym_set_error("ym_moe_ffn: smem attr %zu: %s", smem, cudaGetErrorString(e)); return YM_ERR_CUDA; }
    e = launch_pdl(kern, dim3(strips, P), dim3(MF_THREADS), smem, st, mx, mw1, mw2, p);
    if (e != cudaSuccess) { ym_set_error("ym_moe_ffn: launch: %s", cudaGetErrorString(e)); return YM_ERR_CUDA; }
    return YM_OK;
}

}  // namespace ym

using namespace ym;

extern "C" int ym_moe_ffn_supported(int C, int HID, int ldx) {
    return mf_encode() != nullptr && ((C == 64 && HID == 128) || (C == 128 && HID == 256)) && ldx % 8 == 0;
}

// Row tiles per CTA: the smallest strip length whose grid fits ONE wave of resident CTAs (2 per SM).  A CTA's fixed cost (tensor-memory
// allocation, barrier set-up, the expert's weights by TMA, pipeline fill: ~5 us) is then paid once per SM slot instead of once per tile -
// at P4 / P5 (13 / 4 row tiles per problem) one-tile CTAs in three waves spent most of their life in that prologue.
static int mf_tiles_per_strip(long long units, int mtiles) {
    static int slots = 0;
    if (!slots) {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        slots = 2 * (sms > 0 ? sms : 148);
    }
    int best = 1;
    long long best_cost = -1;
    for (int tps = 1; tps <= mtiles; ++tps) {          // makespan ~ waves x (tiles per CTA + ~one tile of CTA prologue); ties -> more CTAs
        const long long ctas = units * ((mtiles + tps - 1) / tps);
        const long long cost = ((ctas + slots - 1) / slots) * (tps + 1);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = tps; }
    }
    return best;
}

extern "C" int ym_moe_ffn_strips(int HW, int P) {
    const int mtiles = (HW + MF_BM - 1) / MF_BM;
    const int tps = mf_tiles_per_strip(P, mtiles);
    return (mtiles + tps - 1) / tps;
}

extern "C" long long ym_moe_ffn_stats_floats(int P, int strips, int N) { return (long long)P * strips * (N / 8) * 2; }

// stage 1: stats != null, out == null: GroupNorm-1 partial sums of h = x W1[e]^T           -> stats [P][strips][HID/8][2]
// stage 2: out  != null              : o = SiLU(GN1(h)) W2[e]^T (fp16) and its partial sums -> out [P][HW][C], stats [P][strips][C/8][2]
static int moe_ffn_impl(int stage, const void* x, int ldx, int B, int HW, int C, int HID, int topk, const void* w1, const void* w2, int E,
                        const int* route_idx, const float* a_scale, const float* a_shift, const GnRaw& gn1, const RouterFin& rf, void* out,
                        float* stats, int strips, void* stream) {
    YM_CHECK_ARG(x && w1 && (route_idx || rf.partial) && stats, "ym_moe_ffn: null pointer");
    YM_CHECK_ARG(stage == 1 || (stage == 2 && w2 && ((a_scale && a_shift) || gn1.stats) && out), "ym_moe_ffn: stage %d needs w2 / GroupNorm-1 affine or statistics / out", stage);
    YM_CHECK_ARG(ym_moe_ffn_supported(C, HID, ldx), "ym_moe_ffn: unsupported shape C=%d HID=%d ldx=%d (64/128 or 128/256)", C, HID, ldx);
    YM_CHECK_ARG((((uintptr_t)x | (uintptr_t)w1 | (uintptr_t)w2 | (uintptr_t)out) & 15) == 0, "ym_moe_ffn: 16-byte alignment");
    YM_CHECK_ARG(topk >= 1 && E >= 1 && HW >= 1 && strips >= 1, "ym_moe_ffn: bad sizes");
    if (B == 0) return YM_OK;
    const int P = B * topk, mtiles = (HW + MF_BM - 1) / MF_BM;
    YM_CHECK_ARG(P <= 65535, "ym_moe_ffn: too many routed problems (%d)", P);
    CUtensorMap mx, mw1, mw2;
    {
        cuuint64_t d[3] = {(cuuint64_t)C, (cuuint64_t)HW, (cuuint64_t)B}, s[2] = {(cuuint64_t)ldx * 2, (cuuint64_t)HW * ldx * 2};
        cuuint32_t bx[3] = {64, MF_BM, 1};
        if (!mf_map(&mx, x, 3, d, s, bx)) { ym_set_error("ym_moe_ffn: tensor map (x) failed"); return YM_ERR_CUDA; }
    }
    {
        cuuint64_t d[2] = {(cuuint64_t)C, (cuuint64_t)E * HID}, s[1] = {(cuuint64_t)C * 2};
        cuuint32_t bx[2] = {64, (cuuint32_t)HID};
        if (!mf_map(&mw1, w1, 2, d, s, bx)) { ym_set_error("ym_moe_ffn: tensor map (w1) failed"); return YM_ERR_CUDA; }
    }
    if (stage == 2) {
        cuuint64_t d[2] = {(cuuint64_t)HID, (cuuint64_t)E * C}, s[1] = {(cuuint64_t)HID * 2};
        cuuint32_t bx[2] = {64, (cuuint32_t)C};
        if (!mf_map(&mw2, w2, 2, d, s, bx)) { ym_set_error("ym_moe_ffn: tensor map (w2) failed"); return YM_ERR_CUDA; }
    } else {
        mw2 = mw1;
    }
    MoeFfnParams p;
    p.route_idx = route_idx; p.a_div = topk; p.HW = HW; p.mtiles = mtiles; p.tiles_per_strip = (mtiles + strips - 1) / strips;
    p.a_scale = a_scale; p.a_shift = a_shift; p.out = (__half*)out; p.stats = stats; p.gn1 = gn1; p.rf = rf;
    YM_CHECK_ARG((long long)strips * p.tiles_per_strip >= mtiles, "ym_moe_ffn: strips do not cover the tiles");
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 64) return stage == 1 ? mf_launch<64, 128, 1>(mx, mw1, mw2, p, strips, P, st) : mf_launch<64, 128, 2>(mx, mw1, mw2, p, strips, P, st);
    return stage == 1 ? mf_launch<128, 256, 1>(mx, mw1, mw2, p, strips, P, st) : mf_launch<128, 256, 2>(mx, mw1, mw2, p, strips, P, st);
}

static bool gn_raw_ok(const GnRaw& g, int Cn) {
    return g.stats && g.gamma && g.beta && g.tiles >= 1 && g.groups >= 1 && Cn % g.groups == 0 && (Cn / g.groups) % 8 == 0 && g.count > 0.f;
}

extern "C" int ym_moe_ffn(int stage, const void* x, int ldx, int B, int HW, int C, int HID, int topk, const void* w1, const void* w2, int E,
                          const int* route_idx, const float* a_scale, const float* a_shift, void* out, float* stats, int strips,
                          void* stream) {
    GnRaw none;
    RouterFin norf;
    memset(&none, 0, sizeof(none));
    memset(&norf, 0, sizeof(norf));
    return moe_ffn_impl(stage, x, ldx, B, HW, C, HID, topk, w1, w2, E, route_idx, a_scale, a_shift, none, norf, out, stats, strips, stream);
}

// Pass 1 (GroupNorm-1 statistics) with the router's finish in its prologue: `partial` are the per-tile sums of ym_router_partial
// ([B][nblk][Cr]), (w2, scale2, shift2) the router's second layer; idx_out / w_out (/ probs_out) are WRITTEN here, for this kernel's own
// weight selection and for the kernels that follow.  Same results as ym_router_topk + ym_moe_ffn(1, ...), one launch less.
extern "C" int ym_moe_ffn_routed(const void* x, int ldx, int B, int HW, int C, int HID, int topk, const void* w1, int E, const float* partial,
                                 int nblk, int Cr, int npix, const float* rw2, const float* rscale2, const float* rshift2, int* idx_out,
                                 float* w_out, float* probs_out, float* stats, int strips, void* stream) {
    YM_CHECK_ARG(partial && rw2 && rscale2 && rshift2 && idx_out && w_out, "ym_moe_ffn_routed: null pointer");
    YM_CHECK_ARG(nblk >= 1 && npix >= 1 && Cr >= 1 && Cr <= 64 && E >= 1 && E <= 64 && topk >= 1 && topk <= 8 && topk <= E,
                 "ym_moe_ffn_routed: need Cr <= 64, 1 <= topk <= min(8, E), E <= 64");
    GnRaw none;
    RouterFin rf;
    memset(&none, 0, sizeof(none));
    rf.partial = partial; rf.nblk = nblk; rf.Cr = Cr; rf.npix = npix; rf.w2 = rw2; rf.scale2 = rscale2; rf.shift2 = rshift2;
    rf.E = E; rf.topk = topk; rf.idx_out = idx_out; rf.w_out = w_out; rf.probs_out = probs_out;
    return moe_ffn_impl(1, x, ldx, B, HW, C, HID, topk, w1, nullptr, E, nullptr, nullptr, nullptr, none, rf, nullptr, stats, strips, stream);
}

// Stage 2 with GroupNorm-1 finalised inside the kernel from stage 1's partial sums (no ym_gn_finalize_tiles launch in between).
extern "C" int ym_moe_ffn_gn(const void* x, int ldx, int B, int HW, int C, int HID, int topk, const void* w1, const void* w2, int E,
                             const int* route_idx, const float* gn1_stats, int gn1_groups, float gn1_count, float gn1_eps, const float* gamma1,
                             const float* beta1, void* out, float* stats, int strips, void* stream) {
    GnRaw g;
    memset(&g, 0, sizeof(g));
    g.stats = gn1_stats; g.tiles = strips; g.groups = gn1_groups; g.count = gn1_count; g.eps = gn1_eps; g.gamma = gamma1; g.beta = beta1;
    YM_CHECK_ARG(gn_raw_ok(g, HID), "ym_moe_ffn_gn: bad GroupNorm-1 description (groups %d over %d channels)", gn1_groups, HID);
    RouterFin norf;
    memset(&norf, 0, sizeof(norf));
    return moe_ffn_impl(2, x, ldx, B, HW, C, HID, topk, w1, w2, E, route_idx, nullptr, nullptr, g, norf, out, stats, strips, stream);
}

// y[b] = SiLU(x[b] Ws^T + bs) + sum_j (o[b*topk+j] * o_scale + o_shift) (+ x[b]): the combine of OptimizedMOEImproved on tcgen05.
extern "C" int ym_moe_combine_tc_supported(int C, int ldx, int ldo) { return mf_encode() != nullptr && (C == 64 || C == 128) && ldx % 8 == 0 && ldo % 8 == 0; }

static int moe_combine_tc_impl(const void* x, int ldx, int B, int HW, int C, const void* ws, const float* bias_s, const void* o,
                               const float* o_scale, const float* o_shift, const GnRaw& gn2, const int* route_idx, int topk, void* out, int ldo,
                               int add_residual, void* stream) {
    YM_CHECK_ARG(x && ws && o && ((o_scale && o_shift) || (gn2.stats && route_idx)) && out, "ym_moe_combine_tc: null pointer");
    YM_CHECK_ARG(ym_moe_combine_tc_supported(C, ldx, ldo), "ym_moe_combine_tc: unsupported shape C=%d ldx=%d ldo=%d", C, ldx, ldo);
    YM_CHECK_ARG((((uintptr_t)x | (uintptr_t)ws | (uintptr_t)o | (uintptr_t)out) & 15) == 0, "ym_moe_combine_tc: 16-byte alignment");
    YM_CHECK_ARG(topk >= 1 && topk <= 2 && HW >= 1 && B >= 0 && B <= 65535, "ym_moe_combine_tc: top_k must be 1 or 2 (got %d)", topk);
    if (B == 0) return YM_OK;
    const int mtiles = (HW + MF_BM - 1) / MF_BM;
    const int tps = mf_tiles_per_strip(B, mtiles);
    const int strips = (mtiles + tps - 1) / tps;
    CUtensorMap mx, mw;
    {
        cuuint64_t d[3] = {(cuuint64_t)C, (cuuint64_t)HW, (cuuint64_t)B}, s[2] = {(cuuint64_t)ldx * 2, (cuuint64_t)HW * ldx * 2};
        cuuint32_t bx[3] = {64, MF_BM, 1};
        if (!mf_map(&mx, x, 3, d, s, bx)) { ym_set_error("ym_moe_combine_tc: tensor map (x) failed"); return YM_ERR_CUDA; }
    }
    {
        cuuint64_t d[2] = {(cuuint64_t)C, (cuuint64_t)C}, s[1] = {(cuuint64_t)C * 2};
        cuuint32_t bx[2] = {64, (cuuint32_t)C};
        if (!mf_map(&mw, ws, 2, d, s, bx)) { ym_set_error("ym_moe_combine_tc: tensor map (ws) failed"); return YM_ERR_CUDA; }
    }
    MoeCombineParams p;
    p.x = (const __half*)x; p.ldx = ldx; p.bias = bias_s; p.o = (const __half*)o; p.o_scale = o_scale; p.o_shift = o_shift;
    p.out = (__half*)out; p.ldo = ldo; p.HW = HW; p.mtiles = mtiles; p.tiles_per_strip = tps; p.topk = topk; p.add_residual = add_residual;
    p.gn2 = gn2; p.route_idx = route_idx;
    const size_t smem = (size_t)2 * MF_BM * C * 2 + (size_t)C * C * 2 + (size_t)5 * C * sizeof(float) + 8 * 32 * 32 * sizeof(float) + 1024;
    cudaError_t e;
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 64) {
        e = cudaFuncSetAttribute(moe_combine_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = launch_pdl(moe_combine_tc_kernel<64>, dim3(strips, B), dim3(MF_THREADS), smem, st, mx, mw, p);
    } else {
        e = cudaFuncSetAttribute(moe_combine_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = launch_pdl(moe_combine_tc_kernel<128>, dim3(strips, B), dim3(MF_THREADS), smem, st, mx, mw, p);
    }
    if (e != cudaSuccess) { ym_set_error("ym_moe_combine_tc: %s", cudaGetErrorString(e)); return YM_ERR_CUDA; }
    return YM_OK;
}

extern "C" int ym_moe_combine_tc(const void* x, int ldx, int B, int HW, int C, const void* ws, const float* bias_s, const void* o,
                                 const float* o_scale, const float* o_shift, int topk, void* out, int ldo, int add_residual, void* stream) {
    GnRaw none;
    memset(&none, 0, sizeof(none));
    return moe_combine_tc_impl(x, ldx, B, HW, C, ws, bias_s, o, o_scale, o_shift, none, nullptr, topk, out, ldo, add_residual, stream);
}

// The combine with GroupNorm-2 (times the routing weight) finalised inside the kernel from pass 2's partial sums.
extern "C" int ym_moe_combine_tc_gn(const void* x, int ldx, int B, int HW, int C, const void* ws, const float* bias_s, const void* o,
                                    const float* gn2_stats, int gn2_tiles, int gn2_groups, float gn2_count, float gn2_eps, const float* gamma2,
                                    const float* beta2, const int* route_idx, const float* route_w, int topk, void* out, int ldo,
                                    int add_residual, void* stream) {
    GnRaw g;
    memset(&g, 0, sizeof(g));
    g.stats = gn2_stats; g.tiles = gn2_tiles; g.groups = gn2_groups; g.count = gn2_count; g.eps = gn2_eps; g.gamma = gamma2; g.beta = beta2;
    g.route_w = route_w;
    YM_CHECK_ARG(gn_raw_ok(g, C) && route_idx, "ym_moe_combine_tc_gn: bad GroupNorm-2 description (groups %d over %d channels)", gn2_groups, C);
    return moe_combine_tc_impl(x, ldx, B, HW, C, ws, bias_s, o, nullptr, nullptr, g, route_idx, topk, out, ldo, add_residual, stream);
}
