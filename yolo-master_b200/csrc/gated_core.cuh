// Gated MoE family (nn/modules/moe/gated.py): bodies of the small fp32 kernels in gated.cu.
//
// Every kernel here is "one CTA per image (or per route), a fixed sequence of phases separated by __syncthreads()", and each
// phase is written as a function of (tid, nthr) that only reads what earlier phases wrote.  That makes the bodies runnable on
// the host by looping tid inside each phase (tests/native/gated_host.cpp, g++), so the arithmetic is compared with the oracle
// in the GPU-less build container.  No warp intrinsics, no atomics: reductions go through shared memory in a fixed order.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__) || defined(YM_HOST_EMU)   // nvcc, or g++ with tests/native/cuda_host_emu.h already included
#ifdef __CUDACC__
#include <cuda_fp16.h>
#endif
#ifndef YM_HD
#ifdef __CUDACC__
#define YM_HD __host__ __device__ __forceinline__
#else
#define YM_HD inline
#endif
#endif
typedef __half ym_half;
YM_HD float ym_h2f(ym_half h) { return __half2float(h); }
YM_HD ym_half ym_f2h(float f) { return __float2half_rn(f); }
#else
#ifndef YM_HD
#define YM_HD inline
#endif
typedef _Float16 ym_half;
YM_HD float ym_h2f(ym_half h) { return (float)h; }
YM_HD ym_half ym_f2h(float f) { return (ym_half)f; }
#endif

namespace ym {
namespace gated {

constexpr int NTHR = 256;      // threads per CTA of every kernel in this file
constexpr int MAXG = 8;        // GroupNorm groups (get_safe_groups(c, 8) <= 8)
constexpr int MAXE = 64;       // experts

YM_HD float sigmoid_f(float v) { return 1.f / (1.f + expf(-v)); }
YM_HD float silu_f32(float v) { return v / (1.f + expf(-v)); }

// Thread t of a CTA owns "column" c0 = t % Cg and row lane pl = t / Cg of a [rows][cols] reduction (Cg = min(cols, nthr)).
struct ColLane {
    int Cg, NL, c0, pl;
    YM_HD ColLane(int cols, int tid, int nthr) {
        Cg = cols < nthr ? cols : nthr;
        NL = nthr / Cg;
        c0 = tid % Cg;
        pl = tid / Cg;
    }
    YM_HD bool active() const { return pl < NL; }
};
YM_HD int col_lane_floats(int cols, int nthr) { return cols > nthr ? cols : nthr; }   // NL * cols <= max(cols, nthr)

// --------------------------------------------------------------------------------------------------------------------
// R0: per-image channel statistics (mean, population std) of x and the pooled fp32 map the local stream reads
//     (DualStreamGateRouter.forward gated.py:129-142).  The map is cut into S slabs of whole pooled rows, one CTA each (grid S x B):
//     the CTA reduces its slab to (mean_s, M2_s) per channel in two passes over data it has just pulled into L1/L2 and writes its pooled
//     rows; r0m_phase then merges the S partials of an image in slab order (Chan's parallel-variance update), so the result does not
//     depend on scheduling.  Eight channels per thread per load (C and the pitch are multiples of 8, as for every activation here).
struct R0Args {
    const ym_half* x;   // [B][H*W][ldx]
    int ldx, H, W, C, pool, Hp, Wp;
    float inv_area;     // 1 / (pool*pool), 1 when the map is not pooled
    float* stats;       // [B][2C]: mean | std
    float* pooled;      // [B][Hp*Wp][C], or null: statistics only
    float* part;        // [B][S][2C]: slab mean | slab M2
    int S, PR;          // slabs per image, pooled rows per slab (the last slab also takes the H % pool leftover rows)
};
constexpr int R0_MAX_SLABS = 32;
constexpr int R0_PHASES = 5;
YM_HD void r0_slabs(int Hp, int* S, int* PR) {
    const int want = Hp < R0_MAX_SLABS ? Hp : R0_MAX_SLABS;
    *PR = (Hp + want - 1) / want;
    *S = (Hp + *PR - 1) / *PR;
}
YM_HD int r0_smem_floats(int C, int nthr) { return (C > nthr * 8 ? C : nthr * 8) + C; }

YM_HD void ym_load8(const ym_half* p, float (&f)[8]) {
#if defined(__CUDA_ARCH__)
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 v = __half22float2(h[j]);
        f[2 * j] = v.x;
        f[2 * j + 1] = v.y;
    }
#else
    for (int j = 0; j < 8; ++j) f[j] = ym_h2f(p[j]);
#endif
}

YM_HD void r0_phase(int ph, const R0Args& a, int img, int slab, int tid, int nthr, float* sm) {
    const int C = a.C, OC = C >> 3;
    const int y0 = slab * a.PR * a.pool, y1 = slab == a.S - 1 ? a.H : (slab + 1) * a.PR * a.pool;
    const int n = (y1 - y0) * a.W;                                     // pixels of this slab
    const ym_half* x = a.x + ((long long)img * a.H * a.W + (long long)y0 * a.W) * a.ldx;
    float* part = sm;
    float* mean = sm + (C > nthr * 8 ? C : nthr * 8);
    const ColLane cl(OC, tid, nthr);                                   // column = channel octet, lane = pixel lane
    float* slab_out = a.part + ((long long)img * a.S + slab) * 2 * C;
    if (ph == 0 || ph == 2) {   // partial sums of x (ph 0) or (x - slab mean)^2 (ph 2)
        if (!cl.active()) return;
        for (int o = cl.c0; o < OC; o += cl.Cg) {
            float m[8], acc[8], v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                m[j] = ph == 2 ? mean[o * 8 + j] : 0.f;
                acc[j] = 0.f;
            }
            for (int p = cl.pl; p < n; p += cl.NL) {
                ym_load8(x + (long long)p * a.ldx + o * 8, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d = v[j] - m[j];
                    acc[j] += ph == 2 ? d * d : d;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) part[cl.pl * C + o * 8 + j] = acc[j];
        }
    } else if (ph == 1 || ph == 3) {
        for (int c = tid; c < C; c += nthr) {
            float s = 0.f;
            for (int l = 0; l < cl.NL; ++l) s += part[l * C + c];
            if (ph == 1) {
                s /= (float)n;
                mean[c] = s;
                slab_out[c] = s;
            } else {
                slab_out[C + c] = s;
            }
        }
    } else {   // ph 4: avg_pool2d(kernel = stride = pool), floor mode, or a plain fp32 copy - this slab's pooled rows
        if (!a.pooled) return;                                        // statistics only (ZeroCostRouter)
        const int r0 = slab * a.PR, r1 = (slab + 1) * a.PR < a.Hp ? (slab + 1) * a.PR : a.Hp;
        float* out = a.pooled + ((long long)img * a.Hp + r0) * a.Wp * C;
        const int ne = (r1 - r0) * a.Wp * OC;
        for (int e = tid; e < ne; e += nthr) {
            const int o = e % OC, pp = e / OC, py = pp / a.Wp, px = pp % a.Wp;
            float s[8], v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] = 0.f;
            for (int dy = 0; dy < a.pool; ++dy)
                for (int dx = 0; dx < a.pool; ++dx) {
                    ym_load8(x + (long long)((py * a.pool + dy) * a.W + px * a.pool + dx) * a.ldx + o * 8, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) s[j] += v[j];
                }
#pragma unroll
            for (int j = 0; j < 8; ++j) out[(long long)pp * C + o * 8 + j] = s[j] * a.inv_area;
        }
    }
}

// Merge of the slab partials of one image, in slab order: n <- n + n_s, delta = mean_s - mean, mean += delta n_s / n,
// M2 += M2_s + delta^2 n_old n_s / n.
YM_HD void r0m_phase(const R0Args& a, int img, int tid, int nthr) {
    const int C = a.C;
    for (int c = tid; c < C; c += nthr) {
        float cnt = 0.f, mean = 0.f, m2 = 0.f;
        for (int s = 0; s < a.S; ++s) {
            const int y0 = s * a.PR * a.pool, y1 = s == a.S - 1 ? a.H : (s + 1) * a.PR * a.pool;
            const float ns = (float)((y1 - y0) * a.W);
            const float* ps = a.part + ((long long)img * a.S + s) * 2 * C;
            const float tot = cnt + ns, delta = ps[c] - mean;
            mean += delta * (ns / tot);
            m2 += ps[C + c] + delta * delta * (cnt * ns / tot);
            cnt = tot;
        }
        a.stats[(long long)img * 2 * C + c] = mean;
        a.stats[(long long)img * 2 * C + C + c] = cnt > 1.f ? sqrtf(m2 / cnt) : 0.f;
    }
}

// --------------------------------------------------------------------------------------------------------------------
// R1: local stream on the pooled map (gated.py:109-118,144-145): dw3x3 -> GN -> SiLU -> 1x1 -> GN -> SiLU -> 1x1 + bias -> mean.
struct R1Args {
    const float* pooled;   // [B][N][C]
    float *t1, *t2;        // scratch [B][N][C], [B][N][R]
    int Hp, Wp, C, R, E, G1, G2;
    float eps;
    const float *dw;       // [C][9]
    const float *g1w, *g1b;   // [C]
    const float* pw1;      // [R][C]
    const float *g2w, *g2b;   // [R]
    const float *pw2, *b2;    // [E][R], [E]
    float* ll;             // [B][E] local logits (pixel_softmax: the spatial mean of the per-pixel expert probabilities)
    int pixel_softmax;     // 1: UltraEfficientRouter routers.py:104-118 - per-pixel clamp(+-30) / T -> softmax over experts, THEN the mean
    float inv_temp;
    float *p1, *p2;        // GroupNorm slab partials [B][S1][2*G1], [B][S2][2*G2]: slab mean | slab M2 per group
    int S1, PS1, S2, PS2;  // pixel slabs of the depthwise kernel / of the 1x1 kernel, pixels per slab
};
constexpr int MAXR = 64;       // reduced channels of the local stream (max(C / 16, 4))
constexpr int R1A_PHASES = 6, R1B_PHASES = 7;
// Slab geometry of the local stream: ~16 pixels per depthwise CTA (at most 32 slabs), and for the 1x1 as many pixels as keep the
// activated [pixels][C+1] tile within 32 KB of shared memory (at most 32).
YM_HD void r1_geom(int N, int C, int* S1, int* PS1, int* S2, int* PS2) {
    int want = (N + 15) / 16;
    want = want < 1 ? 1 : (want > 32 ? 32 : want);
    *PS1 = (N + want - 1) / want;
    *S1 = (N + *PS1 - 1) / *PS1;
    int ps = 8192 / (C + 1);
    ps = ps < 1 ? 1 : (ps > 32 ? 32 : ps);
    ps = ps > N ? N : ps;
    *PS2 = ps;
    *S2 = (N + ps - 1) / ps;
}
YM_HD int r1a_smem_floats(int C, int nthr) { return col_lane_floats(C, nthr) + C + MAXG; }
YM_HD int r1b_smem_floats(int C, int R, int PS2) { return PS2 * (C + 1) + PS2 * R + R + 3 * MAXG; }
YM_HD int r1_part_floats(int E, int nthr) { return nthr * (E > MAXG ? E : MAXG); }      // per-thread partials per expert (pixel softmax)
YM_HD int r1_smem_floats(int R, int E, int nthr) { return r1_part_floats(E, nthr) + 2 * MAXG + col_lane_floats(R, nthr); }

// Slab partials [S][2G] of group g merged in slab order (Chan): slab s holds min(PS, N - s*PS) pixels x cpg channels.
YM_HD void slab_merge(const float* part, int S, int G, int g, int N, int PS, int cpg, float* mean, float* var) {
    float cnt = 0.f, mu = 0.f, m2 = 0.f;
    for (int s = 0; s < S; ++s) {
        const int n = N - s * PS < PS ? N - s * PS : PS;
        const float ns = (float)n * cpg, tot = cnt + ns, delta = part[s * 2 * G + g] - mu;
        mu += delta * (ns / tot);
        m2 += part[s * 2 * G + G + g] + delta * delta * (cnt * ns / tot);
        cnt = tot;
    }
    *mean = mu;
    *var = m2 / cnt;
}

// R1a, CTA (slab, image): depthwise 3x3 (zero padding) of the slab's pixels -> t1, and the slab's GroupNorm-1 partials.
YM_HD void r1a_phase(int ph, const R1Args& a, int img, int slab, int tid, int nthr, float* sm) {
    const int N = a.Hp * a.Wp, C = a.C, cpg = C / a.G1;
    const int p0 = slab * a.PS1, p1 = p0 + a.PS1 < N ? p0 + a.PS1 : N, n = p1 - p0;
    const float* src = a.pooled + (long long)img * N * C;
    float* t1 = a.t1 + (long long)img * N * C;
    float* part = sm;                                  // [lanes][C]
    float* chs = sm + col_lane_floats(C, nthr);        // [C]
    float* mean = chs + C;                             // [G1]
    float* out = a.p1 + ((long long)img * a.S1 + slab) * 2 * a.G1;
    const ColLane cl(C, tid, nthr);
    if (ph == 0 || ph == 3) {
        if (!cl.active()) return;
        for (int c = cl.c0; c < C; c += cl.Cg) {
            float acc = 0.f;
            if (ph == 0) {
                float w[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) w[k] = a.dw[c * 9 + k];
                for (int p = p0 + cl.pl; p < p1; p += cl.NL) {
                    const int y = p / a.Wp, x = p % a.Wp;
                    float s = 0.f;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int yy = y + ky - 1;
                        if (yy < 0 || yy >= a.Hp) continue;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const int xx = x + kx - 1;
                            if (xx < 0 || xx >= a.Wp) continue;
                            s += w[ky * 3 + kx] * src[(long long)(yy * a.Wp + xx) * C + c];
                        }
                    }
                    t1[(long long)p * C + c] = s;
                    acc += s;
                }
            } else {
                const float m = mean[c / cpg];
                for (int p = p0 + cl.pl; p < p1; p += cl.NL) {
                    const float d = t1[(long long)p * C + c] - m;
                    acc += d * d;
                }
            }
            part[cl.pl * C + c] = acc;
        }
    } else if (ph == 1 || ph == 4) {
        for (int c = tid; c < C; c += nthr) {
            float s = 0.f;
            for (int l = 0; l < cl.NL; ++l) s += part[l * C + c];
            chs[c] = s;
        }
    } else if (tid < a.G1) {
        float s = 0.f;
        for (int c = tid * cpg; c < (tid + 1) * cpg; ++c) s += chs[c];
        if (ph == 2) {
            s /= (float)n * cpg;
            mean[tid] = s;
            out[tid] = s;
        } else {
            out[a.G1 + tid] = s;
        }
    }
}

// R1b, CTA (slab, image): GroupNorm-1 (merged partials) + SiLU of the slab's pixels into a shared tile, the 1x1 C -> R from the tile
// -> t2, and the slab's GroupNorm-2 partials.  Lanes of a warp walk the pixels of one output channel: the weight row is a broadcast,
// the tile rows are one bank apart (pitch C + 1).
YM_HD void r1b_phase(int ph, const R1Args& a, int img, int slab, int tid, int nthr, float* sm) {
    const int N = a.Hp * a.Wp, C = a.C, R = a.R, cpg1 = C / a.G1, cpg2 = R / a.G2, TP = C + 1;
    const int p0 = slab * a.PS2, p1 = p0 + a.PS2 < N ? p0 + a.PS2 : N, n = p1 - p0;
    const float* t1 = a.t1 + ((long long)img * N + p0) * C;
    float* t2 = a.t2 + ((long long)img * N + p0) * R;
    float* tile = sm;                          // [PS2][C+1]
    float* o2 = tile + a.PS2 * TP;             // [PS2][R]
    float* chs = o2 + a.PS2 * R;               // [R]
    float* mean1 = chs + R;
    float* rstd1 = mean1 + MAXG;
    float* mean2 = rstd1 + MAXG;
    float* out = a.p2 + ((long long)img * a.S2 + slab) * 2 * a.G2;
    switch (ph) {
        case 0:
            if (tid < a.G1) {
                float mu, var;
                slab_merge(a.p1 + (long long)img * a.S1 * 2 * a.G1, a.S1, a.G1, tid, N, a.PS1, cpg1, &mu, &var);
                mean1[tid] = mu;
                rstd1[tid] = 1.f / sqrtf(var + a.eps);
            }
            break;
        case 1:
            for (int e = tid; e < n * C; e += nthr) {
                const int c = e % C, p = e / C, g = c / cpg1;
                tile[p * TP + c] = silu_f32((t1[e] - mean1[g]) * rstd1[g] * a.g1w[c] + a.g1b[c]);
            }
            break;
        case 2:
            for (int e = tid; e < n * R; e += nthr) {
                const int p = e % n, r = e / n;
                const float* row = tile + p * TP;
                const float* w = a.pw1 + (long long)r * C;
                float s = 0.f;
                for (int c = 0; c < C; ++c) s += w[c] * row[c];
                t2[(long long)p * R + r] = s;
                o2[p * R + r] = s;
            }
            break;
        case 3:
        case 5:
            for (int r = tid; r < R; r += nthr) {
                const float m = ph == 5 ? mean2[r / cpg2] : 0.f;
                float s = 0.f;
                for (int p = 0; p < n; ++p) {
                    const float d = o2[p * R + r] - m;
                    s += ph == 5 ? d * d : d;
                }
                chs[r] = s;
            }
            break;
        default:   // 4: slab mean of group tid, 6: slab M2
            if (tid < a.G2) {
                float s = 0.f;
                for (int r = tid * cpg2; r < (tid + 1) * cpg2; ++r) s += chs[r];
                if (ph == 4) {
                    s /= (float)n * cpg2;
                    mean2[tid] = s;
                    out[tid] = s;
                } else {
                    out[a.G2 + tid] = s;
                }
            }
            break;
    }
}
// R1c, CTA (image): GroupNorm-2 from the merged partials, SiLU, the last 1x1 and the spatial mean (or the per-pixel softmax variant).
constexpr int R1_TAIL_PHASES = 3;
YM_HD void r1_tail_phase(int ph, const R1Args& a, int img, int tid, int nthr, float* sm) {
    const int N = a.Hp * a.Wp, R = a.R;
    const float* t2 = a.t2 + (long long)img * N * R;
    float* part = sm;
    float* mean = sm + r1_part_floats(a.E, nthr);
    float* rstd = mean + MAXG;
    float* colp = rstd + MAXG;
    const int cpg2 = R / a.G2;
    if (ph == 0) {
        if (tid < a.G2) {
            float mu, var;
            slab_merge(a.p2 + (long long)img * a.S2 * 2 * a.G2, a.S2, a.G2, tid, N, a.PS2, cpg2, &mu, &var);
            mean[tid] = mu;
            rstd[tid] = 1.f / sqrtf(var + a.eps);
        }
    } else if (ph == 1 && a.pixel_softmax) {   // per pixel: SiLU(GN2) -> 1x1 + bias -> clamp -> / T -> softmax; per-thread sums per expert
        float acc[MAXE];
        for (int e = 0; e < a.E; ++e) acc[e] = 0.f;
        for (int p = tid; p < N; p += nthr) {
            float h[MAXR], l[MAXE];
            for (int r = 0; r < R; ++r) {
                const int g = r / cpg2;
                h[r] = silu_f32((t2[(long long)p * R + r] - mean[g]) * rstd[g] * a.g2w[r] + a.g2b[r]);
            }
            float mx = -3.0e38f;
            for (int e = 0; e < a.E; ++e) {
                float s = a.b2[e];
                for (int r = 0; r < R; ++r) s += a.pw2[e * R + r] * h[r];
                s = (s < -30.f ? -30.f : (s > 30.f ? 30.f : s)) * a.inv_temp;
                l[e] = s;
                mx = s > mx ? s : mx;
            }
            float den = 0.f;
            for (int e = 0; e < a.E; ++e) {
                l[e] = expf(l[e] - mx);
                den += l[e];
            }
            for (int e = 0; e < a.E; ++e) acc[e] += l[e] / den;
        }
        for (int e = 0; e < a.E; ++e) part[tid * a.E + e] = acc[e];
    } else if (ph == 1) {   // column sums over the pixels of SiLU(GN2(t2)): the last 1x1 and the spatial mean commute
        const ColLane cl(R, tid, nthr);
        if (!cl.active()) return;
        for (int r = cl.c0; r < R; r += cl.Cg) {
            const int g = r / cpg2;
            float s = 0.f;
            for (int p = cl.pl; p < N; p += cl.NL) s += silu_f32((t2[(long long)p * R + r] - mean[g]) * rstd[g] * a.g2w[r] + a.g2b[r]);
            colp[cl.pl * R + r] = s;
        }
    } else if (a.pixel_softmax) {
        for (int e = tid; e < a.E; e += nthr) {
            float s = 0.f;
            for (int t = 0; t < nthr; ++t) s += part[t * a.E + e];
            a.ll[(long long)img * a.E + e] = s / (float)N;
        }
    } else {
        const ColLane cl(R, tid, nthr);
        for (int e = tid; e < a.E; e += nthr) {
            float s = a.b2[e];
            for (int r = 0; r < R; ++r) {
                float m = 0.f;
                for (int l = 0; l < cl.NL; ++l) m += colp[l * R + r];
                s += a.pw2[e * R + r] * (m / (float)N);
            }
            a.ll[(long long)img * a.E + e] = s;
        }
    }
}

// --------------------------------------------------------------------------------------------------------------------
// R2 (one CTA for the whole batch): batch-level complexity scalar (gated.py:455-461), stream blend, softmax(/T), top-k,
// renormalisation (gated.py:147-151) and the complexity gate (gated.py:469-490).
struct R2Args {
    const float* stats;   // [B][2C]
    const float* ll;      // [B][E]
    const float* wg;      // [E][2C] global_fc
    const float *wc;      // [C] complexity_estimator conv weight
    float bc, alpha, inv_temp;   // alpha = sigmoid(self.alpha)
    const float *ln_w, *ln_b;    // [2C] LayerNorm over the statistics in front of global_fc (DualStreamGateRouterV2 gated.py:215,226), or null
    float ln_eps;
    const float* prior;          // [E] learnable expert prior added to the blended logits before the clamp (gated.py:216,238), or null
    int B, C, E, topk;
    int zero_cost;        // 0: DualStreamGateRouter + complexity GATE (ranks dropped);  1: ZeroCostRouter (gated.py:953-968: softmax of
                          //    the global stream, / T, clamp, softmax again) + complexity SCALE (weights multiplied, modules.py:1663-1670);
                          // 2: `ll` already holds expert probabilities (UltraEfficientRouter): top-k, w / max(sum, 1e-6), no complexity,
                          //    weights <= w_min zeroed (BatchedExpertComputation's eval threshold, moe/utils.py:172-173)
    float w_min;
    float* cx;            // scratch [B]: per-image complexity, written by g_phase
    float* gl;            // scratch [B][E]: global-stream logits, written by g_phase
    float *w, *probs;     // [B][topk], [B][E] (probs nullable)
    int* idx;             // [B][topk]
};
constexpr int R2_PHASES = 2;
YM_HD int r2_smem_floats() { return 4; }

// G, CTA (image), runs behind the statistics merge in the same kernel: the optional LayerNorm of the [mean | std] row, the global
// stream's E dot products of length 2C and the complexity estimator's dot product of length C.  One output row per warp-sized group
// of 32 threads (coalesced weight rows), partials reduced through shared memory in lane order.
constexpr int G_PHASES = 7;
YM_HD int g_smem_floats(int C, int E, int nthr) { return nthr + 2 + 2 * C + (E + 1) * 32; }

YM_HD void g_phase(int ph, const R2Args& a, int img, int tid, int nthr, float* sm) {
    const int C2 = 2 * a.C;
    const float* st = a.stats + (long long)img * C2;
    float* part = sm;              // [nthr]
    float* lnp = sm + nthr;        // mean, rstd of the row
    float* row = lnp + 2;          // [2C] the (normalised) statistics
    float* rowp = row + C2;        // [(E+1)][32]
    const bool ln = a.ln_w != nullptr;
    switch (ph) {
        case 0:
        case 2: {
            if (!ln) return;
            const float m = ph == 2 ? lnp[0] : 0.f;
            float s = 0.f;
            for (int j = tid; j < C2; j += nthr) {
                const float d = st[j] - m;
                s += ph == 2 ? d * d : d;
            }
            part[tid] = s;
            break;
        }
        case 1:
        case 3: {
            if (!ln || tid != 0) return;
            float s = 0.f;
            for (int t = 0; t < nthr; ++t) s += part[t];
            s /= (float)C2;
            if (ph == 1) lnp[0] = s;
            else lnp[1] = 1.f / sqrtf(s + a.ln_eps);
            break;
        }
        case 4:
            for (int j = tid; j < C2; j += nthr) row[j] = ln ? (st[j] - lnp[0]) * lnp[1] * a.ln_w[j] + a.ln_b[j] : st[j];
            break;
        case 5: {
            const int lane = tid & 31, nrow = nthr >> 5;
            for (int r = tid >> 5; r <= a.E; r += nrow) {
                float s = 0.f;
                if (r < a.E) {
                    if (a.wg)
                        for (int j = lane; j < C2; j += 32) s += a.wg[(long long)r * C2 + j] * row[j];
                } else if (a.wc) {
                    for (int c = lane; c < a.C; c += 32) s += a.wc[c] * st[c];
                }
                rowp[r * 32 + lane] = s;
            }
            break;
        }
        default: {
            if (tid > a.E) return;
            float s = 0.f;
            for (int l = 0; l < 32; ++l) s += rowp[tid * 32 + l];
            if (tid < a.E) {
                if (a.wg) a.gl[(long long)img * a.E + tid] = s;
            } else if (a.wc) {
                a.cx[img] = sigmoid_f(a.bc + s);
            }
            break;
        }
    }
}

YM_HD void r2_phase(int ph, const R2Args& a, int tid, int nthr, float* sm) {
    if (ph == 0) {
        if (tid != 0 || a.zero_cost == 2) return;
        float s = 0.f;
        for (int b = 0; b < a.B; ++b) s += a.cx[b];
        s /= (float)a.B;
        if (!(s == s) || s > 3.0e38f || s < -3.0e38f) s = 1.f;      // non-finite -> 1.0
        s = s < 0.3f ? 0.3f : (s > 1.5f ? 1.5f : s);
        float keep = rintf(s * (float)a.topk);                        // torch.round: half to even
        keep = keep < 1.f ? 1.f : (keep > (float)a.topk ? (float)a.topk : keep);
        sm[0] = keep;
        sm[1] = s;                                                    // the clamped complexity itself (zero-cost mode scales by it)
    } else {
        const float keep = sm[0], cscale = sm[1];
        for (int b = tid; b < a.B; b += nthr) {
            float p[MAXE];
            if (a.zero_cost == 2) {   // probabilities come straight from the per-pixel stream
                float wsel2[MAXE], tot = 0.f;
                for (int e = 0; e < a.E; ++e) {
                    p[e] = a.ll[(long long)b * a.E + e];
                    if (a.probs) a.probs[(long long)b * a.E + e] = p[e];
                }
                for (int j = 0; j < a.topk; ++j) {
                    int best = -1;
                    for (int e = 0; e < a.E; ++e)
                        if (p[e] >= 0.f && (best < 0 || p[e] > p[best])) best = e;
                    if (best < 0) best = j;
                    a.idx[(long long)b * a.topk + j] = best;
                    wsel2[j] = p[best];
                    tot += p[best];
                    p[best] = -1.f;
                }
                for (int j = 0; j < a.topk; ++j) {
                    const float v = wsel2[j] / (tot < 1e-6f ? 1e-6f : tot);
                    a.w[(long long)b * a.topk + j] = v > a.w_min ? v : 0.f;
                }
                continue;
            }
            float mx = -3.0e38f;
            for (int e = 0; e < a.E; ++e) {
                const float gl = a.gl[(long long)b * a.E + e];
                float l = a.zero_cost ? gl : a.alpha * gl + (1.f - a.alpha) * a.ll[(long long)b * a.E + e];
                if (a.prior) l += a.prior[e];
                if (!a.zero_cost) {
                    l = l < -30.f ? -30.f : (l > 30.f ? 30.f : l);
                    l *= a.inv_temp;
                }
                p[e] = l;
                mx = p[e] > mx ? p[e] : mx;
            }
            float den = 0.f;
            for (int e = 0; e < a.E; ++e) {
                p[e] = expf(p[e] - mx);
                den += p[e];
            }
            for (int e = 0; e < a.E; ++e) p[e] /= den;
            if (a.zero_cost) {   // the Sequential already ends in a Softmax: its output / T, clamped, goes through softmax again
                mx = -3.0e38f;
                for (int e = 0; e < a.E; ++e) {
                    float l = p[e] * a.inv_temp;
                    l = l < -30.f ? -30.f : (l > 30.f ? 30.f : l);
                    p[e] = l;
                    mx = l > mx ? l : mx;
                }
                den = 0.f;
                for (int e = 0; e < a.E; ++e) {
                    p[e] = expf(p[e] - mx);
                    den += p[e];
                }
                for (int e = 0; e < a.E; ++e) p[e] /= den;
            }
            if (a.probs)
                for (int e = 0; e < a.E; ++e) a.probs[(long long)b * a.E + e] = p[e];
            float wsel[MAXE];
            float tot = 0.f;
            for (int j = 0; j < a.topk; ++j) {
                int best = -1;
                for (int e = 0; e < a.E; ++e)
                    if (p[e] >= 0.f && (best < 0 || p[e] > p[best])) best = e;
                if (best < 0) best = j;                               // all-NaN row: keep the indices valid
                a.idx[(long long)b * a.topk + j] = best;
                wsel[j] = p[best];
                tot += p[best];
                p[best] = -1.f;                                       // taken
            }
            float tot2 = 0.f;
            for (int j = 0; j < a.topk; ++j) {
                wsel[j] /= tot + 1e-6f;
                if (!a.zero_cost && a.topk > 1 && (float)(j + 1) > keep) wsel[j] = 0.f;
                tot2 += wsel[j];
            }
            for (int j = 0; j < a.topk; ++j) {
                float v = wsel[j];
                if (a.zero_cost) v *= cscale;
                else if (a.topk > 1) v /= tot2 < 1e-6f ? 1e-6f : tot2;
                a.w[(long long)b * a.topk + j] = v;
            }
        }
    }
}

// --------------------------------------------------------------------------------------------------------------------
// Per-image two-layer gate on a pooled vector: out = offset + scale * sigmoid(W2 . silu(W1 . v) + b2)
// (se_gate gated.py:325-332; feature_gate moe/hooks.py:50-57 with scale = tanh(refine_scale); CrossPathGate gated.py:2402-2404 with
// offset 0.5 and scale 0.5 * tanh(gate_scale)).
struct FcArgs {
    const ym_half* v;   // [B][ldv]
    int ldv, Cin, Cr, Cout;
    const float *w1, *w2, *b2;   // [Cr][Cin], [Cout][Cr], [Cout] (nullable)
    float scale, offset;
    float* out;         // [B][Cout]
};
constexpr int FC_PHASES = 2;
YM_HD int fc_smem_floats(int Cr) { return Cr; }

YM_HD void fc_phase(int ph, const FcArgs& a, int img, int tid, int nthr, float* sm) {
    const ym_half* v = a.v + (long long)img * a.ldv;
    if (ph == 0) {
        for (int r = tid; r < a.Cr; r += nthr) {
            float s = 0.f;
            for (int c = 0; c < a.Cin; ++c) s += a.w1[(long long)r * a.Cin + c] * ym_h2f(v[c]);
            sm[r] = silu_f32(s);
        }
    } else {
        for (int o = tid; o < a.Cout; o += nthr) {
            float s = a.b2 ? a.b2[o] : 0.f;
            for (int r = 0; r < a.Cr; ++r) s += a.w2[(long long)o * a.Cr + r] * sm[r];
            a.out[(long long)img * a.Cout + o] = a.offset + a.scale * sigmoid_f(s);
        }
    }
}

// --------------------------------------------------------------------------------------------------------------------
// Global average pool of an NHWC fp16 map -> fp16 [B][ldo] (nn.AdaptiveAvgPool2d(1) of the SE / feature / cross gates, the latent
// tokens and Classify).  One CTA per (image, 64-channel slab): 8 channel octets x 32 pixel lanes, 16-byte loads, fixed-order reduction.
struct GapArgs {
    const ym_half* x;   // [B][HW][ldx]
    int ldx, HW, C;
    ym_half* out;       // [B][ldo]
    int ldo;
};
constexpr int GAP_SLAB = 64;     // channels per CTA
constexpr int GAP_PHASES = 2;
YM_HD int gap_smem_floats(int nthr) { return nthr * 8; }

YM_HD void gap_phase(int ph, const GapArgs& a, int img, int slab, int tid, int nthr, float* sm) {
    const int c0 = slab * GAP_SLAB, oct = tid & 7, lane = tid >> 3, nl = nthr >> 3;    // 8 octets x nl pixel lanes
    const int c = c0 + oct * 8;
    if (ph == 0) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        if (c < a.C) {
            const ym_half* base = a.x + (long long)img * a.HW * a.ldx + c;
            float v[8];
            for (int p = lane; p < a.HW; p += nl) {
                ym_load8(base + (long long)p * a.ldx, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += v[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sm[tid * 8 + j] = acc[j];
    } else {
        if (tid >= GAP_SLAB || c0 + tid >= a.C) return;
        const int o = tid >> 3, j = tid & 7;                  // channel c0 + tid lives in octet o, slot j
        float s = 0.f;
        for (int l = 0; l < nl; ++l) s += sm[(l * 8 + o) * 8 + j];
        a.out[(long long)img * a.ldo + c0 + tid] = ym_f2h(s / (float)a.HW);
    }
}

// --------------------------------------------------------------------------------------------------------------------
// LatentRouter.forward (nn/modules/latent_mixture.py:219-241, per_token = False) on the pooled scale tokens of LatentMixture: mean over
// tokens of (token + scale embedding) -> LayerNorm -> Linear + SiLU -> Linear + SiLU -> expert head -> clamp(+-30) -> softmax(/ T).
constexpr int LR_MAX_TOKENS = 4;
struct LrArgs {
    const ym_half* tok[LR_MAX_TOKENS];   // [B][ld] pooled token vectors (fp16, as adaptive_avg_pool2d of an fp16 map leaves them)
    int ld[LR_MAX_TOKENS];
    int T, C, hid, E;
    const float* emb;                    // [T][C] scale embedding, or null
    const float *ln_w, *ln_b;            // [C]
    float ln_eps, inv_temp;              // inv_temp = 1 / max(temperature, 0.1)
    const float *w1, *b1, *w2, *b2, *wh, *bh;   // [hid][C],[hid]  [C][hid],[C]  [E][C],[E]
    float *logits, *probs;               // [B][E]
};
constexpr int LR_PHASES = 9;
YM_HD int lr_smem_floats(int C, int hid, int E, int nthr) { return 2 * C + hid + E + nthr + 2; }

YM_HD void lr_phase(int ph, const LrArgs& a, int img, int tid, int nthr, float* sm) {
    const int C = a.C;
    float* r = sm;
    float* h2 = r + C;
    float* h1 = h2 + C;
    float* lg = h1 + a.hid;
    float* red = lg + a.E;
    float* st = red + nthr;
    switch (ph) {
        case 0: case 2: {
            float s = 0.f;
            for (int c = tid; c < C; c += nthr) {
                if (ph == 0) {
                    float v = 0.f;
                    for (int t = 0; t < a.T; ++t) v += ym_h2f(a.tok[t][(long long)img * a.ld[t] + c]) + (a.emb ? a.emb[t * C + c] : 0.f);
                    r[c] = v / (float)a.T;
                    s += r[c];
                } else {
                    const float d = r[c] - st[0];
                    s += d * d;
                }
            }
            red[tid] = s;
            break;
        }
        case 1: case 3: {
            if (tid != 0) return;
            float s = 0.f;
            for (int t = 0; t < nthr; ++t) s += red[t];
            if (ph == 1) st[0] = s / (float)C;
            else st[1] = 1.f / sqrtf(s / (float)C + a.ln_eps);
            break;
        }
        case 4:
            for (int c = tid; c < C; c += nthr) r[c] = (r[c] - st[0]) * st[1] * a.ln_w[c] + a.ln_b[c];
            break;
        case 5:
            for (int j = tid; j < a.hid; j += nthr) {
                float s = a.b1[j];
                for (int c = 0; c < C; ++c) s += a.w1[(long long)j * C + c] * r[c];
                h1[j] = silu_f32(s);
            }
            break;
        case 6:
            for (int c = tid; c < C; c += nthr) {
                float s = a.b2[c];
                for (int j = 0; j < a.hid; ++j) s += a.w2[(long long)c * a.hid + j] * h1[j];
                h2[c] = silu_f32(s);
            }
            break;
        case 7:
            for (int e = tid; e < a.E; e += nthr) {
                float s = a.bh[e];
                for (int c = 0; c < C; ++c) s += a.wh[(long long)e * C + c] * h2[c];
                if (!(s == s)) s = 0.f;                                   // nan_to_num(nan = 0), then the +-30 clamp covers +-inf
                lg[e] = s < -30.f ? -30.f : (s > 30.f ? 30.f : s);
            }
            break;
        default: {
            if (tid != 0) return;
            float mx = -3.0e38f, den = 0.f;
            for (int e = 0; e < a.E; ++e) mx = lg[e] * a.inv_temp > mx ? lg[e] * a.inv_temp : mx;
            for (int e = 0; e < a.E; ++e) den += expf(lg[e] * a.inv_temp - mx);
            for (int e = 0; e < a.E; ++e) {
                a.logits[(long long)img * a.E + e] = lg[e];
                a.probs[(long long)img * a.E + e] = expf(lg[e] * a.inv_temp - mx) / den;
            }
        }
    }
}

// --------------------------------------------------------------------------------------------------------------------
// Classify tail (nn/modules/head.py:823-832): logits = W . v + b on the pooled vector, probs = softmax(logits).  One CTA per image.
struct ClsArgs {
    const ym_half* v;   // [B][ldv] pooled features
    int ldv, Cin, nc;
    const float *w, *b;   // [nc][Cin], [nc]
    float *logits, *probs;   // [B][nc]
};
constexpr int CLS_PHASES = 4;
YM_HD int cls_smem_floats(int nthr) { return nthr + 2; }

YM_HD void cls_phase(int ph, const ClsArgs& a, int img, int tid, int nthr, float* sm) {
    const ym_half* v = a.v + (long long)img * a.ldv;
    float* lg = a.logits + (long long)img * a.nc;
    float* pr = a.probs + (long long)img * a.nc;
    if (ph == 0) {          // logits, per-thread running maximum
        float mx = -3.0e38f;
        for (int o = tid; o < a.nc; o += nthr) {
            float s = a.b ? a.b[o] : 0.f;
            const float* w = a.w + (long long)o * a.Cin;
            for (int c = 0; c < a.Cin; ++c) s += w[c] * ym_h2f(v[c]);
            lg[o] = s;
            mx = s > mx ? s : mx;
        }
        sm[tid] = mx;
    } else if (ph == 1) {   // row maximum
        if (tid != 0) return;
        float mx = sm[0];
        for (int t = 1; t < nthr; ++t) mx = sm[t] > mx ? sm[t] : mx;
        sm[nthr] = mx;
    } else if (ph == 2) {   // exponentials, per-thread partial sums (barrier before sm[tid] is overwritten: sm[nthr] already read)
        const float mx = sm[nthr];
        float s = 0.f;
        for (int o = tid; o < a.nc; o += nthr) {
            const float e = expf(lg[o] - mx);
            pr[o] = e;
            s += e;
        }
        sm[tid] = s;
    } else {                // normalise; every thread re-adds the partial sums in the same fixed order
        float den = 0.f;
        for (int t = 0; t < nthr; ++t) den += sm[t];
        for (int o = tid; o < a.nc; o += nthr) pr[o] /= den;
    }
}

// --------------------------------------------------------------------------------------------------------------------
// FusedExpertGroup tail (gated.py:1061-1081): for route (b, j) with expert e = idx[b][j], GroupNorm (no affine) over the
// expert's channel slice of the all-expert conv output, then the expert's affine; S0 produces per-(route, channel) scale/shift.
struct S0Args {
    const ym_half* fo;   // [B][HW][ldf], expert e occupies channels [e*oc, (e+1)*oc)
    int ldf, HW, oc, G, topk;
    float eps;
    const int* idx;      // [B][topk]
    const float *gamma, *beta;   // [E][oc]
    float *sc, *sh;      // [B*topk][oc]
    float* part;         // [B*topk][S][2G]: slab mean | slab M2 per group
    int S, PS;           // pixel slabs per route, pixels per slab
};
constexpr int S0_MAX_SLABS = 32;
constexpr int S0_PHASES = 6;
constexpr int S0M_PHASES = 2;
YM_HD void s0_slabs(int HW, int* S, int* PS) {   // about 256 pixels per CTA
    int want = (HW + 255) / 256;
    want = want < 1 ? 1 : (want > S0_MAX_SLABS ? S0_MAX_SLABS : want);
    *PS = (HW + want - 1) / want;
    *S = (HW + *PS - 1) / *PS;
}
YM_HD int s0_smem_floats(int oc, int nthr) { return nthr * 8 + oc + 2 * MAXG; }   // oc <= 8 * nthr
YM_HD long long s0_scratch_floats(int B, int topk, int oc) { return (long long)B * topk * (2LL * oc + 2LL * S0_MAX_SLABS * MAXG); }

// Slab CTA (route, slab): per-group (mean, M2) of the slab's pixels of the selected expert's channel slice, two passes.
YM_HD void s0_phase(int ph, const S0Args& a, int route, int slab, int tid, int nthr, float* sm) {
    const int b = route / a.topk, ex = a.idx[route], cpg = a.oc / a.G, OC = a.oc >> 3;
    const int p0 = slab * a.PS, p1 = p0 + a.PS < a.HW ? p0 + a.PS : a.HW, n = p1 - p0;
    const ym_half* base = a.fo + ((long long)b * a.HW + p0) * a.ldf + (long long)ex * a.oc;
    float* part = sm;                 // [lanes][oc]
    float* chs = sm + nthr * 8;       // [oc]
    float* mean = chs + a.oc;         // [G]
    const ColLane cl(OC, tid, nthr);
    if (ph == 0 || ph == 3) {
        if (!cl.active()) return;
        for (int o = cl.c0; o < OC; o += cl.Cg) {
            float m[8], acc[8], v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                m[j] = ph == 3 ? mean[(o * 8 + j) / cpg] : 0.f;
                acc[j] = 0.f;
            }
            for (int p = cl.pl; p < n; p += cl.NL) {
                ym_load8(base + (long long)p * a.ldf + o * 8, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d = v[j] - m[j];
                    acc[j] += ph == 3 ? d * d : d;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) part[cl.pl * a.oc + o * 8 + j] = acc[j];
        }
    } else if (ph == 1 || ph == 4) {
        for (int c = tid; c < a.oc; c += nthr) {
            float s = 0.f;
            for (int l = 0; l < cl.NL; ++l) s += part[l * a.oc + c];
            chs[c] = s;
        }
    } else if (tid < a.G) {           // ph 2: slab mean of group tid; ph 5: slab M2, both to the partials
        float s = 0.f;
        for (int c = tid * cpg; c < (tid + 1) * cpg; ++c) s += chs[c];
        float* out = a.part + ((long long)route * a.S + slab) * 2 * a.G;
        if (ph == 2) {
            s /= (float)n * cpg;
            mean[tid] = s;
            out[tid] = s;
        } else {
            out[a.G + tid] = s;
        }
    }
}

// Merge CTA (route): slab partials in slab order (Chan), then the per-channel scale / shift of GroupNorm + the expert's affine.
YM_HD void s0m_phase(int ph, const S0Args& a, int route, int tid, int nthr, float* sm) {
    const int ex = a.idx[route], cpg = a.oc / a.G;
    float* mean = sm;
    float* rstd = sm + MAXG;
    if (ph == 0) {
        if (tid >= a.G) return;
        float cnt = 0.f, mu = 0.f, m2 = 0.f;
        for (int s = 0; s < a.S; ++s) {
            const int p0 = s * a.PS, p1 = p0 + a.PS < a.HW ? p0 + a.PS : a.HW;
            const float ns = (float)(p1 - p0) * cpg;
            const float* ps = a.part + ((long long)route * a.S + s) * 2 * a.G;
            const float tot = cnt + ns, delta = ps[tid] - mu;
            mu += delta * (ns / tot);
            m2 += ps[a.G + tid] + delta * delta * (cnt * ns / tot);
            cnt = tot;
        }
        mean[tid] = mu;
        rstd[tid] = 1.f / sqrtf(m2 / cnt + a.eps);
    } else {
        for (int c = tid; c < a.oc; c += nthr) {
            const int g = c / cpg;
            const float gm = a.gamma[(long long)ex * a.oc + c];
            a.sc[(long long)route * a.oc + c] = rstd[g] * gm;
            a.sh[(long long)route * a.oc + c] = a.beta[(long long)ex * a.oc + c] - mean[g] * rstd[g] * gm;
        }
    }
}

// S1, one output element: sum_j w[b][j] * SiLU(fo[b, p, e_j*oc + c] * sc[route][c] + sh[route][c]).
YM_HD float s1_element(const S0Args& a, const float* w, int b, int p, int c) {
    float s = 0.f;
    for (int j = 0; j < a.topk; ++j) {
        const int route = b * a.topk + j, ex = a.idx[route];
        const float v = ym_h2f(a.fo[((long long)b * a.HW + p) * a.ldf + (long long)ex * a.oc + c]);
        s += w[route] * silu_f32(v * a.sc[(long long)route * a.oc + c] + a.sh[(long long)route * a.oc + c]);
    }
    return s;
}

// --------------------------------------------------------------------------------------------------------------------
// PyramidContextMixer mean of the three context maps (gated.py:1213-1219): a full-res, b and c at (h2,w2) / (h4,w4), upsampled
// with F.interpolate(mode="nearest"): src = min(floor(dst * in / out), in - 1), the ratio in float32.
struct CtxArgs {
    const ym_half *a, *b, *c;
    int lda, ldb, ldc, H, W, C, h2, w2, h4, w4;
    float sy2, sx2, sy4, sx4;   // (float)in / out per axis
};
YM_HD int nearest_src(int dst, float scale, int in) {
    const int s = (int)floorf((float)dst * scale);
    return s < in - 1 ? s : in - 1;
}
YM_HD float ctx_element(const CtxArgs& a, int img, int y, int x, int ch) {
    const float va = ym_h2f(a.a[((long long)(img * a.H + y) * a.W + x) * a.lda + ch]);
    const int y2 = nearest_src(y, a.sy2, a.h2), x2 = nearest_src(x, a.sx2, a.w2);
    const int y4 = nearest_src(y, a.sy4, a.h4), x4 = nearest_src(x, a.sx4, a.w4);
    const float vb = ym_h2f(a.b[((long long)(img * a.h2 + y2) * a.w2 + x2) * a.ldb + ch]);
    const float vc = ym_h2f(a.c[((long long)(img * a.h4 + y4) * a.w4 + x4) * a.ldc + ch]);
    return (va + vb + vc) / 3.f;
}

}  // namespace gated
}  // namespace ym
