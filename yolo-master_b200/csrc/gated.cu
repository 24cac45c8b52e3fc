// Gated MoE family (VisualEnhancedAdaptiveGateMoE, nn/modules/moe/gated.py): the small fp32 kernels
// between the convolutions.  The kernel bodies are in gated_core.cuh (phase functions shared with the host-side check).
//   ym_gate_router    DualStreamGateRouter + batch complexity gate -> per-image top-k experts and weights
//   ym_fc_gate        SE gate / feature gate: two-layer MLP on a pooled vector -> per-(image, channel) multiplier
//   ym_gated_select   FusedExpertGroup tail: GroupNorm of the routed experts' channel slices, affine, SiLU, weighted sum
//   ym_ctx_mean3      PyramidContextMixer: mean of the local map and two nearest-upsampled pooled maps
#include "ym_common.cuh"

#include "gated_core.cuh"

namespace ym {
using namespace gated;

__global__ void __launch_bounds__(NTHR) gate_r0_kernel(const R0Args a) {   // grid = (slabs, B)
    YM_DYN_SMEM(float, sm);
    for (int ph = 0; ph < R0_PHASES; ++ph) {
        r0_phase(ph, a, blockIdx.y, blockIdx.x, threadIdx.x, NTHR, sm);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(NTHR) gate_r0m_kernel(const R0Args a, const R2Args r) {   // grid = B: merge, then the global stream
    YM_DYN_SMEM(float, sm);
    r0m_phase(a, blockIdx.x, threadIdx.x, NTHR);
    __syncthreads();
    for (int ph = 0; ph < G_PHASES; ++ph) {
        g_phase(ph, r, blockIdx.x, threadIdx.x, NTHR, sm);
        __syncthreads();
    }
}

// Slab statistics + pooled map of every image.
static void launch_r0(R0Args& a, int B, float* part, cudaStream_t st) {
    a.part = part;
    r0_slabs(a.Hp, &a.S, &a.PR);
    YM_LAUNCH(gate_r0_kernel, dim3(a.S, B), NTHR, r0_smem_floats(a.C, NTHR) * sizeof(float), st, a);
}

__global__ void __launch_bounds__(NTHR) gate_r1a_kernel(const R1Args a) {   // grid = (S1, B)
    YM_DYN_SMEM(float, sm);
    for (int ph = 0; ph < R1A_PHASES; ++ph) {
        r1a_phase(ph, a, blockIdx.y, blockIdx.x, threadIdx.x, NTHR, sm);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(NTHR) gate_r1b_kernel(const R1Args a) {   // grid = (S2, B)
    YM_DYN_SMEM(float, sm);
    for (int ph = 0; ph < R1B_PHASES; ++ph) {
        r1b_phase(ph, a, blockIdx.y, blockIdx.x, threadIdx.x, NTHR, sm);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(NTHR) gate_r1c_kernel(const R1Args a) {   // grid = B
    YM_DYN_SMEM(float, sm);
    for (int ph = 0; ph < R1_TAIL_PHASES; ++ph) {
        r1_tail_phase(ph, a, blockIdx.x, threadIdx.x, NTHR, sm);
        __syncthreads();
    }
}

// Local stream of every image: depthwise + GN1 partials, GN1 + 1x1 + GN2 partials, GN2 + head + spatial mean.
static void launch_r1(R1Args& a, int B, float* p1, cudaStream_t st) {
    r1_geom(a.Hp * a.Wp, a.C, &a.S1, &a.PS1, &a.S2, &a.PS2);
    a.p1 = p1;
    a.p2 = p1 + (long long)B * a.S1 * 2 * MAXG;
    YM_LAUNCH(gate_r1a_kernel, dim3(a.S1, B), NTHR, r1a_smem_floats(a.C, NTHR) * sizeof(float), st, a);
    YM_LAUNCH(gate_r1b_kernel, dim3(a.S2, B), NTHR, r1b_smem_floats(a.C, a.R, a.PS2) * sizeof(float), st, a);
    YM_LAUNCH(gate_r1c_kernel, B, NTHR, r1_smem_floats(a.R, a.E, NTHR) * sizeof(float), st, a);
}

__global__ void __launch_bounds__(NTHR) gate_r2_kernel(const R2Args a) {
    __shared__ float sm[4];
    for (int ph = 0; ph < R2_PHASES; ++ph) {
        r2_phase(ph, a, threadIdx.x, NTHR, sm);
        __syncthreads();
    }
}

// Per image: merge of the slab statistics, global stream logits and complexity; then the batch-level finish.
static void launch_finish(const R0Args& a0, const R2Args& a2, cudaStream_t st) {
    if (a2.zero_cost != 2) YM_LAUNCH(gate_r0m_kernel, a2.B, NTHR, g_smem_floats(a2.C, a2.E, NTHR) * sizeof(float), st, a0, a2);
    YM_LAUNCH(gate_r2_kernel, 1, NTHR, 0, st, a2);
}

__global__ void __launch_bounds__(NTHR) fc_gate_kernel(const FcArgs a) {
    YM_DYN_SMEM(float, sm);
    for (int ph = 0; ph < FC_PHASES; ++ph) {
        fc_phase(ph, a, blockIdx.x, threadIdx.x, NTHR, sm);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(NTHR) gap_kernel(const GapArgs a) {   // grid = (channel slabs, B)
    YM_DYN_SMEM(float, sm);
    for (int ph = 0; ph < GAP_PHASES; ++ph) {
        gap_phase(ph, a, blockIdx.y, blockIdx.x, threadIdx.x, NTHR, sm);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(NTHR) latent_router_kernel(const LrArgs a) {
    YM_DYN_SMEM(float, sm);
    for (int ph = 0; ph < LR_PHASES; ++ph) {
        lr_phase(ph, a, blockIdx.x, threadIdx.x, NTHR, sm);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(NTHR) classify_kernel(const ClsArgs a) {
    YM_DYN_SMEM(float, sm);
    for (int ph = 0; ph < CLS_PHASES; ++ph) {
        cls_phase(ph, a, blockIdx.x, threadIdx.x, NTHR, sm);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(NTHR) select_s0_kernel(const S0Args a) {   // grid = (slabs, routes)
    YM_DYN_SMEM(float, sm);
    for (int ph = 0; ph < S0_PHASES; ++ph) {
        s0_phase(ph, a, blockIdx.y, blockIdx.x, threadIdx.x, NTHR, sm);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(NTHR) select_s0m_kernel(const S0Args a) {   // grid = routes
    __shared__ float sm[2 * MAXG];
    for (int ph = 0; ph < S0M_PHASES; ++ph) {
        s0m_phase(ph, a, blockIdx.x, threadIdx.x, NTHR, sm);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) select_s1_kernel(const S0Args a, const float* __restrict__ w, __half* __restrict__ out,
                                                        int ldo, long long total) {
    const int cv = a.oc >> 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / cv;
        const int c = (int)(i - row * cv) << 3, b = (int)(row / a.HW), p = (int)(row - (long long)b * a.HW);
        Half8 o;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            o.v[j] = __floats2half2_rn(s1_element(a, w, b, p, c + 2 * j), s1_element(a, w, b, p, c + 2 * j + 1));
        *reinterpret_cast<Half8*>(out + row * ldo + c) = o;
    }
}

__global__ void __launch_bounds__(256) ctx_mean3_kernel(const CtxArgs a, __half* __restrict__ out, int ldo, long long total) {
    const int cv = a.C >> 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / cv;
        const int c = (int)(i - row * cv) << 3;
        const int img = (int)(row / ((long long)a.H * a.W)), rem = (int)(row - (long long)img * a.H * a.W);
        const int y = rem / a.W, x = rem - y * a.W;
        Half8 o;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            o.v[j] = __floats2half2_rn(ctx_element(a, img, y, x, c + 2 * j), ctx_element(a, img, y, x, c + 2 * j + 1));
        *reinterpret_cast<Half8*>(out + row * ldo + c) = o;
    }
}

static int grid_for(long long total) {
    long long nb = (total + 255) / 256;
    return (int)(nb > 148LL * 16 ? 148LL * 16 : (nb < 1 ? 1 : nb));
}

}  // namespace ym

using namespace ym;

static void pooled_dims(int H, int W, int pool, int* Hp, int* Wp, int* eff) {
    const bool pooling = pool > 1 && H > pool && W > pool;   // gated.py:139-142
    *eff = pooling ? pool : 1;
    *Hp = H / *eff;
    *Wp = W / *eff;
}

extern "C" long long ym_gate_router_scratch_floats(int B, int H, int W, int C, int R, int E, int pool) {
    int Hp, Wp, eff;
    pooled_dims(H, W, pool, &Hp, &Wp, &eff);
    const long long N = (long long)Hp * Wp;
    int S1, PS1, S2, PS2;
    r1_geom((int)N, C, &S1, &PS1, &S2, &PS2);
    return (long long)B * (2LL * C + N * C * 2 + N * R + E + 1 + 2LL * R0_MAX_SLABS * C + 2LL * MAXG * (S1 + S2) + E);
}

extern "C" int ym_gate_router(const void* x, int ldx, int B, int H, int W, int C, int pool, const float* global_fc,
                              const float* dw, const float* gn1_w, const float* gn1_b, int G1, const float* pw1, int R,
                              const float* gn2_w, const float* gn2_b, int G2, const float* pw2, const float* b2, int E,
                              float gn_eps, float alpha, float temperature, const float* cx_w, float cx_b, int topk,
                              const float* ln_w, const float* ln_b, float ln_eps, const float* prior, float* scratch,
                              float* w_out, int* idx_out, float* probs_out, void* stream) {
    YM_CHECK_ARG(x && scratch && w_out && idx_out, "ym_gate_router: null pointer");
    YM_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && C > 0 && R > 0, "ym_gate_router: empty problem");
    YM_CHECK_ARG(C % 8 == 0 && ldx % 8 == 0 && ldx >= C, "ym_gate_router: C and the pitch must be multiples of 8");
    YM_CHECK_ARG(E >= 1 && E <= MAXE && topk >= 1 && topk <= E, "ym_gate_router: 1 <= topk <= E <= %d", MAXE);
    YM_CHECK_ARG(G1 >= 1 && G1 <= MAXG && C % G1 == 0 && G2 >= 1 && G2 <= MAXG && R % G2 == 0, "ym_gate_router: GroupNorm groups");
    YM_CHECK_ARG(temperature > 0.f, "ym_gate_router: temperature must be positive");
    YM_CHECK_ARG((ln_w == nullptr) == (ln_b == nullptr), "ym_gate_router: LayerNorm weight and bias come together");
    int Hp, Wp, eff;
    pooled_dims(H, W, pool, &Hp, &Wp, &eff);
    const long long N = (long long)Hp * Wp;
    {
        int S1, PS1, S2, PS2;
        r1_geom((int)N, C, &S1, &PS1, &S2, &PS2);
        YM_CHECK_ARG(r1b_smem_floats(C, R, PS2) <= 12288 && r1a_smem_floats(C, NTHR) <= 12288, "%s: the local stream does not fit 48 KB of shared memory", "ym_gate_router");
    }
    float* stats = scratch;
    float* pooled = stats + (long long)B * 2 * C;
    float* t1 = pooled + (long long)B * N * C;
    float* t2 = t1 + (long long)B * N * C;
    float* ll = t2 + (long long)B * N * R;
    float* cx = ll + (long long)B * E;
    float* part = cx + B;
    cudaStream_t st = (cudaStream_t)stream;
    R0Args a0;
    a0.x = (const __half*)x; a0.ldx = ldx; a0.H = H; a0.W = W; a0.C = C; a0.pool = eff; a0.Hp = Hp; a0.Wp = Wp;
    a0.inv_area = 1.f / (float)(eff * eff); a0.stats = stats; a0.pooled = pooled;
    launch_r0(a0, B, part, st);
    R1Args a1;
    a1.pooled = pooled; a1.t1 = t1; a1.t2 = t2; a1.Hp = Hp; a1.Wp = Wp; a1.C = C; a1.R = R; a1.E = E; a1.G1 = G1; a1.G2 = G2;
    a1.eps = gn_eps; a1.dw = dw; a1.g1w = gn1_w; a1.g1b = gn1_b; a1.pw1 = pw1; a1.g2w = gn2_w; a1.g2b = gn2_b; a1.pw2 = pw2;
    a1.b2 = b2; a1.ll = ll; a1.pixel_softmax = 0; a1.inv_temp = 1.f;
    launch_r1(a1, B, part + (long long)B * 2 * R0_MAX_SLABS * C, st);
    R2Args a2;
    a2.stats = stats; a2.ll = ll; a2.wg = global_fc; a2.wc = cx_w; a2.bc = cx_b; a2.alpha = alpha; a2.inv_temp = 1.f / temperature;
    a2.B = B; a2.C = C; a2.E = E; a2.topk = topk; a2.zero_cost = 0; a2.w_min = 0.f; a2.cx = cx; a2.w = w_out; a2.probs = probs_out; a2.idx = idx_out;
    a2.gl = a1.p2 + (long long)B * a1.S2 * 2 * MAXG;
    a2.ln_w = ln_w; a2.ln_b = ln_b; a2.ln_eps = ln_eps; a2.prior = prior;
    launch_finish(a0, a2, st);
    YM_CHECK_LAUNCH("gate_router");
    return YM_OK;
}

extern "C" int ym_pixel_router(const void* x, int ldx, int B, int H, int W, int C, int pool, const float* dw, const float* gn1_w,
                               const float* gn1_b, int G1, const float* pw1, int R, const float* gn2_w, const float* gn2_b, int G2,
                               const float* pw2, const float* b2, int E, float gn_eps, float temperature, float w_min, int topk,
                               float* scratch, float* w_out, int* idx_out, float* probs_out, void* stream) {
    YM_CHECK_ARG(x && scratch && w_out && idx_out, "ym_pixel_router: null pointer");
    YM_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && C > 0 && R > 0 && R <= MAXR, "ym_pixel_router: empty problem / more than %d reduced channels", MAXR);
    YM_CHECK_ARG(C % 8 == 0 && ldx % 8 == 0 && ldx >= C, "ym_pixel_router: C and the pitch must be multiples of 8");
    YM_CHECK_ARG(E >= 1 && E <= 32 && topk >= 1 && topk <= E, "ym_pixel_router: 1 <= topk <= E <= 32");
    YM_CHECK_ARG(G1 >= 1 && G1 <= MAXG && C % G1 == 0 && G2 >= 1 && G2 <= MAXG && R % G2 == 0, "ym_pixel_router: GroupNorm groups");
    YM_CHECK_ARG(temperature > 0.f, "ym_pixel_router: temperature must be positive");
    int Hp, Wp, eff;
    pooled_dims(H, W, pool, &Hp, &Wp, &eff);
    const long long N = (long long)Hp * Wp;
    {
        int S1, PS1, S2, PS2;
        r1_geom((int)N, C, &S1, &PS1, &S2, &PS2);
        YM_CHECK_ARG(r1b_smem_floats(C, R, PS2) <= 12288 && r1a_smem_floats(C, NTHR) <= 12288, "%s: the local stream does not fit 48 KB of shared memory", "ym_pixel_router");
    }
    float* stats = scratch;
    float* pooled = stats + (long long)B * 2 * C;
    float* t1 = pooled + (long long)B * N * C;
    float* t2 = t1 + (long long)B * N * C;
    float* ll = t2 + (long long)B * N * R;
    float* part = ll + (long long)B * E + B;
    cudaStream_t st = (cudaStream_t)stream;
    R0Args a0;
    a0.x = (const __half*)x; a0.ldx = ldx; a0.H = H; a0.W = W; a0.C = C; a0.pool = eff; a0.Hp = Hp; a0.Wp = Wp;
    a0.inv_area = 1.f / (float)(eff * eff); a0.stats = stats; a0.pooled = pooled;
    launch_r0(a0, B, part, st);
    R1Args a1;
    a1.pooled = pooled; a1.t1 = t1; a1.t2 = t2; a1.Hp = Hp; a1.Wp = Wp; a1.C = C; a1.R = R; a1.E = E; a1.G1 = G1; a1.G2 = G2;
    a1.eps = gn_eps; a1.dw = dw; a1.g1w = gn1_w; a1.g1b = gn1_b; a1.pw1 = pw1; a1.g2w = gn2_w; a1.g2b = gn2_b; a1.pw2 = pw2;
    a1.b2 = b2; a1.ll = ll; a1.pixel_softmax = 1; a1.inv_temp = 1.f / temperature;
    launch_r1(a1, B, part + (long long)B * 2 * R0_MAX_SLABS * C, st);
    R2Args a2;
    a2.stats = stats; a2.ll = ll; a2.wg = nullptr; a2.wc = nullptr; a2.bc = 0.f; a2.alpha = 0.f; a2.inv_temp = 1.f;
    a2.B = B; a2.C = C; a2.E = E; a2.topk = topk; a2.zero_cost = 2; a2.w_min = w_min; a2.cx = nullptr; a2.w = w_out; a2.probs = probs_out;
    a2.gl = nullptr;
    a2.idx = idx_out; a2.ln_w = nullptr; a2.ln_b = nullptr; a2.ln_eps = 0.f; a2.prior = nullptr;
    launch_finish(a0, a2, st);
    YM_CHECK_LAUNCH("pixel_router");
    return YM_OK;
}

extern "C" long long ym_zero_cost_router_scratch_floats(int B, int C) { return (long long)B * (2LL * C + 1 + 2LL * R0_MAX_SLABS * C + MAXE); }

extern "C" int ym_zero_cost_router(const void* x, int ldx, int B, int H, int W, int C, const float* fc, int E, float temperature,
                                   const float* cx_w, float cx_b, int topk, float* scratch, float* w_out, int* idx_out,
                                   float* probs_out, void* stream) {
    YM_CHECK_ARG(x && fc && cx_w && scratch && w_out && idx_out, "ym_zero_cost_router: null pointer");
    YM_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && C > 0, "ym_zero_cost_router: empty problem");
    YM_CHECK_ARG(C % 8 == 0 && ldx % 8 == 0 && ldx >= C, "ym_zero_cost_router: C and the pitch must be multiples of 8");
    YM_CHECK_ARG(E >= 1 && E <= MAXE && topk >= 1 && topk <= E, "ym_zero_cost_router: 1 <= topk <= E <= %d", MAXE);
    YM_CHECK_ARG(temperature > 0.f, "ym_zero_cost_router: temperature must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    float* stats = scratch;
    float* cx = stats + (long long)B * 2 * C;
    float* part = cx + B;
    R0Args a0;
    a0.x = (const __half*)x; a0.ldx = ldx; a0.H = H; a0.W = W; a0.C = C; a0.pool = 1; a0.Hp = H; a0.Wp = W; a0.inv_area = 1.f;
    a0.stats = stats; a0.pooled = nullptr;
    launch_r0(a0, B, part, st);
    R2Args a2;
    a2.stats = stats; a2.ll = nullptr; a2.wg = fc; a2.wc = cx_w; a2.bc = cx_b; a2.alpha = 1.f; a2.inv_temp = 1.f / temperature;
    a2.B = B; a2.C = C; a2.E = E; a2.topk = topk; a2.zero_cost = 1; a2.w_min = 0.f; a2.cx = cx; a2.w = w_out; a2.probs = probs_out; a2.idx = idx_out;
    a2.gl = part + (long long)B * 2 * R0_MAX_SLABS * C;
    a2.ln_w = nullptr; a2.ln_b = nullptr; a2.ln_eps = 0.f; a2.prior = nullptr;
    launch_finish(a0, a2, st);
    YM_CHECK_LAUNCH("zero_cost_router");
    return YM_OK;
}

extern "C" int ym_fc_gate(const void* v, int ldv, int B, int Cin, const float* w1, int Cr, const float* w2, const float* b2,
                          int Cout, float scale, float offset, float* out, void* stream) {
    YM_CHECK_ARG(v && w1 && w2 && out, "ym_fc_gate: null pointer");
    YM_CHECK_ARG(B > 0 && Cin > 0 && Cr > 0 && Cr <= 8192 && Cout > 0 && ldv >= Cin, "ym_fc_gate: bad sizes");
    FcArgs a;
    a.v = (const __half*)v; a.ldv = ldv; a.Cin = Cin; a.Cr = Cr; a.Cout = Cout; a.w1 = w1; a.w2 = w2; a.b2 = b2; a.scale = scale;
    a.offset = offset; a.out = out;
    YM_LAUNCH(fc_gate_kernel, B, NTHR, fc_smem_floats(Cr) * sizeof(float), (cudaStream_t)stream, a);
    YM_CHECK_LAUNCH("fc_gate");
    return YM_OK;
}

extern "C" int ym_gap_nhwc(const void* x, int ldx, int B, int HW, int C, void* out, int ldo, void* stream) {
    YM_CHECK_ARG(x && out, "ym_gap_nhwc: null pointer");
    YM_CHECK_ARG(B > 0 && B <= 65535 && HW > 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldx >= C && ldo >= C, "ym_gap_nhwc: bad sizes (C and the pitch multiples of 8)");
    GapArgs a;
    a.x = (const __half*)x; a.ldx = ldx; a.HW = HW; a.C = C; a.out = (__half*)out; a.ldo = ldo;
    YM_LAUNCH(gap_kernel, dim3((C + GAP_SLAB - 1) / GAP_SLAB, B), NTHR, gap_smem_floats(NTHR) * sizeof(float), (cudaStream_t)stream, a);
    YM_CHECK_LAUNCH("gap_nhwc");
    return YM_OK;
}

extern "C" int ym_latent_router(int T, const void* const* tokens, const int* lds, int B, int C, const float* emb, const float* ln_w,
                                const float* ln_b, float ln_eps, const float* w1, const float* b1, int hid, const float* w2,
                                const float* b2, const float* wh, const float* bh, int E, float temperature, float* logits,
                                float* probs, void* stream) {
    YM_CHECK_ARG(tokens && lds && ln_w && ln_b && w1 && b1 && w2 && b2 && wh && bh && logits && probs, "ym_latent_router: null pointer");
    YM_CHECK_ARG(T >= 1 && T <= LR_MAX_TOKENS && B > 0 && C > 0 && hid > 0 && E >= 1 && E <= MAXE, "ym_latent_router: 1..%d tokens, E <= %d",
                 LR_MAX_TOKENS, MAXE);
    YM_CHECK_ARG(lr_smem_floats(C, hid, E, NTHR) * sizeof(float) <= 48 * 1024, "ym_latent_router: latent / hidden width too large");
    LrArgs a;
    for (int t = 0; t < T; ++t) {
        YM_CHECK_ARG(tokens[t] && lds[t] >= C, "ym_latent_router: bad token %d", t);
        a.tok[t] = (const __half*)tokens[t];
        a.ld[t] = lds[t];
    }
    a.T = T; a.C = C; a.hid = hid; a.E = E; a.emb = emb; a.ln_w = ln_w; a.ln_b = ln_b; a.ln_eps = ln_eps;
    a.inv_temp = 1.f / (temperature < 0.1f ? 0.1f : temperature);
    a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = b2; a.wh = wh; a.bh = bh; a.logits = logits; a.probs = probs;
    YM_LAUNCH(latent_router_kernel, B, NTHR, lr_smem_floats(C, hid, E, NTHR) * sizeof(float), (cudaStream_t)stream, a);
    YM_CHECK_LAUNCH("latent_router");
    return YM_OK;
}

extern "C" int ym_classify_head(const void* v, int ldv, int B, int Cin, const float* w, const float* b, int nc, float* logits,
                                float* probs, void* stream) {
    YM_CHECK_ARG(v && w && logits && probs, "ym_classify_head: null pointer");
    YM_CHECK_ARG(B > 0 && Cin > 0 && nc > 0 && ldv >= Cin, "ym_classify_head: bad sizes");
    ClsArgs a;
    a.v = (const __half*)v; a.ldv = ldv; a.Cin = Cin; a.nc = nc; a.w = w; a.b = b; a.logits = logits; a.probs = probs;
    YM_LAUNCH(classify_kernel, B, NTHR, cls_smem_floats(NTHR) * sizeof(float), (cudaStream_t)stream, a);
    YM_CHECK_LAUNCH("classify_head");
    return YM_OK;
}

extern "C" long long ym_gated_select_scratch_floats(int B, int topk, int oc) { return s0_scratch_floats(B, topk, oc); }

extern "C" int ym_gated_select(const void* fo, int ldf, int B, int HW, int E, int oc, int G, float eps, const int* idx,
                               const float* w, int topk, const float* gamma, const float* beta, float* scratch, void* out,
                               int ldo, void* stream) {
    YM_CHECK_ARG(fo && idx && w && gamma && beta && scratch && out, "ym_gated_select: null pointer");
    YM_CHECK_ARG(B > 0 && HW > 0 && E >= 1 && topk >= 1 && topk <= E && (long long)B * topk <= 65535, "ym_gated_select: bad sizes");
    YM_CHECK_ARG(oc % 8 == 0 && oc <= 8 * NTHR && ldo % 8 == 0 && ldf % 8 == 0 && ldf >= E * oc, "ym_gated_select: oc, ldf, ldo multiples of 8; ldf >= E*oc");
    YM_CHECK_ARG(G >= 1 && G <= MAXG && oc % G == 0, "ym_gated_select: GroupNorm groups");
    S0Args a;
    a.fo = (const __half*)fo; a.ldf = ldf; a.HW = HW; a.oc = oc; a.G = G; a.topk = topk; a.eps = eps; a.idx = idx; a.gamma = gamma;
    a.beta = beta; a.sc = scratch; a.sh = scratch + (long long)B * topk * oc; a.part = a.sh + (long long)B * topk * oc;
    s0_slabs(HW, &a.S, &a.PS);
    cudaStream_t st = (cudaStream_t)stream;
    YM_LAUNCH(select_s0_kernel, dim3(a.S, B * topk), NTHR, s0_smem_floats(oc, NTHR) * sizeof(float), st, a);
    YM_LAUNCH(select_s0m_kernel, B * topk, NTHR, 0, st, a);
    const long long total = (long long)B * HW * (oc / 8);
    YM_LAUNCH(select_s1_kernel, grid_for(total), 256, 0, st, a, w, (__half*)out, ldo, total);
    YM_CHECK_LAUNCH("gated_select");
    return YM_OK;
}

extern "C" int ym_ctx_mean3(const void* a, int lda, const void* b, int ldb, const void* c, int ldc, int B, int H, int W, int C,
                            int h2, int w2, int h4, int w4, void* out, int ldo, void* stream) {
    YM_CHECK_ARG(a && b && c && out, "ym_ctx_mean3: null pointer");
    YM_CHECK_ARG(B > 0 && H > 0 && W > 0 && h2 > 0 && w2 > 0 && h4 > 0 && w4 > 0, "ym_ctx_mean3: empty map");
    YM_CHECK_ARG(C % 8 == 0 && ldo % 8 == 0, "ym_ctx_mean3: C, ldo multiples of 8");
    CtxArgs g;
    g.a = (const __half*)a; g.b = (const __half*)b; g.c = (const __half*)c; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.H = H; g.W = W; g.C = C; g.h2 = h2; g.w2 = w2; g.h4 = h4; g.w4 = w4;
    g.sy2 = (float)h2 / (float)H; g.sx2 = (float)w2 / (float)W; g.sy4 = (float)h4 / (float)H; g.sx4 = (float)w4 / (float)W;
    const long long total = (long long)B * H * W * (C / 8);
    YM_LAUNCH(ctx_mean3_kernel, grid_for(total), 256, 0, (cudaStream_t)stream, g, (__half*)out, ldo, total);
    YM_CHECK_LAUNCH("ctx_mean3");
    return YM_OK;
}
