// TEST INFRASTRUCTURE (g++ only, no CUDA): runs the phase functions of yolo-master_b200/csrc/gated_core.cuh on the host - for
// each CTA, every phase is executed for tid = 0..NTHR-1 before the next phase starts, which is what __syncthreads() gives the
// kernels in gated.cu - so tests/test_gated_host.py can compare the arithmetic of ym_gate_router / ym_fc_gate / ym_gated_select /
// ym_ctx_mean3 with the oracle in the GPU-less build container.  Same argument lists as the C ABI, minus the stream.
#include <vector>

#include "gated_core.cuh"

using namespace ym::gated;

// gate_r1a / r1b / r1c kernels as launch_r1 in gated.cu issues them.
static void host_r1(R1Args& a, int B, std::vector<float>& p12) {
    r1_geom(a.Hp * a.Wp, a.C, &a.S1, &a.PS1, &a.S2, &a.PS2);
    p12.assign((size_t)B * (a.S1 + a.S2) * 2 * MAXG, 0.f);
    a.p1 = p12.data();
    a.p2 = a.p1 + (size_t)B * a.S1 * 2 * MAXG;
    std::vector<float> sm(r1a_smem_floats(a.C, NTHR) + r1b_smem_floats(a.C, a.R, a.PS2) + r1_smem_floats(a.R, a.E, NTHR));
    for (int b = 0; b < B; ++b)
        for (int s = 0; s < a.S1; ++s)
            for (int ph = 0; ph < R1A_PHASES; ++ph)
                for (int t = 0; t < NTHR; ++t) r1a_phase(ph, a, b, s, t, NTHR, sm.data());
    for (int b = 0; b < B; ++b)
        for (int s = 0; s < a.S2; ++s)
            for (int ph = 0; ph < R1B_PHASES; ++ph)
                for (int t = 0; t < NTHR; ++t) r1b_phase(ph, a, b, s, t, NTHR, sm.data());
    for (int b = 0; b < B; ++b)
        for (int ph = 0; ph < R1_TAIL_PHASES; ++ph)
            for (int t = 0; t < NTHR; ++t) r1_tail_phase(ph, a, b, t, NTHR, sm.data());
}

// gate_r0_kernel (grid = slabs x images), as launch_r0 in gated.cu issues it.
static void host_r0(R0Args& a, int B, std::vector<float>& part, float* sm) {
    r0_slabs(a.Hp, &a.S, &a.PR);
    part.assign((size_t)B * a.S * 2 * a.C, 0.f);
    a.part = part.data();
    for (int b = 0; b < B; ++b)
        for (int s = 0; s < a.S; ++s)
            for (int ph = 0; ph < R0_PHASES; ++ph)
                for (int t = 0; t < NTHR; ++t) r0_phase(ph, a, b, s, t, NTHR, sm);
}

// gate_r0m_kernel (merge + global stream, grid = images) and gate_r2_kernel (one CTA), as launch_finish in gated.cu issues them.
static void host_finish(const R0Args& a0, R2Args& a2, std::vector<float>& gl) {
    gl.assign((size_t)a2.B * a2.E, 0.f);
    a2.gl = gl.data();
    std::vector<float> sm(g_smem_floats(a2.C, a2.E, NTHR) + 16);
    if (a2.zero_cost != 2)
        for (int b = 0; b < a2.B; ++b) {
            for (int t = 0; t < NTHR; ++t) r0m_phase(a0, b, t, NTHR);
            for (int ph = 0; ph < G_PHASES; ++ph)
                for (int t = 0; t < NTHR; ++t) g_phase(ph, a2, b, t, NTHR, sm.data());
        }
    for (int ph = 0; ph < R2_PHASES; ++ph)
        for (int t = 0; t < NTHR; ++t) r2_phase(ph, a2, t, NTHR, sm.data());
}

extern "C" int host_gate_router(const void* x, int ldx, int B, int H, int W, int C, int pool, const float* global_fc,
                                const float* dw, const float* gn1_w, const float* gn1_b, int G1, const float* pw1, int R,
                                const float* gn2_w, const float* gn2_b, int G2, const float* pw2, const float* b2, int E,
                                float gn_eps, float alpha, float temperature, const float* cx_w, float cx_b, int topk,
                                const float* ln_w, const float* ln_b, float ln_eps, const float* prior, float* w_out, int* idx_out,
                                float* probs_out) {
    const bool pooling = pool > 1 && H > pool && W > pool;
    const int eff = pooling ? pool : 1, Hp = H / eff, Wp = W / eff;
    const long long N = (long long)Hp * Wp;
    std::vector<float> stats((size_t)B * 2 * C), pooled((size_t)B * N * C), t1((size_t)B * N * C), t2((size_t)B * N * R),
        ll((size_t)B * E), cx(B);
    R0Args a0;
    a0.x = (const ym_half*)x; a0.ldx = ldx; a0.H = H; a0.W = W; a0.C = C; a0.pool = eff; a0.Hp = Hp; a0.Wp = Wp;
    a0.inv_area = 1.f / (float)(eff * eff); a0.stats = stats.data(); a0.pooled = pooled.data();
    std::vector<float> sm(r0_smem_floats(C, NTHR) + r1_smem_floats(R, E, NTHR) + 16);
    std::vector<float> part;
    host_r0(a0, B, part, sm.data());
    R1Args a1;
    a1.pooled = pooled.data(); a1.t1 = t1.data(); a1.t2 = t2.data(); a1.Hp = Hp; a1.Wp = Wp; a1.C = C; a1.R = R; a1.E = E;
    a1.G1 = G1; a1.G2 = G2; a1.eps = gn_eps; a1.dw = dw; a1.g1w = gn1_w; a1.g1b = gn1_b; a1.pw1 = pw1; a1.g2w = gn2_w;
    a1.g2b = gn2_b; a1.pw2 = pw2; a1.b2 = b2; a1.ll = ll.data(); a1.pixel_softmax = 0; a1.inv_temp = 1.f;
    std::vector<float> p12;
    host_r1(a1, B, p12);
    R2Args a2;
    a2.stats = stats.data(); a2.ll = ll.data(); a2.wg = global_fc; a2.wc = cx_w; a2.bc = cx_b; a2.alpha = alpha;
    a2.inv_temp = 1.f / temperature; a2.B = B; a2.C = C; a2.E = E; a2.topk = topk; a2.zero_cost = 0; a2.w_min = 0.f; a2.cx = cx.data(); a2.w = w_out;
    a2.probs = probs_out; a2.idx = idx_out; a2.ln_w = ln_w; a2.ln_b = ln_b; a2.ln_eps = ln_eps; a2.prior = prior;
    std::vector<float> gl;
    host_finish(a0, a2, gl);
    return 0;
}

extern "C" int host_pixel_router(const void* x, int ldx, int B, int H, int W, int C, int pool, const float* dw, const float* gn1_w,
                                 const float* gn1_b, int G1, const float* pw1, int R, const float* gn2_w, const float* gn2_b, int G2,
                                 const float* pw2, const float* b2, int E, float gn_eps, float temperature, float w_min, int topk,
                                 float* w_out, int* idx_out, float* probs_out) {
    const bool pooling = pool > 1 && H > pool && W > pool;
    const int eff = pooling ? pool : 1, Hp = H / eff, Wp = W / eff;
    const long long N = (long long)Hp * Wp;
    std::vector<float> stats((size_t)B * 2 * C), pooled((size_t)B * N * C), t1((size_t)B * N * C), t2((size_t)B * N * R), ll((size_t)B * E);
    std::vector<float> sm(r0_smem_floats(C, NTHR) + r1_smem_floats(R, E, NTHR) + 16);
    R0Args a0;
    a0.x = (const ym_half*)x; a0.ldx = ldx; a0.H = H; a0.W = W; a0.C = C; a0.pool = eff; a0.Hp = Hp; a0.Wp = Wp;
    a0.inv_area = 1.f / (float)(eff * eff); a0.stats = stats.data(); a0.pooled = pooled.data();
    R1Args a1;
    a1.pooled = pooled.data(); a1.t1 = t1.data(); a1.t2 = t2.data(); a1.Hp = Hp; a1.Wp = Wp; a1.C = C; a1.R = R; a1.E = E;
    a1.G1 = G1; a1.G2 = G2; a1.eps = gn_eps; a1.dw = dw; a1.g1w = gn1_w; a1.g1b = gn1_b; a1.pw1 = pw1; a1.g2w = gn2_w;
    a1.g2b = gn2_b; a1.pw2 = pw2; a1.b2 = b2; a1.ll = ll.data(); a1.pixel_softmax = 1; a1.inv_temp = 1.f / temperature;
    std::vector<float> part;
    host_r0(a0, B, part, sm.data());
    std::vector<float> p12;
    host_r1(a1, B, p12);
    R2Args a2;
    a2.stats = stats.data(); a2.ll = ll.data(); a2.wg = nullptr; a2.wc = nullptr; a2.bc = 0.f; a2.alpha = 0.f; a2.inv_temp = 1.f;
    a2.B = B; a2.C = C; a2.E = E; a2.topk = topk; a2.zero_cost = 2; a2.w_min = w_min; a2.cx = nullptr; a2.w = w_out; a2.probs = probs_out;
    a2.idx = idx_out; a2.ln_w = nullptr; a2.ln_b = nullptr; a2.ln_eps = 0.f; a2.prior = nullptr;
    std::vector<float> gl;
    host_finish(a0, a2, gl);
    return 0;
}

extern "C" int host_zero_cost_router(const void* x, int ldx, int B, int H, int W, int C, const float* fc, int E, float temperature,
                                     const float* cx_w, float cx_b, int topk, float* w_out, int* idx_out, float* probs_out) {
    std::vector<float> stats((size_t)B * 2 * C), cx(B), sm(r0_smem_floats(C, NTHR) + 16);
    R0Args a0;
    a0.x = (const ym_half*)x; a0.ldx = ldx; a0.H = H; a0.W = W; a0.C = C; a0.pool = 1; a0.Hp = H; a0.Wp = W; a0.inv_area = 1.f;
    a0.stats = stats.data(); a0.pooled = nullptr;
    std::vector<float> part;
    host_r0(a0, B, part, sm.data());
    R2Args a2;
    a2.stats = stats.data(); a2.ll = nullptr; a2.wg = fc; a2.wc = cx_w; a2.bc = cx_b; a2.alpha = 1.f; a2.inv_temp = 1.f / temperature;
    a2.B = B; a2.C = C; a2.E = E; a2.topk = topk; a2.zero_cost = 1; a2.w_min = 0.f; a2.cx = cx.data(); a2.w = w_out; a2.probs = probs_out;
    a2.idx = idx_out; a2.ln_w = nullptr; a2.ln_b = nullptr; a2.ln_eps = 0.f; a2.prior = nullptr;
    std::vector<float> gl;
    host_finish(a0, a2, gl);
    return 0;
}

extern "C" int host_fc_gate(const void* v, int ldv, int B, int Cin, const float* w1, int Cr, const float* w2, const float* b2,
                            int Cout, float scale, float offset, float* out) {
    FcArgs a;
    a.v = (const ym_half*)v; a.ldv = ldv; a.Cin = Cin; a.Cr = Cr; a.Cout = Cout; a.w1 = w1; a.w2 = w2; a.b2 = b2; a.scale = scale;
    a.offset = offset; a.out = out;
    std::vector<float> sm(fc_smem_floats(Cr) + 1);
    for (int b = 0; b < B; ++b)
        for (int ph = 0; ph < FC_PHASES; ++ph)
            for (int t = 0; t < NTHR; ++t) fc_phase(ph, a, b, t, NTHR, sm.data());
    return 0;
}

extern "C" int host_classify_head(const void* v, int ldv, int B, int Cin, const float* w, const float* b, int nc, float* logits,
                                  float* probs) {
    ClsArgs a;
    a.v = (const ym_half*)v; a.ldv = ldv; a.Cin = Cin; a.nc = nc; a.w = w; a.b = b; a.logits = logits; a.probs = probs;
    std::vector<float> sm(cls_smem_floats(NTHR));
    for (int i = 0; i < B; ++i)
        for (int ph = 0; ph < CLS_PHASES; ++ph)
            for (int t = 0; t < NTHR; ++t) cls_phase(ph, a, i, t, NTHR, sm.data());
    return 0;
}

extern "C" int host_gated_select(const void* fo, int ldf, int B, int HW, int E, int oc, int G, float eps, const int* idx,
                                 const float* w, int topk, const float* gamma, const float* beta, void* out, int ldo) {
    std::vector<float> sc((size_t)B * topk * oc), sh((size_t)B * topk * oc), sm(s0_smem_floats(oc, NTHR));
    S0Args a;
    s0_slabs(HW, &a.S, &a.PS);
    std::vector<float> part((size_t)B * topk * a.S * 2 * G);
    a.part = part.data();
    a.fo = (const ym_half*)fo; a.ldf = ldf; a.HW = HW; a.oc = oc; a.G = G; a.topk = topk; a.eps = eps; a.idx = idx; a.gamma = gamma;
    a.beta = beta; a.sc = sc.data(); a.sh = sh.data();
    for (int r = 0; r < B * topk; ++r)
        for (int s = 0; s < a.S; ++s)
            for (int ph = 0; ph < S0_PHASES; ++ph)
                for (int t = 0; t < NTHR; ++t) s0_phase(ph, a, r, s, t, NTHR, sm.data());
    for (int r = 0; r < B * topk; ++r)
        for (int ph = 0; ph < S0M_PHASES; ++ph)
            for (int t = 0; t < NTHR; ++t) s0m_phase(ph, a, r, t, NTHR, sm.data());
    ym_half* o = (ym_half*)out;
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < HW; ++p)
            for (int c = 0; c < oc; ++c) o[((long long)b * HW + p) * ldo + c] = ym_f2h(s1_element(a, w, b, p, c));
    (void)E;
    return 0;
}

extern "C" int host_ctx_mean3(const void* a, int lda, const void* b, int ldb, const void* c, int ldc, int B, int H, int W, int C,
                              int h2, int w2, int h4, int w4, void* out, int ldo) {
    CtxArgs g;
    g.a = (const ym_half*)a; g.b = (const ym_half*)b; g.c = (const ym_half*)c; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.H = H; g.W = W; g.C = C; g.h2 = h2; g.w2 = w2; g.h4 = h4; g.w4 = w4;
    g.sy2 = (float)h2 / (float)H; g.sx2 = (float)w2 / (float)W; g.sy4 = (float)h4 / (float)H; g.sx4 = (float)w4 / (float)W;
    ym_half* o = (ym_half*)out;
    for (int img = 0; img < B; ++img)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                for (int ch = 0; ch < C; ++ch) o[((long long)(img * H + y) * W + x) * ldo + ch] = ym_f2h(ctx_element(g, img, y, x, ch));
    return 0;
}
