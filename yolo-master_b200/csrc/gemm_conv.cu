// Implicit-GEMM convolution / image-grouped expert GEMM for NHWC fp16 activations (sm_100a).
//
// One kernel template covers
//   * Conv (1x1 / 3x3, stride 1/2, groups=1) + folded-BN bias + SiLU + residual      [conv.py:69-89]
//   * SimpleExpert GEMM1 / GEMM2 of the routed MoE-FFN, one problem per (image, k)   [moe/experts.py:73-88]
//       - weight-pointer indirection through the router's index table (no gather copy)
//       - GroupNorm statistics accumulated in the epilogue (sum, sumsq per (problem, group))
//       - GroupNorm-normalise + SiLU applied to the A operand in registers (GEMM2 prologue)
//   * MoE combine: shared expert GEMM + sum_j w_j*GN2(o_j) + residual in one epilogue [moe/modules.py:1085-1157]
//
// Tiling: CTA = 128 threads (4 warps), BM=128 output pixels x BN output channels, BK=32,
// 3-stage cp.async pipeline, ldmatrix + mma.sync.m16n8k16 (fp32 accumulate).
// The N-scale layers this serves are HBM-bound (AI 15..260 FLOP/B < ridge), so the
// design goal is one pass over the activation with 16-byte coalesced accesses, not tensor peak.
#include "ym_common.cuh"

namespace ym {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int SK = BK + 8;  // padded smem row (halves): 80 B -> conflict-free ldmatrix
constexpr int STAGES = 3;
constexpr int NTHREADS = 128;

enum { EPI_STD = 0, EPI_STATS = 1, EPI_MOE_COMBINE = 2 };

struct GemmConvParams {
    const __half* x;   int ldx;
    const __half* w;   int Kpad;
    const float* bias;
    void* out;         int ldo;   int out_f32;
    const __half* res; int ldr;
    int B, H, W, Cin, Ho, Wo, Cout, KH, KW, stride, pad;
    int M;     // rows per problem (B*Ho*Wo for plain conv, H*W for grouped)
    int K;     // KH*KW*Cin
    int act;   // 0 none, 1 SiLU
    // grouped (problem = blockIdx.z)
    const int* route_idx;      // expert per problem (null -> plain); a negative entry = dropped route (CTA exits)
    long long w_expert_stride; // elements between experts' packed weights
    int a_div;                 // A problem = p / a_div
    int bias_expert_stride;    // >0: bias + expert * stride (per-expert folded-BN bias, ES_MOE)
    const float* prob_scale;   // optional per-problem output scale applied after the activation (routing weight)
    // A-operand GroupNorm+SiLU prologue: per (problem, k) scale/shift
    const float* a_scale; const float* a_shift;
    // stats epilogue
    float* stats; int groups;  // partial sums [P][m tiles][Cout/8][2]
    // combine epilogue
    const __half* o; int ldo_o; const float* o_scale; const float* o_shift; int topk; int HW;
};

template <int BN, int EPI, bool A_XFORM>
// (A register cap of 96 for the BN = 64 tiles - 5 instead of 4 CTAs per SM at 16-88 bytes of spill - was measured: the whole
// step got 1 % slower, so the kernel keeps the compiler's 124-128 registers.)
__global__ void __launch_bounds__(NTHREADS) gemm_conv_kernel(const GemmConvParams p) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sA = reinterpret_cast<__half*>(smem_raw);                 // [STAGES][BM][SK]
    __half* sB = sA + STAGES * BM * SK;                               // [STAGES][BN][SK]
    float* sXf = reinterpret_cast<float*>(sB + STAGES * BN * SK);     // A_XFORM: scale[K], shift[K]; STATS: [BN/8][2]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int prob = blockIdx.z;

    const __half* xb = p.x;
    const __half* wb = p.w;
    char* outb = reinterpret_cast<char*>(p.out);
    const float* biasb = p.bias;
    float pscale = 1.f;
    if (p.route_idx != nullptr) {
        const int e = p.route_idx[prob];
        if (e < 0) return;                     // dropped route: nothing to compute, the combine kernel skips the slot
        if (p.bias_expert_stride > 0 && biasb != nullptr) biasb += (long long)e * p.bias_expert_stride;
        if (p.prob_scale != nullptr) pscale = p.prob_scale[prob];
        wb += (long long)e * p.w_expert_stride;
        xb += (long long)(prob / p.a_div) * p.M * p.ldx;
        outb += (long long)prob * p.M * p.ldo * (p.out_f32 ? 4 : 2);
    }

    if (A_XFORM) {
        const float* sc = p.a_scale + (long long)prob * p.K;
        const float* sh = p.a_shift + (long long)prob * p.K;
        for (int i = tid; i < p.K; i += NTHREADS) {
            sXf[i] = sc[i];
            sXf[p.K + i] = sh[i];
        }
    }

    // ---- per-thread A-row bookkeeping: thread owns chunk c=tid%4 of rows tid/4 + 32*i
    const int a_chunk = tid & 3;
    int a_iy0[4], a_ix0[4];
    long long a_base[4];
    bool a_rowok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = (tid >> 2) + 32 * i;
        const int m = m0 + r;
        a_rowok[i] = m < p.M;
        const int mm = a_rowok[i] ? m : 0;
        const int hw = p.Ho * p.Wo;
        const int b = mm / hw;
        const int rem = mm - b * hw;
        const int oy = rem / p.Wo;
        const int ox = rem - oy * p.Wo;
        a_iy0[i] = oy * p.stride - p.pad;
        a_ix0[i] = ox * p.stride - p.pad;
        a_base[i] = (long long)b * p.H * p.W;
    }
    // B rows: thread owns chunk tid%4 of rows tid/4 + 32*i (i < BN/32, at least 1)
    constexpr int B_ITERS = (BN * 4 + NTHREADS - 1) / NTHREADS;

    const int KT = (p.K + BK - 1) / BK;

    auto load_stage = [&](int stage, int kt) {
        const int k = kt * BK + a_chunk * 8;
        const bool kok = k < p.K;
        int ky = 0, kx = 0, ci = k;
        if (p.KH * p.KW > 1) {
            const int tap = k / p.Cin;
            ci = k - tap * p.Cin;
            ky = tap / p.KW;
            kx = tap - ky * p.KW;
        }
        __half* dA = sA + stage * BM * SK;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = (tid >> 2) + 32 * i;
            const int iy = a_iy0[i] + ky, ix = a_ix0[i] + kx;
            const bool ok = kok && a_rowok[i] && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
            const __half* src = ok ? xb + (a_base[i] + (long long)iy * p.W + ix) * p.ldx + ci : xb;
            cp_async16(dA + r * SK + a_chunk * 8, src, ok ? 16 : 0);
        }
        __half* dB = sB + stage * BN * SK;
#pragma unroll
        for (int i = 0; i < B_ITERS; ++i) {
            const int idx = tid + i * NTHREADS;
            const int r = idx >> 2, c = idx & 3;
            if (r < BN) {
                const int n = n0 + r;
                const bool ok = n < p.Cout;
                const __half* src = ok ? wb + (long long)n * p.Kpad + kt * BK + c * 8 : wb;
                cp_async16(dB + r * SK + c * 8, src, ok ? 16 : 0);
            }
        }
    };

    constexpr int NT = BN / 8;  // n8 tiles per warp (warp covers 32 rows x BN cols)
    float acc[2][NT][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[mi][ni][q] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < KT) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const __half* tA = sA + (kt % STAGES) * BM * SK + warp * 32 * SK;
        const __half* tB = sB + (kt % STAGES) * BN * SK;
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
            uint32_t af[2][4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const __half* pa = tA + (mi * 16 + (lane & 15)) * SK + ks * 16 + (lane >> 4) * 8;
                ldmatrix_x4(af[mi][0], af[mi][1], af[mi][2], af[mi][3], pa);
            }
            if (A_XFORM) {
                // a0:(row g, k 2t..2t+1) a1:(row g+8, same k) a2:(row g, k+8) a3:(row g+8, k+8)
                const int kb = kt * BK + ks * 16 + 2 * (lane & 3);
                const float s0 = sXf[kb], s1 = sXf[kb + 1], s8 = sXf[kb + 8], s9 = sXf[kb + 9];
                const float h0 = sXf[p.K + kb], h1 = sXf[p.K + kb + 1], h8 = sXf[p.K + kb + 8], h9 = sXf[p.K + kb + 9];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float2 v = unpack_half2(af[mi][q]);
                        const bool hi = q >= 2;
                        v.x = silu_f(fmaf(v.x, hi ? s8 : s0, hi ? h8 : h0));
                        v.y = silu_f(fmaf(v.y, hi ? s9 : s1, hi ? h9 : h1));
                        af[mi][q] = pack_half2(v.x, v.y);
                    }
                }
            }
            if (NT >= 2) {
#pragma unroll
                for (int nj = 0; nj < NT / 2; ++nj) {
                    uint32_t b0, b1, b2, b3;
                    const __half* pb = tB + (nj * 16 + (lane & 7) + (lane >> 4) * 8) * SK + ks * 16 + ((lane >> 3) & 1) * 8;
                    ldmatrix_x4(b0, b1, b2, b3, pb);
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi) {
                        mma_16816(acc[mi][2 * nj], af[mi], b0, b1);
                        mma_16816(acc[mi][2 * nj + 1], af[mi], b2, b3);
                    }
                }
            } else {
                uint32_t b0, b1;
                const __half* pb = tB + (lane & 7) * SK + ks * 16 + ((lane >> 3) & 1) * 8;
                ldmatrix_x2(b0, b1, pb);
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) mma_16816(acc[mi][0], af[mi], b0, b1);
            }
        }
    }
    cp_async_wait<0>();

    // ---------------- epilogue
    // Phase 1: accumulators (+bias, +activation) -> warp-private fp32 staging tile in shared memory (reusing the
    // pipeline buffers); GroupNorm statistics are taken here from the values as they will be stored.
    // Phase 2: each lane owns one 16-byte (8-channel) chunk of an output row: residual / expert outputs are read and
    // the result written with full 128-byte-line coalescing.
    __syncthreads();  // every warp is done reading sA/sB
    constexpr int SP = BN + 8;  // staging pitch (floats): == 8 mod 32 -> conflict-free float2 stores
    float* stg = reinterpret_cast<float*>(smem_raw) + warp * 32 * SP;
    const int g = lane >> 2, t = lane & 3;
    float* sS = sXf + (A_XFORM ? 2 * p.K : 0);
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) {
        const int nl = ni * 8 + 2 * t;
        const int n = n0 + nl;
        const bool nok = n < p.Cout;  // an odd Cout (1-class heads, the OBB angle) leaves the second channel of the last pair out
        float bias0 = 0.f, bias1 = 0.f;
        if (nok && biasb != nullptr) {
            bias0 = biasb[n];
            if (n + 1 < p.Cout) bias1 = biasb[n + 1];
        }
        float ssum = 0.f, ssq = 0.f;
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int rl = mi * 16 + g + hf * 8;
                const int m = m0 + warp * 32 + rl;
                float v0 = acc[mi][ni][hf * 2 + 0] + bias0;
                float v1 = acc[mi][ni][hf * 2 + 1] + bias1;
                if (EPI == EPI_STATS && m < p.M && nok) {
                    const float2 r = __half22float2(__floats2half2_rn(v0, v1));
                    ssum += r.x + r.y;
                    ssq += r.x * r.x + r.y * r.y;
                }
                if (p.act == 1) {
                    v0 = silu_f(v0);
                    v1 = silu_f(v1);
                }
                v0 *= pscale;
                v1 *= pscale;
                *reinterpret_cast<float2*>(stg + rl * SP + nl) = make_float2(v0, v1);
            }
        }
        if (EPI == EPI_STATS) {
            ssum = warp_sum(ssum);
            ssq = warp_sum(ssq);
            if (lane == 0) {  // one slot per (warp, n8 tile): no atomics -> bit-reproducible statistics
                sS[(warp * NT + ni) * 2 + 0] = ssum;
                sS[(warp * NT + ni) * 2 + 1] = ssq;
            }
        }
    }
    __syncwarp();
    {
        constexpr int CH = BN / 8;          // 16-byte chunks per row
        constexpr int RPI = 32 / CH;        // rows per iteration (BN=64: 4, 32: 8, 16: 16, 8: 32)
        const int ch = lane % CH, rsub = lane / CH;
        const int n = n0 + ch * 8;
#pragma unroll
        for (int it = 0; it < 32 / RPI; ++it) {
            const int rl = it * RPI + rsub;
            const int m = m0 + warp * 32 + rl;
            if (m >= p.M || n >= p.Cout) continue;
            float v[8];
            {
                const float4 a0 = *reinterpret_cast<const float4*>(stg + rl * SP + ch * 8);
                const float4 a1 = *reinterpret_cast<const float4*>(stg + rl * SP + ch * 8 + 4);
                v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
            }
            const bool full = n + 8 <= p.Cout;  // Cout % 8 != 0 only for tiny heads (e.g. 4 box channels)
            if (EPI == EPI_MOE_COMBINE) {
                const int bimg = m / p.HW;
                const int r = m - bimg * p.HW;
                for (int j = 0; j < p.topk; ++j) {
                    const long long pr = (long long)bimg * p.topk + j;
                    const Half8 ov = *reinterpret_cast<const Half8*>(p.o + (pr * p.HW + r) * p.ldo_o + n);
                    const float4 s0 = *reinterpret_cast<const float4*>(p.o_scale + pr * p.Cout + n);
                    const float4 s1 = *reinterpret_cast<const float4*>(p.o_scale + pr * p.Cout + n + 4);
                    const float4 h0 = *reinterpret_cast<const float4*>(p.o_shift + pr * p.Cout + n);
                    const float4 h1 = *reinterpret_cast<const float4*>(p.o_shift + pr * p.Cout + n + 4);
                    const float2 o0 = __half22float2(ov.v[0]), o1 = __half22float2(ov.v[1]);
                    const float2 o2 = __half22float2(ov.v[2]), o3 = __half22float2(ov.v[3]);
                    v[0] += fmaf(o0.x, s0.x, h0.x); v[1] += fmaf(o0.y, s0.y, h0.y);
                    v[2] += fmaf(o1.x, s0.z, h0.z); v[3] += fmaf(o1.y, s0.w, h0.w);
                    v[4] += fmaf(o2.x, s1.x, h1.x); v[5] += fmaf(o2.y, s1.y, h1.y);
                    v[6] += fmaf(o3.x, s1.z, h1.z); v[7] += fmaf(o3.y, s1.w, h1.w);
                }
            }
            if (full) {
                if (p.res != nullptr) {
                    const Half8 rv = *reinterpret_cast<const Half8*>(p.res + (long long)m * p.ldr + n);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 rf = __half22float2(rv.v[q]);
                        v[2 * q] += rf.x;
                        v[2 * q + 1] += rf.y;
                    }
                }
                if (p.out_f32) {
                    float4* dst = reinterpret_cast<float4*>(outb + ((long long)m * p.ldo + n) * 4);
                    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
                } else {
                    Half8 o;
#pragma unroll
                    for (int q = 0; q < 4; ++q) o.v[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
                    *reinterpret_cast<Half8*>(outb + ((long long)m * p.ldo + n) * 2) = o;
                }
            } else {  // ragged channel tail: element-wise
                for (int q = 0; q < 8 && n + q < p.Cout; ++q) {
                    float r = v[q];
                    if (p.res != nullptr) r += __half2float(p.res[(long long)m * p.ldr + n + q]);
                    if (p.out_f32) reinterpret_cast<float*>(outb)[(long long)m * p.ldo + n + q] = r;
                    else reinterpret_cast<__half*>(outb)[(long long)m * p.ldo + n + q] = __float2half_rn(r);
                }
            }
        }
    }
    if (EPI == EPI_STATS) {
        __syncthreads();
        if (tid < NT) {
            const int n = n0 + tid * 8;
            if (n < p.Cout) {
                float a = 0.f, q = 0.f;
#pragma unroll
                for (int w = 0; w < NTHREADS / 32; ++w) {
                    a += sS[(w * NT + tid) * 2 + 0];
                    q += sS[(w * NT + tid) * 2 + 1];
                }
                // partial statistics: [problem][m tile][n8 tile][2], reduced in fixed order by gn_finalize
                float* dst = p.stats + (((long long)prob * gridDim.x + blockIdx.x) * (p.Cout / 8) + (n >> 3)) * 2;
                dst[0] = a;
                dst[1] = q;
            }
        }
    }
}

template <int BN, int EPI, bool A_XFORM>
static int launch_gemm(const GemmConvParams& p, int problems, cudaStream_t stream) {
    size_t smem = (size_t)STAGES * (BM + BN) * SK * sizeof(__half);
    static_assert((size_t)STAGES * (BM + BN) * SK * sizeof(__half) >= (size_t)BM * (BN + 8) * sizeof(float),
                  "epilogue staging tile must fit in the pipeline buffers");
    if (A_XFORM) smem += 2 * (size_t)p.K * sizeof(float);
    if (EPI == EPI_STATS) smem += (NTHREADS / 32) * (BN / 8) * 2 * sizeof(float);
    auto kern = gemm_conv_kernel<BN, EPI, A_XFORM>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            ym_set_error("gemm_conv: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
            return YM_ERR_CUDA;
        }
    }
    dim3 grid((p.M + BM - 1) / BM, (p.Cout + BN - 1) / BN, problems);
    launch_pdl(kern, grid, NTHREADS, smem, stream, p);
    YM_CHECK_LAUNCH("gemm_conv");
    return YM_OK;
}

template <int EPI, bool A_XFORM>
static int dispatch_bn(const GemmConvParams& p, int problems, cudaStream_t stream) {
    if (p.Cout <= 8) return launch_gemm<8, EPI, A_XFORM>(p, problems, stream);
    if (p.Cout <= 16) return launch_gemm<16, EPI, A_XFORM>(p, problems, stream);
    if (p.Cout <= 32 || p.Cout % 64 == 32) return launch_gemm<32, EPI, A_XFORM>(p, problems, stream);
    return launch_gemm<64, EPI, A_XFORM>(p, problems, stream);
}

// small_conv.cu: patch-staged kernel for 3x3 layers over 8 / 16 input channels
int small_conv3_supported(int Cin, int Cout, int KH, int KW, int stride, int pad, int Kpad, int out_f32, int B);
int small_conv3_run(const void* x, int ldx, int B, int H, int W, int Cin, const void* w, int Kpad, const float* bias, int Cout, int stride,
                    void* out, int ldo, const void* res, int ldr, int act, cudaStream_t st);

}  // namespace ym

using namespace ym;

extern "C" int ym_conv2d_nhwc(const void* x, int ldx, int B, int H, int W, int Cin, const void* w, int Kpad,
                              const float* bias, int Cout, int KH, int KW, int stride, int pad, void* out, int ldo,
                              int out_f32, const void* res, int ldr, int act, void* stream) {
    YM_CHECK_ARG(x && w && out, "ym_conv2d_nhwc: null pointer");
    YM_CHECK_ARG(Cin % 8 == 0 && ldx % 8 == 0, "ym_conv2d_nhwc: Cin (%d) and ldx (%d) must be multiples of 8", Cin, ldx);
    YM_CHECK_ARG(Cout % 2 == 0, "ym_conv2d_nhwc: Cout (%d) must be even (the host side pads odd widths to a multiple of 8)", Cout);
    YM_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0, "ym_conv2d_nhwc: x/w must be 16-byte aligned");
    YM_CHECK_ARG(Cout % 8 != 0 || (((uintptr_t)out & 15) == 0 && ldo % (out_f32 ? 4 : 8) == 0),
                 "ym_conv2d_nhwc: out must be 16-byte aligned with a pitch (%d) that keeps rows 16-byte aligned", ldo);
    YM_CHECK_ARG(Kpad % BK == 0 && Kpad >= KH * KW * Cin, "ym_conv2d_nhwc: bad Kpad %d", Kpad);
    YM_CHECK_ARG(stride >= 1 && KH >= 1 && KW >= 1 && pad >= 0, "ym_conv2d_nhwc: bad geometry");
    YM_CHECK_ARG(res == nullptr || (ldr % 8 == 0 && ((uintptr_t)res & 15) == 0), "ym_conv2d_nhwc: residual must be 16-byte aligned");
    if (B == 0) return YM_OK;
    if (small_conv3_supported(Cin, Cout, KH, KW, stride, pad, Kpad, out_f32, B)) {
        const int rc = small_conv3_run(x, ldx, B, H, W, Cin, w, Kpad, bias, Cout, stride, out, ldo, res, ldr, act, (cudaStream_t)stream);
        if (rc) return rc;
        YM_CHECK_LAUNCH("small_conv3");
        return YM_OK;
    }
    GemmConvParams p;
    memset(&p, 0, sizeof(p));
    p.x = (const __half*)x; p.ldx = ldx; p.w = (const __half*)w; p.Kpad = Kpad; p.bias = bias;
    p.out = out; p.ldo = ldo; p.out_f32 = out_f32; p.res = (const __half*)res; p.ldr = ldr;
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad;
    p.Ho = (H + 2 * pad - KH) / stride + 1;
    p.Wo = (W + 2 * pad - KW) / stride + 1;
    p.Cout = Cout; p.M = B * p.Ho * p.Wo; p.K = KH * KW * Cin; p.act = act; p.a_div = 1;
    YM_CHECK_ARG(p.Ho > 0 && p.Wo > 0, "ym_conv2d_nhwc: empty output");
    return dispatch_bn<EPI_STD, false>(p, 1, (cudaStream_t)stream);
}

// Grouped 1x1 "image GEMM" for the routed experts.  problem p in [0,P): expert = route_idx[p],
// A = a + (p / a_div) * HW * lda  (HW rows, K cols),  out = out + p * HW * ldo (HW rows, N cols).
// Optional A prologue SiLU(a*scale[p][k] + shift[p][k]); optional GroupNorm stats epilogue.
extern "C" int ym_moe_expert_gemm(const void* a, int lda, int a_div, int P, int HW, int K, const void* w, int Kpad,
                                  long long w_expert_stride, const int* route_idx, int N, void* out, int ldo,
                                  const float* a_scale, const float* a_shift, float* stats, int groups, void* stream) {
    YM_CHECK_ARG(a && w && out && route_idx, "ym_moe_expert_gemm: null pointer");
    YM_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && N % 8 == 0 && ldo % 8 == 0, "ym_moe_expert_gemm: bad dims K=%d N=%d", K, N);
    YM_CHECK_ARG((((uintptr_t)a | (uintptr_t)out) & 15) == 0, "ym_moe_expert_gemm: 16-byte alignment");
    YM_CHECK_ARG(Kpad % BK == 0 && Kpad >= K, "ym_moe_expert_gemm: bad Kpad");
    YM_CHECK_ARG((a_scale == nullptr) == (a_shift == nullptr), "ym_moe_expert_gemm: scale/shift mismatch");
    YM_CHECK_ARG(a_scale == nullptr || K % BK == 0, "ym_moe_expert_gemm: A prologue needs K %% 32 == 0");
    YM_CHECK_ARG(stats == nullptr || (groups > 0 && N % groups == 0 && (N / groups) % 8 == 0),
                 "ym_moe_expert_gemm: GroupNorm group width must be a multiple of 8 (N=%d groups=%d)", N, groups);
    YM_CHECK_ARG(a_div >= 1, "ym_moe_expert_gemm: a_div");
    if (P == 0) return YM_OK;
    GemmConvParams p;
    memset(&p, 0, sizeof(p));
    p.x = (const __half*)a; p.ldx = lda; p.w = (const __half*)w; p.Kpad = Kpad;
    p.out = out; p.ldo = ldo;
    p.B = 1; p.H = HW; p.W = 1; p.Cin = K; p.KH = 1; p.KW = 1; p.stride = 1; p.pad = 0; p.Ho = HW; p.Wo = 1;
    p.Cout = N; p.M = HW; p.K = K; p.act = 0;
    p.route_idx = route_idx; p.w_expert_stride = w_expert_stride; p.a_div = a_div;
    p.a_scale = a_scale; p.a_shift = a_shift; p.stats = stats; p.groups = groups;
    cudaStream_t st = (cudaStream_t)stream;
    if (stats != nullptr) {
        if (a_scale) return dispatch_bn<EPI_STATS, true>(p, P, st);
        return dispatch_bn<EPI_STATS, false>(p, P, st);
    }
    if (a_scale) return dispatch_bn<EPI_STD, true>(p, P, st);
    return dispatch_bn<EPI_STD, false>(p, P, st);
}

// out = x + SiLU(x*Ws^T + bs) + sum_j (o[b*topk+j] * o_scale[b*topk+j] + o_shift[b*topk+j])
extern "C" int ym_moe_combine(const void* x, int ldx, int B, int HW, int C, const void* ws, int Kpad, const float* bias_s,
                              const void* o, int ldo_o, const float* o_scale, const float* o_shift, int topk, void* out,
                              int ldo, int add_residual, void* stream) {
    YM_CHECK_ARG(x && ws && o && o_scale && o_shift && out, "ym_moe_combine: null pointer");
    YM_CHECK_ARG(C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && ldo_o % 8 == 0, "ym_moe_combine: bad dims");
    YM_CHECK_ARG((((uintptr_t)x | (uintptr_t)o | (uintptr_t)out | (uintptr_t)o_scale | (uintptr_t)o_shift) & 15) == 0,
                 "ym_moe_combine: 16-byte alignment");
    YM_CHECK_ARG(Kpad % BK == 0 && Kpad >= C, "ym_moe_combine: bad Kpad");
    if (B == 0) return YM_OK;
    GemmConvParams p;
    memset(&p, 0, sizeof(p));
    p.x = (const __half*)x; p.ldx = ldx; p.w = (const __half*)ws; p.Kpad = Kpad; p.bias = bias_s;
    p.out = out; p.ldo = ldo;
    if (add_residual) { p.res = (const __half*)x; p.ldr = ldx; }
    p.B = B; p.H = HW; p.W = 1; p.Cin = C; p.KH = 1; p.KW = 1; p.stride = 1; p.pad = 0; p.Ho = HW; p.Wo = 1;
    p.Cout = C; p.M = B * HW; p.K = C; p.act = 1; p.a_div = 1;
    p.o = (const __half*)o; p.ldo_o = ldo_o; p.o_scale = o_scale; p.o_shift = o_shift; p.topk = topk; p.HW = HW;
    return dispatch_bn<EPI_MOE_COMBINE, false>(p, 1, (cudaStream_t)stream);
}

// GroupNorm partial statistics [P][mtiles][C/8][2] -> per-(problem, channel) scale/shift.  One warp per (problem, group):
// lanes stride over the partials, fixed-shape shuffle tree (deterministic), then write the group's channels.
// scale = rw*rstd*gamma, shift = rw*(beta - mean*rstd*gamma)
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float* __restrict__ stats, int P, int mtiles, int groups, int C,
                                                          float count, float eps, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const int* __restrict__ route_idx,
                                                          const float* __restrict__ route_w, float* __restrict__ scale,
                                                          float* __restrict__ shift) {
    pdl_prologue();
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= P * groups) return;
    const int pr = wid / groups, grp = wid - pr * groups;
    const int cpg = C / groups, tpg = cpg / 8, nt = C / 8;
    float s = 0.f, q = 0.f;
    for (int i = lane; i < mtiles * tpg; i += 32) {
        const int mt = i / tpg, t = grp * tpg + i % tpg;
        const float* src = stats + (((long long)pr * mtiles + mt) * nt + t) * 2;
        s += src[0];
        q += src[1];
    }
    s = ym::warp_sum(s);
    q = ym::warp_sum(q);
    const float mean = s / count;
    const float var = fmaxf(q / count - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    const int e = route_idx[pr];
    const float rw = route_w ? route_w[pr] : 1.f;
    for (int j = lane; j < cpg; j += 32) {
        const int c = grp * cpg + j;
        const float gm = gamma[e * C + c], bt = beta[e * C + c];
        scale[(long long)pr * C + c] = rw * rstd * gm;
        shift[(long long)pr * C + c] = rw * (bt - mean * rstd * gm);
    }
}

extern "C" long long ym_moe_stats_floats(int P, int HW, int N) { return (long long)P * ((HW + BM - 1) / BM) * (N / 8) * 2; }

extern "C" int ym_gn_finalize(const float* stats, int P, int HW, int groups, int C, float count, float eps, const float* gamma,
                              const float* beta, const int* route_idx, const float* route_w, float* scale, float* shift,
                              void* stream) {
    YM_CHECK_ARG(stats && gamma && beta && route_idx && scale && shift, "ym_gn_finalize: null pointer");
    YM_CHECK_ARG(groups > 0 && C % groups == 0 && (C / groups) % 8 == 0, "ym_gn_finalize: bad groups");
    if (P == 0) return YM_OK;
    const int n = P * groups * 32;
    launch_pdl(gn_finalize_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, stats, P, (HW + BM - 1) / BM, groups, C, count, eps, gamma, beta,
                                                                          route_idx, route_w, scale, shift);
    YM_CHECK_LAUNCH("gn_finalize");
    return YM_OK;
}

// Same with an explicit number of partial-sum tiles per problem (ym_moe_ffn writes one set per STRIP of row tiles, tc_moe.cu).
extern "C" int ym_gn_finalize_tiles(const float* stats, int P, int tiles, int groups, int C, float count, float eps, const float* gamma,
                                    const float* beta, const int* route_idx, const float* route_w, float* scale, float* shift,
                                    void* stream) {
    YM_CHECK_ARG(stats && gamma && beta && route_idx && scale && shift, "ym_gn_finalize_tiles: null pointer");
    YM_CHECK_ARG(groups > 0 && C % groups == 0 && (C / groups) % 8 == 0 && tiles >= 1, "ym_gn_finalize_tiles: bad groups / tiles");
    if (P == 0) return YM_OK;
    const int n = P * groups * 32;
    launch_pdl(gn_finalize_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, stats, P, tiles, groups, C, count, eps, gamma, beta,
               route_idx, route_w, scale, shift);
    YM_CHECK_LAUNCH("gn_finalize_tiles");
    return YM_OK;
}

// ES_MOE expert pointwise stage: y[p] = route_w[p] * SiLU(t[p] * Wpw[e_p]^T + bias[e_p])  (experts.py:280-296 with BN folded),
// one problem per (image, k); problems whose route_idx is negative (dropped by the dynamic threshold) are skipped.
extern "C" int ym_esmoe_pointwise(const void* t, int ldt, int P, int HW, int K, const void* w, int Kpad, long long w_expert_stride,
                                  const float* bias_all, int N, const int* route_idx, const float* route_w, void* y, int ldy,
                                  void* stream) {
    YM_CHECK_ARG(t && w && bias_all && route_idx && route_w && y, "ym_esmoe_pointwise: null pointer");
    YM_CHECK_ARG(K % 8 == 0 && ldt % 8 == 0 && N % 8 == 0 && ldy % 8 == 0, "ym_esmoe_pointwise: bad dims K=%d N=%d", K, N);
    YM_CHECK_ARG(Kpad % BK == 0 && Kpad >= K, "ym_esmoe_pointwise: bad Kpad");
    YM_CHECK_ARG((((uintptr_t)t | (uintptr_t)y | (uintptr_t)w) & 15) == 0, "ym_esmoe_pointwise: 16-byte alignment");
    if (P == 0) return YM_OK;
    GemmConvParams p;
    memset(&p, 0, sizeof(p));
    p.x = (const __half*)t; p.ldx = ldt; p.w = (const __half*)w; p.Kpad = Kpad; p.bias = bias_all; p.bias_expert_stride = N;
    p.out = y; p.ldo = ldy;
    p.B = 1; p.H = HW; p.W = 1; p.Cin = K; p.KH = 1; p.KW = 1; p.stride = 1; p.pad = 0; p.Ho = HW; p.Wo = 1;
    p.Cout = N; p.M = HW; p.K = K; p.act = 1;
    p.route_idx = route_idx; p.w_expert_stride = w_expert_stride; p.a_div = 1; p.prob_scale = route_w;
    return dispatch_bn<EPI_STD, false>(p, P, (cudaStream_t)stream);
}
