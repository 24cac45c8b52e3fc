// 3x3 convolution over 8 or 16 input channels (stride 1 / 2, pad 1) + folded-BN bias + SiLU + residual, NHWC fp16 (sm_100a).
//
// The layers this serves - the P2 stage of the backbone (16->32 s2 @320^2, the 16->8->16 bottleneck @160^2) and the c = 16 bottlenecks
// and box towers further down (conv.py:69-89 behind block.py's Bottleneck / C3k2, head.py's cv2) - move 60-160 MB each at bs32 and do
// almost no arithmetic (K = 72 / 144).  As implicit GEMMs they ran at 1.3-1.5 TB/s (profiles/r02_launch_roofline.txt, r02g: 84 / 61 /
// 63 us): the generic kernel gathers every one of the nine taps of every pixel from L2 as 16-byte cp.async requests (9x the input
// in load instructions), and the TMA kernel is worse still at 32-byte rows (ops.py, `small_k`).
//
// Here a CTA stages the input patch of its 8 x 32 output tile in shared memory ONCE (coalesced 16-byte cp.async, zero-filled
// borders) together with the whole weight matrix, and builds every A fragment of mma.sync.m16n8k16 from the patch with ldmatrix:
// the K axis is cut into slices of 8 channels, slice = tap * (Cin / 8) + c8, and one k-step takes two slices - the four 8x8 matrices
// of ldmatrix.x4 take independent row addresses, so the two halves of a fragment may come from different taps (Cin = 8) or from the
// two channel halves of one tap (Cin = 16).  Pixels are 48 bytes apart in the patch (Cin = 16; 16 bytes at Cin = 8): eight
// consecutive pixels fall in eight different 16-byte bank groups.  For stride 2 the even and odd patch columns live in separate planes,
// so the output pixels of a fragment are adjacent in shared memory for every tap.
// Weights are read in the layout ym_conv2d_nhwc already takes ([Cout][Kpad] fp16, k = tap * Cin + ci): the entry point does not change.
// warp = one output row of the tile = two 16-pixel fragments; B fragments are read once per k-step for both.
#include "ym_common.cuh"

namespace ym {

struct SmallConvParams {
    const __half* __restrict__ x;   int ldx;
    const __half* __restrict__ w;   int Kpad;
    const float* __restrict__ bias;
    __half* __restrict__ out;       int ldo;
    const __half* __restrict__ res; int ldr;
    int B, H, W, Ho, Wo, tiles_x, act;
};

template <int CIN, int S, int COUT>
struct SmallConvCfg {
    static constexpr int TW = 32, TH = 8;
    static constexpr int IH = (TH - 1) * S + 3, IW = (TW - 1) * S + 3;
    static constexpr int PW = (IW + S - 1) / S;               // pixels per plane row
    static constexpr int PP = CIN == 16 ? 48 : 16;            // bytes between neighbouring pixels of a plane row
    static constexpr int ROWB = PW * PP;
    static constexpr int PATCH_BYTES = IH * S * ROWB;
    static constexpr int CH = CIN / 8;                        // 16-byte chunks per pixel
    static constexpr int SLICES = 9 * CH, KSTEPS = (SLICES + 1) / 2;
    static constexpr int NT = COUT / 8;
    static constexpr int SP = COUT + 8;                       // staging pitch in halves: 16-byte rows stay aligned, (SP / 2) % 8 == 4
    static constexpr int STG_BYTES = 8 * 16 * SP * 2;
};

// byte offset of (patch row ry, patch column rx, chunk c8)
template <typename Cfg, int S>
__device__ __forceinline__ int patch_off(int ry, int rx, int c8) {
    return ((ry * S + (rx % S)) * Cfg::PW + rx / S) * Cfg::PP + c8 * 16;
}

template <int CIN, int S, int COUT>
__global__ void __launch_bounds__(256) small_conv3_kernel(const SmallConvParams p) {
    using Cfg = SmallConvCfg<CIN, S, COUT>;
    extern __shared__ __align__(128) unsigned char sc_smem[];
    unsigned char* patch = sc_smem;
    const int wpitch = (p.Kpad + 8) * 2;                      // bytes per weight row: (Kpad + 8) / 2 words == 4 mod 8 for Kpad = 96 / 160
    unsigned char* sW = patch + Cfg::PATCH_BYTES;
    __half* stg_all = reinterpret_cast<__half*>(sW + ((COUT * wpitch + 127) & ~127));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int oy0 = (blockIdx.x / p.tiles_x) * Cfg::TH, ox0 = (blockIdx.x % p.tiles_x) * Cfg::TW;
    const int iy0 = oy0 * S - 1, ix0 = ox0 * S - 1;

    // weights do not depend on the producer grid: stage them before the dependency wait
    {
        const int chunks_per_row = p.Kpad / 8;
        for (int i = tid; i < COUT * chunks_per_row; i += 256) {
            const int n = i / chunks_per_row, c = i - n * chunks_per_row;
            cp_async16(sW + n * wpitch + c * 16, p.w + (long long)n * p.Kpad + c * 8, 16);
        }
    }
    pdl_prologue();
    {
        // chunk i = (row ry, column-chunk rem) with rem = rx * CH + c8: each thread walks i = tid, tid + 256, ... by carrying (ry, rem)
        constexpr int RC = Cfg::IW * Cfg::CH, DQ = 256 / RC, DR = 256 % RC;
        const __half* xb = p.x + (long long)b * p.H * p.W * p.ldx;
        int ry = tid / RC, rem = tid - ry * RC;
        while (ry < Cfg::IH) {
            const int rx = rem / Cfg::CH, c8 = rem - rx * Cfg::CH;
            const int gy = iy0 + ry, gx = ix0 + rx;
            const bool ok = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
            const __half* src = ok ? xb + ((long long)gy * p.W + gx) * p.ldx + c8 * 8 : p.x;
            cp_async16(patch + patch_off<Cfg, S>(ry, rx, c8), src, ok ? 16 : 0);
            ry += DQ;
            rem += DR;
            if (rem >= RC) { rem -= RC; ++ry; }
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    const int oy = oy0 + warp;
    if (oy >= p.Ho) return;                                    // warp-uniform; no CTA-wide barrier follows
    const int g = lane >> 2, t = lane & 3;
    // ldmatrix row address of this lane: matrix mi = lane / 8 -> fragment rows (mi & 1) * 8 + lane % 8, k half mi >> 1
    const int mi = lane >> 3, frow = (mi & 1) * 8 + (lane & 7), khalf = mi >> 1;
    const uint32_t a_lane = smem_u32(patch) + (uint32_t)((warp * S * S * Cfg::PW + frow) * Cfg::PP);
    const uint32_t w_lane = smem_u32(sW) + (uint32_t)(g * wpitch + t * 4);

    float acc[2][Cfg::NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < Cfg::NT; ++nt)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[mt][nt][q] = 0.f;

#pragma unroll
    for (int j = 0; j < Cfg::KSTEPS; ++j) {
        // slices 2j (k 0..7 of the step) and 2j+1 (k 8..15); a slice past the last one (Cin = 8: the tenth) multiplies zero weights and
        // re-reads slice 0 so that the operand is finite
        const int s0 = 2 * j, s1 = (2 * j + 1 < Cfg::SLICES) ? 2 * j + 1 : 0;
        const int tap0 = s0 / Cfg::CH, c80 = s0 % Cfg::CH, tap1 = s1 / Cfg::CH, c81 = s1 % Cfg::CH;
        const int off0 = patch_off<Cfg, S>(tap0 / 3, tap0 % 3, c80), off1 = patch_off<Cfg, S>(tap1 / 3, tap1 % 3, c81);
        const uint32_t a_addr = a_lane + (uint32_t)(khalf ? off1 : off0);
        uint32_t a[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                         : "=r"(a[mt][0]), "=r"(a[mt][1]), "=r"(a[mt][2]), "=r"(a[mt][3])
                         : "r"(a_addr + (uint32_t)(mt * 16 * Cfg::PP)));
#pragma unroll
        for (int nt = 0; nt < Cfg::NT; ++nt) {
            uint32_t b0, b1;
            const uint32_t wa = w_lane + (uint32_t)(nt * 8 * wpitch + j * 32);
            asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(b0) : "r"(wa));
            asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(b1) : "r"(wa + 16));
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma_16816(acc[mt][nt], a[mt], b0, b1);
        }
    }

    // ---- epilogue: bias, SiLU and the residual in fragment layout, rounded ONCE to fp16 into a warp-private staging tile, then one
    // 16-byte (8-channel) chunk of an output pixel per lane to global memory
    __half* stg = stg_all + warp * 16 * Cfg::SP;
    constexpr int RPI = 32 / Cfg::NT;                          // pixels per pass (Cout 32: 8, 16: 16, 8: 32)
    const int ch = lane % Cfg::NT, rsub = lane / Cfg::NT;
    float bias2[Cfg::NT][2];
#pragma unroll
    for (int nt = 0; nt < Cfg::NT; ++nt) {
        bias2[nt][0] = p.bias != nullptr ? p.bias[nt * 8 + 2 * t] : 0.f;
        bias2[nt][1] = p.bias != nullptr ? p.bias[nt * 8 + 2 * t + 1] : 0.f;
    }
    const long long row_pix = ((long long)b * p.Ho + oy) * p.Wo;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int r = g + hf * 8, ox = ox0 + mt * 16 + r;
#pragma unroll
            for (int nt = 0; nt < Cfg::NT; ++nt) {
                float v0 = acc[mt][nt][2 * hf] + bias2[nt][0], v1 = acc[mt][nt][2 * hf + 1] + bias2[nt][1];
                if (p.act == 1) { v0 = silu_f(v0); v1 = silu_f(v1); }
                if (p.res != nullptr && ox < p.Wo) {
                    const float2 rf = __half22float2(*reinterpret_cast<const __half2*>(p.res + (row_pix + ox) * p.ldr + nt * 8 + 2 * t));
                    v0 += rf.x;
                    v1 += rf.y;
                }
                *reinterpret_cast<__half2*>(stg + r * Cfg::SP + nt * 8 + 2 * t) = __floats2half2_rn(v0, v1);
            }
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < (16 + RPI - 1) / RPI; ++it) {
            const int r = it * RPI + rsub;
            const int ox = ox0 + mt * 16 + r;
            if (r < 16 && ox < p.Wo)
                *reinterpret_cast<uint4*>(p.out + (row_pix + ox) * p.ldo + ch * 8) = *reinterpret_cast<const uint4*>(stg + r * Cfg::SP + ch * 8);
        }
        __syncwarp();
    }
}

template <int CIN, int S, int COUT>
static int small_conv_launch(const SmallConvParams& p, cudaStream_t st) {
    using Cfg = SmallConvCfg<CIN, S, COUT>;
    const size_t smem = (size_t)Cfg::PATCH_BYTES + (((size_t)COUT * (p.Kpad + 8) * 2 + 127) & ~(size_t)127) + Cfg::STG_BYTES;
    auto kern = small_conv3_kernel<CIN, S, COUT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ym_set_error("small_conv: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e)); return YM_ERR_CUDA; }
    }
    const int tiles_y = (p.Ho + Cfg::TH - 1) / Cfg::TH;
    launch_pdl(kern, dim3(p.tiles_x * tiles_y, p.B), 256, smem, st, p);
    return YM_OK;
}

static int g_small_conv_impl = 1;

// 1 when ym_small_conv3 takes the layer (ym_conv2d_nhwc asks before falling back to the implicit GEMM)
int small_conv3_supported(int Cin, int Cout, int KH, int KW, int stride, int pad, int Kpad, int out_f32, int B) {
    return g_small_conv_impl == 1 && KH == 3 && KW == 3 && pad == 1 && (stride == 1 || stride == 2) && (Cin == 8 || Cin == 16) &&
           (Cout == 8 || Cout == 16 || Cout == 32) && !out_f32 && Kpad >= ((9 * (Cin / 8) + 1) / 2) * 16 && Kpad % 8 == 0 && B <= 65535;
}

int small_conv3_run(const void* x, int ldx, int B, int H, int W, int Cin, const void* w, int Kpad, const float* bias, int Cout, int stride,
                    void* out, int ldo, const void* res, int ldr, int act, cudaStream_t st) {
    SmallConvParams p;
    p.x = (const __half*)x; p.ldx = ldx; p.w = (const __half*)w; p.Kpad = Kpad; p.bias = bias;
    p.out = (__half*)out; p.ldo = ldo; p.res = (const __half*)res; p.ldr = ldr;
    p.B = B; p.H = H; p.W = W; p.Ho = (H + 2 - 3) / stride + 1; p.Wo = (W + 2 - 3) / stride + 1;
    p.tiles_x = (p.Wo + 31) / 32; p.act = act;
#define YM_SC(CI, S_, CO) return small_conv_launch<CI, S_, CO>(p, st)
    if (Cin == 16 && stride == 1) { if (Cout == 8) YM_SC(16, 1, 8); if (Cout == 16) YM_SC(16, 1, 16); YM_SC(16, 1, 32); }
    if (Cin == 16 && stride == 2) { if (Cout == 8) YM_SC(16, 2, 8); if (Cout == 16) YM_SC(16, 2, 16); YM_SC(16, 2, 32); }
    if (Cin == 8 && stride == 1) { if (Cout == 8) YM_SC(8, 1, 8); if (Cout == 16) YM_SC(8, 1, 16); YM_SC(8, 1, 32); }
    if (Cout == 8) YM_SC(8, 2, 8);
    if (Cout == 16) YM_SC(8, 2, 16);
    YM_SC(8, 2, 32);
#undef YM_SC
}

}  // namespace ym

// 0 = every small 3x3 layer stays on the implicit-GEMM kernel of ym_conv2d_nhwc, 1 = patch-staged kernel above (default).  Returns the
// previous setting (A/B measurements and tests).
extern "C" int ym_set_small_conv_impl(int impl) {
    const int old = ym::g_small_conv_impl;
    if (impl == 0 || impl == 1) ym::g_small_conv_impl = impl;
    return old;
}
