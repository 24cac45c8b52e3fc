// Depthwise 7x7 (stride 1, pad 3) on the tensor cores: the `pe` positional convolutions of the area-attention blocks (block.py:1688,1731;
// :1311,1331) and every other 7x7 DWConv with C % 32 == 0.  Same entry point (ym_dwconv_nhwc), same weights ([49][C] fp16 tap-major).
//
// The FFMA kernel (elementwise.cu, dwconv_tiled_kernel<7, 8>) spends 49 FFMA + ~30 conversions / shared-memory reads per output and ran at
// 27 % of the FP32 peak (66 us per P3 launch, 231 us per yolo26-master-n step: profiles/r02_launch_roofline.txt).  A depthwise filter
// has no channel reduction to feed a GEMM - but it has a SPATIAL one: for one channel and one filter row ky,
//     out[y][x0 + m] += sum_k T[m][k] * in[y + ky - 3][x0 - 3 + k],      T[m][k] = w[ky][k - m]  (0 <= k - m < 7, else 0)
// is a 16 x 32 Toeplitz matrix times a 32 x (rows) matrix of input pixels, the SAME T for every output row y of the channel.  So per
// channel:  D[16 px][8 rows] += T_ky[16][32] * In_ky[32][8 rows]  over the 7 filter rows = 14 mma.sync.m16n8k16 for 128 outputs x 49
// taps (196 FFMA warp-instructions of arithmetic in the other kernel).  fp16 operands, fp32 accumulation: the same arithmetic precision.
//   * A (Toeplitz) fragments: element (m, k) depends on k - m only, so a thread's registers are pairs (w[d], w[d+1]) at the five offsets
//     d0 - 8, d0, d0 + 8, d0 + 16, d0 + 24 (d0 = 2t - g) per filter row - and since the filter is 7 wide exactly ONE of them overlaps it.
//     The CTA builds the eight non-trivial pairs (d = -1 .. 6) of every (channel, filter row) once in shared memory (7 KB); a lane reads
//     its one word per filter row and places it with selects (which of the five slots is a property of the lane, not of the channel).
//   * B (input) fragments: the CTA stages its haloed input tile TRANSPOSED, [channel][row][x] with x contiguous (16-byte global loads,
//     2-byte shared-memory scatter), so one ldmatrix.x4 delivers both k-steps of a filter row for 8 output rows; the fourth 8-column
//     segment (k 24..31) points at a zero strip, columns 22-23 of each row are zero.
//   * epilogue: accumulators + bias -> fp32 staging tile [row][px][channel] -> SiLU / residual / one rounding -> 16-byte NHWC stores.
// CTA = 16 x 16 output pixels x 32 channels, warp w = channels 4w .. 4w+3.
#include "ym_common.cuh"

namespace ym {

constexpr int DT_TW = 16, DT_TH = 16, DT_CB = 32, DT_R = 3, DT_K = 7;  // (TH = 8 fits four CTAs per SM but loads 2.4x instead of 1.9x the tile: 68 us against 56)
constexpr int DT_HH = DT_TH + 2 * DT_R;          // 22 input rows
constexpr int DT_HW = DT_TW + 2 * DT_R;          // 22 input columns
constexpr int DT_RP = 24;                        // row pitch in halves (48 B): 22 pixels + 2 zeros; 8 rows fall in 8 distinct 16-byte bank groups
constexpr int DT_CH_HALVES = DT_HH * DT_RP;      // 528 halves per channel
constexpr int DT_PAIRS = 8;                      // pairs (w[d], w[d+1]), d = -1 .. 6, per (channel, filter row)
constexpr int DT_PIXP = 36;                      // staging: floats per pixel (32 channels + 4)
constexpr int DT_ROWP = DT_TW * DT_PIXP + 4;     // staging: floats per row (580)

struct DwTcParams {
    const __half* __restrict__ x;   int ldx;
    int grp_w, grp_stride, grp_off;
    const __half* __restrict__ w;                 // [49][C]
    const float* __restrict__ bias;
    const __half* __restrict__ add; int ldadd;
    __half* __restrict__ out;       int ldo;
    int H, W, C, act, tiles_x;
};

__global__ void __launch_bounds__(256, 2) dwconv7_tc_kernel(const DwTcParams p) {
    extern __shared__ __align__(16) unsigned char dt_smem[];
    __half* sIn = reinterpret_cast<__half*>(dt_smem);                                   // [32 ch][22 rows][24]
    __half* sW = sIn + DT_CB * DT_CH_HALVES;                                            // [49][32]
    uint32_t* sTab = reinterpret_cast<uint32_t*>(sW + 49 * DT_CB);                      // [32 ch][7 ky][8 pairs]
    float* sOut = reinterpret_cast<float*>(sTab + DT_CB * DT_K * DT_PAIRS);             // [16 rows][580]
    __half* sZero = reinterpret_cast<__half*>(sOut + DT_TH * DT_ROWP);                   // 8 rows x 16 B of zeros (one ldmatrix segment)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, b = blockIdx.y, cb0 = blockIdx.z * DT_CB;
    const int ty0 = (tile / p.tiles_x) * DT_TH, tx0 = (tile % p.tiles_x) * DT_TW;

    // weights of the CTA's 32 channels (independent of the producer grid) and the zero strips
    for (int i = tid; i < 49 * (DT_CB / 8); i += 256) {
        const int tap = i / (DT_CB / 8), c8 = i - tap * (DT_CB / 8);
        *reinterpret_cast<uint4*>(sW + tap * DT_CB + c8 * 8) = *reinterpret_cast<const uint4*>(p.w + (long long)tap * p.C + cb0 + c8 * 8);
    }
    if (tid < 64) reinterpret_cast<uint32_t*>(sZero)[tid % 32] = 0u;
    for (int i = tid; i < DT_CB * DT_HH; i += 256)                                       // columns 22, 23 of every staged row
        *reinterpret_cast<uint32_t*>(sIn + i * DT_RP + DT_HW) = 0u;
    pdl_prologue();

    // ---- haloed input tile, transposed to [channel][row][x].  Item = (8-channel group, pixel) with the pixel fastest: the 2-byte scatter of a
    // warp then lands on consecutive x of one row and channel (conflict-free); the 16-byte global reads of a warp touch 32 pixels.
    {
        const __half* xb = p.x + (long long)b * p.H * p.W * p.ldx;
        constexpr int NPIX = DT_HH * DT_HW, ITEMS = NPIX * (DT_CB / 8), ITERS = (ITEMS + 255) / 256;
        // all (up to eight) 16-byte loads of a thread are in flight before the first scatter: behind one another each paid the full L2 / HBM
        // latency (the first version of this kernel was no faster than the FFMA one for that reason alone)
        uint4 v[ITERS];
#pragma unroll
        for (int k = 0; k < ITERS; ++k) {
            const int i = tid + k * 256;
            const int c8 = i / NPIX, pp = i - c8 * NPIX;
            const int hy = pp / DT_HW, hx = pp - hy * DT_HW;
            const int iy = ty0 + hy - DT_R, ix = tx0 + hx - DT_R;
            v[k] = make_uint4(0u, 0u, 0u, 0u);
            if (i < ITEMS && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
                const int c = cb0 + c8 * 8;
                const int csrc = (c / p.grp_w) * p.grp_stride + p.grp_off + (c % p.grp_w);
                v[k] = *reinterpret_cast<const uint4*>(xb + ((long long)iy * p.W + ix) * p.ldx + csrc);
            }
        }
#pragma unroll
        for (int k = 0; k < ITERS; ++k) {
            const int i = tid + k * 256;
            if (i < ITEMS) {
                const int c8 = i / NPIX, pp = i - c8 * NPIX;
                const int hy = pp / DT_HW, hx = pp - hy * DT_HW;
                const __half* hv = reinterpret_cast<const __half*>(&v[k]);
                __half* dst = sIn + (c8 * 8) * DT_CH_HALVES + hy * DT_RP + hx;
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j * DT_CH_HALVES] = hv[j];
            }
        }
    }
    __syncthreads();                                              // sW is complete: build the pair table
    for (int i = tid; i < DT_CB * DT_K * DT_PAIRS; i += 256) {
        const int cl = i / (DT_K * DT_PAIRS), rem = i - cl * (DT_K * DT_PAIRS);
        const int ky = rem / DT_PAIRS, d = rem - ky * DT_PAIRS - 1;             // d = -1 .. 6
        const unsigned short lo = d >= 0 ? __half_as_ushort(sW[(ky * DT_K + d) * DT_CB + cl]) : (unsigned short)0;
        const unsigned short hi = d + 1 < DT_K ? __half_as_ushort(sW[(ky * DT_K + d + 1) * DT_CB + cl]) : (unsigned short)0;
        sTab[i] = (uint32_t)lo | ((uint32_t)hi << 16);
    }
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    // the lane's five pair offsets are d0 - 8 + 8 j (j = 0 .. 4), d0 = 2t - g in [-7, 6]: exactly one lies in [-1, 6]
    const int d0 = 2 * t - g;
    const int jstar = d0 >= -1 ? 1 : 2;                            // d0 in [-1, 6] -> slot 1 (d0 itself); d0 in [-7, -2] -> slot 2 (d0 + 8)
    const int dstar = d0 >= -1 ? d0 : d0 + 8;                      // in [-1, 6]
    // ldmatrix row address of this lane: matrix = lane / 8 (k segment), row = lane % 8
    const int seg = lane >> 3, lrow = lane & 7;

#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
        const int cl = warp * 4 + cc;                               // CTA-local channel
        uint32_t aw[DT_K][5];                                       // pairs at d0-8, d0, d0+8, d0+16, d0+24
#pragma unroll
        for (int ky = 0; ky < DT_K; ++ky) {
            const uint32_t wv = sTab[(cl * DT_K + ky) * DT_PAIRS + dstar + 1];
            aw[ky][0] = 0u;
            aw[ky][1] = jstar == 1 ? wv : 0u;
            aw[ky][2] = jstar == 2 ? wv : 0u;
            aw[ky][3] = 0u;
            aw[ky][4] = 0u;
        }
        const float bias = p.bias != nullptr ? p.bias[cb0 + cl] : 0.f;
        const uint32_t in_base = smem_u32(sIn + cl * DT_CH_HALVES);
        const uint32_t zero_addr = smem_u32(sZero) + (uint32_t)(lrow * 16);
#pragma unroll
        for (int nb = 0; nb < DT_TH / 8; ++nb) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ky = 0; ky < DT_K; ++ky) {
                const uint32_t addr = seg < 3 ? in_base + (uint32_t)(((nb * 8 + lrow + ky) * DT_RP + seg * 8) * 2) : zero_addr;
                uint32_t b0, b1, b2, b3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(addr));
                const uint32_t a_lo[4] = {aw[ky][1], aw[ky][0], aw[ky][2], aw[ky][1]};      // k 0..15: d0, d0-8, d0+8, d0
                const uint32_t a_hi[4] = {aw[ky][3], aw[ky][2], aw[ky][4], aw[ky][3]};      // k 16..31: d0+16, d0+8, d0+24, d0+16
                mma_16816(acc, a_lo, b0, b1);
                mma_16816(acc, a_hi, b2, b3);
            }
            // D: (px g, rows 2t, 2t+1), (px g+8, rows 2t, 2t+1) of this 8-row block
            float* so = sOut + (nb * 8 + 2 * t) * DT_ROWP + cl;
            so[g * DT_PIXP] = acc[0] + bias;
            so[DT_ROWP + g * DT_PIXP] = acc[1] + bias;
            so[(g + 8) * DT_PIXP] = acc[2] + bias;
            so[DT_ROWP + (g + 8) * DT_PIXP] = acc[3] + bias;
        }
    }
    __syncthreads();

    // ---- staging tile -> SiLU / residual / one rounding -> NHWC
    for (int i = tid; i < DT_TH * DT_TW * (DT_CB / 8); i += 256) {
        const int c8 = i & 3, pp = i >> 2;
        const int row = pp / DT_TW, px = pp - row * DT_TW;
        const int oy = ty0 + row, ox = tx0 + px;
        if (oy >= p.H || ox >= p.W) continue;
        const float* src = sOut + row * DT_ROWP + px * DT_PIXP + c8 * 8;
        const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
        float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        if (p.act == 1) {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = silu_f(v[q]);
        }
        const long long pix = ((long long)b * p.H + oy) * p.W + ox;
        const int c = cb0 + c8 * 8;
        if (p.add != nullptr) {
            const Half8 av = *reinterpret_cast<const Half8*>(p.add + pix * p.ldadd + c);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 af = __half22float2(av.v[q]);
                v[2 * q] += af.x;
                v[2 * q + 1] += af.y;
            }
        }
        Half8 o;
#pragma unroll
        for (int q = 0; q < 4; ++q) o.v[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
        *reinterpret_cast<Half8*>(p.out + pix * p.ldo + c) = o;
    }
}

// Measured at bs32 (profiles/r02_dwconv_ncu.txt, profiles/r02_op_bench.json): 25.8 M instructions against the FFMA kernel's 43.6 M, but the same
// 56 us per P3 launch - the time went from the FFMAs into the TRANSPOSING load of the input tile ([row][x][channel] in memory, [channel][row][x]
// for ldmatrix): eight 2-byte scatters per 16-byte load, and the tile is read 1.9x with its halo.  An 8-row tile (four CTAs per SM instead of
// two) reads 2.4x and takes 68 us: the kernel scales with the bytes it transposes, not with its occupancy.  Kept selectable (ym_set_dwconv_tc); the FFMA kernel stays the default.
static int g_dwconv_tc = 0;

int dwconv7_tc_supported(int C, int ksize, int grp_w, int B, const void* route_idx) {
    return g_dwconv_tc == 1 && ksize == 7 && C % DT_CB == 0 && grp_w % 8 == 0 && route_idx == nullptr && B <= 65535 && C / DT_CB <= 65535;
}

int dwconv7_tc_run(const void* x, int ldx, int grp_w, int grp_stride, int grp_off, const void* w, const float* bias, int B, int H, int W, int C,
                   int act, const void* add, int ldadd, void* out, int ldo, cudaStream_t st) {
    DwTcParams p;
    p.x = (const __half*)x; p.ldx = ldx; p.grp_w = grp_w; p.grp_stride = grp_stride; p.grp_off = grp_off;
    p.w = (const __half*)w; p.bias = bias; p.add = (const __half*)add; p.ldadd = ldadd; p.out = (__half*)out; p.ldo = ldo;
    p.H = H; p.W = W; p.C = C; p.act = act; p.tiles_x = (W + DT_TW - 1) / DT_TW;
    const int tiles_y = (H + DT_TH - 1) / DT_TH;
    const size_t smem = (size_t)DT_CB * DT_CH_HALVES * 2 + 49 * DT_CB * 2 + (size_t)DT_CB * DT_K * DT_PAIRS * 4 + (size_t)DT_TH * DT_ROWP * 4 + 128;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(dwconv7_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ym_set_error("dwconv7_tc: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e)); return YM_ERR_CUDA; }
        attr_set = true;
    }
    launch_pdl(dwconv7_tc_kernel, dim3(p.tiles_x * tiles_y, B, C / DT_CB), 256, smem, st, p);
    return YM_OK;
}

}  // namespace ym

// 1 = depthwise 7x7 layers with C % 32 == 0 run on the mma.sync Toeplitz kernel, 0 = the FFMA kernel (default).  Returns the previous setting.
extern "C" int ym_set_dwconv_tc(int on) {
    const int old = ym::g_dwconv_tc;
    if (on == 0 || on == 1) ym::g_dwconv_tc = on;
    return old;
}
